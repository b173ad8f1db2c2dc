// Device-side restatement of the System.Numerics / MathF operations on the hot
// path.  This translation unit is compiled with
// -fmad=false (no contraction), IEEE div/sqrt and denormals kept:
// every fused operation below is an explicit __fmaf_rn, exactly where the
// reference calls FusedMultiplyAdd or where .NET 9's Vector3.Cross/Transform
// fuse (BN_NET9_FMA).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#ifdef BN_EXP_SHARED_RCP
#include "../../../include/bn_portable_math.h"
#endif

#ifndef BN_NET9_FMA
#define BN_NET9_FMA 1
#endif

#define BN_DEV __device__ __forceinline__

namespace bn {

constexpr float kPi = 3.14159274101257324f;              // MathF.PI
constexpr float kSingleEpsilon = 1.401298464324817e-45f;  // Single.Epsilon

BN_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
BN_DEV float3 splat(float s) { return make_float3(s, s, s); }
BN_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
BN_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
BN_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
BN_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
BN_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
BN_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
// a / s, IEEE round-to-nearest (what `/` compiles to here), with one case kept off nvcc's slow path:
// a zero numerator fails the FCHK range test and sends the whole warp through the out-of-line
// special-case routine, and axis-aligned geometry (Cornell-box normals and tangents) puts exact
// zeros into every normalize().  0 / s for a normal finite s is +-0 with the XOR of the signs, so
// that case is answered directly and the divider only ever sees a non-zero numerator.
BN_DEV float div_ieee(float a, float s) {
  const float as = fabsf(s);
  const bool zero_num = a == 0.f && as >= 1.17549435e-38f && as <= 3.402823466e38f;
  const float q = (zero_num ? 1.f : a) / s;
  return zero_num ? __uint_as_float((__float_as_uint(a) ^ __float_as_uint(s)) & 0x80000000u) : q;
}
#ifdef BN_EXP_SHARED_RCP
// experiment queued for the next GPU session (default off): the three divisions of a normalize() share one correctly
// rounded reciprocal (include/bn_portable_math.h: bn_div_by_rcp — same bits as the IEEE division on the guarded domain)
#ifdef BN_HOSTSIM
#define BN_NOINLINE_DEV static __attribute__((noinline))
#else
#define BN_NOINLINE_DEV static __device__ __noinline__
#endif
BN_NOINLINE_DEV float3 div3_ieee(float3 a, float s) { return f3(div_ieee(a.x, s), div_ieee(a.y, s), div_ieee(a.z, s)); }  // one cold copy
BN_DEV float3 operator/(float3 a, float s) {
  const bool ok = bn_div_rcp_ok(s) && (a.x == 0.f || bn_div_rcp_ok(a.x)) && (a.y == 0.f || bn_div_rcp_ok(a.y)) && (a.z == 0.f || bn_div_rcp_ok(a.z));
  if (!ok) return div3_ieee(a, s);
  const float r = __frcp_rn(s);
  return f3(bn_div_by_rcp(a.x, s, r), bn_div_by_rcp(a.y, s, r), bn_div_by_rcp(a.z, s, r));
}
#else
BN_DEV float3 operator/(float3 a, float s) { return f3(div_ieee(a.x, s), div_ieee(a.y, s), div_ieee(a.z, s)); }
#endif
// Vector3.FusedMultiplyAdd
BN_DEV float3 vfma(float3 a, float3 b, float3 c) { return f3(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y), __fmaf_rn(a.z, b.z, c.z)); }
// Vector3.Dot: ((x*x' + y*y') + z*z')
BN_DEV float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
BN_DEV float length_sq(float3 a) { return dot(a, a); }
BN_DEV float length(float3 a) { return __fsqrt_rn(dot(a, a)); }
BN_DEV float3 normalize(float3 a) { return a / length(a); }  // 3 IEEE divisions, as Vector3.Normalize
BN_DEV float3 cross(float3 a, float3 b) {
#if BN_NET9_FMA
  return f3(__fmaf_rn(-a.z, b.y, a.y * b.z), __fmaf_rn(-a.x, b.z, a.z * b.x), __fmaf_rn(-a.y, b.x, a.x * b.y));
#else
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
#endif
}
// Vector3.MinNative/MaxNative: minps/maxps (second operand on NaN)
BN_DEV float min_native(float a, float b) { return a < b ? a : b; }
BN_DEV float max_native(float a, float b) { return a > b ? a : b; }
BN_DEV float3 min_native(float3 a, float3 b) { return f3(min_native(a.x, b.x), min_native(a.y, b.y), min_native(a.z, b.z)); }
BN_DEV float3 max_native(float3 a, float3 b) { return f3(max_native(a.x, b.x), max_native(a.y, b.y), max_native(a.z, b.z)); }
// Math.Max / Math.Min / MathF.Max / MathF.Min: IEEE 754-2019 maximum/minimum
// (NaN-propagating, +0 > -0) — one instruction on sm_100.
#ifdef BN_HOSTSIM
// host build of these device functions for tests/hostsim (never part of the product): the same IEEE 754-2019
// maximum / minimum spelled out
BN_DEV float net_max(float a, float b) {
  if (a != a || b != b) return __uint_as_float(0x7fffffffu);  // max.NaN.f32 returns the canonical NaN
  if (a == b) return (__float_as_uint(a) & 0x80000000u) ? b : a;  // +0 > -0
  return a > b ? a : b;
}
BN_DEV float net_min(float a, float b) {
  if (a != a || b != b) return __uint_as_float(0x7fffffffu);
  if (a == b) return (__float_as_uint(a) & 0x80000000u) ? a : b;
  return a < b ? a : b;
}
#else
BN_DEV float net_max(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
BN_DEV float net_min(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
#endif

// 4x3 slice of a row-major Matrix4x4 (rows M1x..M4x, columns 1..3): all that
// Vector3.Transform reads.
struct Mat43 {
  float m[12];  // r0c0 r0c1 r0c2 | r1c0 ... | r3c0 r3c1 r3c2
};
BN_DEV Mat43 load_mat43(const float4* p) {
  float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  Mat43 M;
  M.m[0] = a.x; M.m[1] = a.y; M.m[2] = a.z; M.m[3] = a.w;
  M.m[4] = b.x; M.m[5] = b.y; M.m[6] = b.z; M.m[7] = b.w;
  M.m[8] = c.x; M.m[9] = c.y; M.m[10] = c.z; M.m[11] = c.w;
  return M;
}
// Vector3.Transform(position, M), row-vector convention
BN_DEV float3 transform_point(float3 p, const Mat43& M) {
#if BN_NET9_FMA
  float x = __fmaf_rn(M.m[6], p.z, __fmaf_rn(M.m[3], p.y, M.m[0] * p.x)) + M.m[9];
  float y = __fmaf_rn(M.m[7], p.z, __fmaf_rn(M.m[4], p.y, M.m[1] * p.x)) + M.m[10];
  float z = __fmaf_rn(M.m[8], p.z, __fmaf_rn(M.m[5], p.y, M.m[2] * p.x)) + M.m[11];
#else
  float x = ((p.x * M.m[0] + p.y * M.m[3]) + p.z * M.m[6]) + M.m[9];
  float y = ((p.x * M.m[1] + p.y * M.m[4]) + p.z * M.m[7]) + M.m[10];
  float z = ((p.x * M.m[2] + p.y * M.m[5]) + p.z * M.m[8]) + M.m[11];
#endif
  return f3(x, y, z);
}
// Transform(dir, M) - M.Translation
BN_DEV float3 transform_dir(float3 d, const Mat43& M) { return transform_point(d, M) - f3(M.m[9], M.m[10], M.m[11]); }

BN_DEV float3 point_at(float3 o, float3 d, float t) { return vfma(splat(t), d, o); }  // Ray.fs:16-17
BN_DEV float3 rcp3(float3 d) { return f3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z)); }  // Vector3.One / d

// AxisAlignedBoundingBox.Intersect (Ray.fs:29-39) split in two: the part that
// does not depend on the running closest t ...
struct Slab {
  float tmin;  // Max(1e-3, Max(lo.x, Max(lo.y, lo.z)))
  float thi;   // Min(hi.x, Min(hi.y, hi.z))
};
// FAST = false: the reference's operations one for one (minps/maxps selects,
// NaN-propagating Math.Max/Min).
// FAST = true: plain fmin/fmax.  Bit-identical to the exact form whenever no lane
// can be NaN, i.e. every component of 1/d is finite and non-zero and the box and
// origin are finite (then (p - o) * inv is never 0 * inf; sign-of-zero differences
// between minps and fmin cannot survive the Max with 1e-3 or change a comparison).
// The caller picks FAST per ray and per space with slab_fast_ok(); rays with a
// zero / denormal / infinite direction component take the exact path.
template <bool FAST>
BN_DEV Slab slab(float3 pmin, float3 pmax, float3 o, float3 inv) {
  const float3 t0 = (pmin - o) * inv;
  const float3 t1 = (pmax - o) * inv;
  Slab s;
  if (FAST) {
    s.tmin = fmaxf(fmaxf(1e-3f, fminf(t0.x, t1.x)), fmaxf(fminf(t0.y, t1.y), fminf(t0.z, t1.z)));
    s.thi = fminf(fmaxf(t0.x, t1.x), fminf(fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z)));
  } else {
    const float3 lo = min_native(t0, t1);
    const float3 hi = max_native(t0, t1);
    s.tmin = net_max(1e-3f, net_max(lo.x, net_max(lo.y, lo.z)));
    s.thi = net_min(hi.x, net_min(hi.y, hi.z));
  }
  return s;
}
BN_DEV bool slab_fast_ok(float3 o, float3 inv) {
  const float big = 3.0e38f;
  return fabsf(inv.x) < big && fabsf(inv.y) < big && fabsf(inv.z) < big && inv.x != 0.f && inv.y != 0.f && inv.z != 0.f &&
         fabsf(o.x) < 1.0e37f && fabsf(o.y) < 1.0e37f && fabsf(o.z) < 1.0e37f;
}
// ... and the part that does: tMin <= Min(t, thi).  For non-NaN t this equals
// (tmin <= t) && (tmin <= thi), which is what lets a deferred child be
// re-checked against the CURRENT t at pop time exactly as the reference does.
template <bool FAST>
BN_DEV bool slab_pass(const Slab& s, float t) { return FAST ? (s.tmin <= fminf(t, s.thi)) : (s.tmin <= net_min(t, s.thi)); }

// OrthonormalBasis (Primitive.fs:9-40)
struct Onb { float3 n, t, b; };
BN_DEV Onb onb_from_n(float3 n) {  // :15-23
  float3 axis = fabsf(n.x) > 0.1f ? f3(0.f, 1.f, 0.f) : f3(1.f, 0.f, 0.f);
  Onb f;
  f.n = n;
  f.t = normalize(cross(n, axis));
  f.b = cross(n, f.t);
  return f;
}
BN_DEV float3 local_to_world(const Onb& f, float3 v) { return (v.x * f.t + v.y * f.b) + v.z * f.n; }
BN_DEV float3 world_to_local(const Onb& f, float3 v) { return f3(dot(v, f.t), dot(v, f.b), dot(v, f.n)); }
BN_DEV Onb transform_onb(const Onb& f, const Mat43& M) {  // :34-38 (n' not renormalised)
  Onb r;
  r.t = normalize(transform_dir(f.t, M));
  r.b = normalize(transform_dir(f.b, M));
  r.n = cross(r.t, r.b);
  return r;
}

}  // namespace bn
