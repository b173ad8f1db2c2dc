// Host-side restatement of the System.Numerics operations the reference's
// scene set-up uses.  The .NET 9 BCL source is not part of
// /root/reference; the conventions below (FMA chains in Vector3.Transform /
// Matrix4x4 multiply when BN_NET9_FMA is on) are the documented assumption.
// Compile with -ffp-contract=off: every fused op here is an explicit fmaf.
#pragma once
#include <cmath>
#include <cstring>

#ifndef BN_NET9_FMA
#define BN_NET9_FMA 1
#endif

namespace bnhost {

struct V3 {
  float x, y, z;
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
// Vector3.MinNative / MaxNative: x64 minps/maxps, `a < b ? a : b`
inline V3 min_native(V3 a, V3 b) { return {a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
inline V3 max_native(V3 a, V3 b) { return {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }

inline V3 cross(V3 a, V3 b) {
#if BN_NET9_FMA
  return {fmaf(-a.z, b.y, a.y * b.z), fmaf(-a.x, b.z, a.z * b.x), fmaf(-a.y, b.x, a.x * b.y)};
#else
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
#endif
}
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }

// Matrix4x4: row-major M11..M44, row-vector convention (v' = v*M).
struct M4 {
  float m[16];
  static M4 identity() {
    M4 r;
    std::memset(r.m, 0, sizeof r.m);
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f;
    return r;
  }
};

inline M4 create_scale(float x, float y, float z) {
  M4 r = M4::identity();
  r.m[0] = x; r.m[5] = y; r.m[10] = z;
  return r;
}
inline M4 create_translation(float x, float y, float z) {
  M4 r = M4::identity();
  r.m[12] = x; r.m[13] = y; r.m[14] = z;
  return r;
}
// Matrix4x4.CreateRotationX/Y/Z (radians; Loader.fs:29-31)
inline M4 create_rotation_x(float a) {
  float c = cosf(a), s = sinf(a);
  M4 r = M4::identity();
  r.m[5] = c; r.m[6] = s; r.m[9] = -s; r.m[10] = c;
  return r;
}
inline M4 create_rotation_y(float a) {
  float c = cosf(a), s = sinf(a);
  M4 r = M4::identity();
  r.m[0] = c; r.m[2] = -s; r.m[8] = s; r.m[10] = c;
  return r;
}
inline M4 create_rotation_z(float a) {
  float c = cosf(a), s = sinf(a);
  M4 r = M4::identity();
  r.m[0] = c; r.m[1] = s; r.m[4] = -s; r.m[5] = c;
  return r;
}

// left * right: result.row_i = left.row_i transformed by `right`
inline M4 mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
#if BN_NET9_FMA
      float v = b.m[j] * a.m[i * 4];
      v = fmaf(b.m[4 + j], a.m[i * 4 + 1], v);
      v = fmaf(b.m[8 + j], a.m[i * 4 + 2], v);
      v = fmaf(b.m[12 + j], a.m[i * 4 + 3], v);
#else
      float v = ((a.m[i * 4] * b.m[j] + a.m[i * 4 + 1] * b.m[4 + j]) + a.m[i * 4 + 2] * b.m[8 + j]) + a.m[i * 4 + 3] * b.m[12 + j];
#endif
      r.m[i * 4 + j] = v;
    }
  }
  return r;
}

// Vector3.Transform(position, M)
inline V3 transform_point(V3 p, const M4& M) {
  float r[3];
  for (int j = 0; j < 3; ++j) {
#if BN_NET9_FMA
    float v = M.m[j] * p.x;
    v = fmaf(M.m[4 + j], p.y, v);
    v = fmaf(M.m[8 + j], p.z, v);
    r[j] = v + M.m[12 + j];
#else
    r[j] = ((p.x * M.m[j] + p.y * M.m[4 + j]) + p.z * M.m[8 + j]) + M.m[12 + j];
#endif
  }
  return {r[0], r[1], r[2]};
}

// Matrix4x4.Invert — general 4x4 inverse by cofactors (adjugate / det), the
// shape of the BCL's scalar fallback.  Host-side only: the oracle and the GPU
// both consume the matrices produced here, so rounding differences against the
// BCL's SIMD path cannot cause oracle-vs-GPU divergence.
inline bool invert(const M4& s, M4& out) {
  const float* q = s.m;
  float a = q[0], b = q[1], c = q[2], d = q[3];
  float e = q[4], f = q[5], g = q[6], h = q[7];
  float i = q[8], j = q[9], k = q[10], l = q[11];
  float m = q[12], n = q[13], o = q[14], p = q[15];

  float kp_lo = k * p - l * o, jp_ln = j * p - l * n, jo_kn = j * o - k * n;
  float ip_lm = i * p - l * m, io_km = i * o - k * m, in_jm = i * n - j * m;

  float a11 = +(f * kp_lo - g * jp_ln + h * jo_kn);
  float a12 = -(e * kp_lo - g * ip_lm + h * io_km);
  float a13 = +(e * jp_ln - f * ip_lm + h * in_jm);
  float a14 = -(e * jo_kn - f * io_km + g * in_jm);

  float det = a * a11 + b * a12 + c * a13 + d * a14;
  if (fabsf(det) < 1.401298464e-45f) {
    for (float& v : out.m) v = NAN;
    return false;
  }
  float inv = 1.0f / det;

  float gp_ho = g * p - h * o, fp_hn = f * p - h * n, fo_gn = f * o - g * n;
  float ep_hm = e * p - h * m, eo_gm = e * o - g * m, en_fm = e * n - f * m;
  float gl_hk = g * l - h * k, fl_hj = f * l - h * j, fk_gj = f * k - g * j;
  float el_hi = e * l - h * i, ek_gi = e * k - g * i, ej_fi = e * j - f * i;

  float* r = out.m;
  r[0] = a11 * inv;
  r[4] = a12 * inv;
  r[8] = a13 * inv;
  r[12] = a14 * inv;
  r[1] = -(b * kp_lo - c * jp_ln + d * jo_kn) * inv;
  r[5] = +(a * kp_lo - c * ip_lm + d * io_km) * inv;
  r[9] = -(a * jp_ln - b * ip_lm + d * in_jm) * inv;
  r[13] = +(a * jo_kn - b * io_km + c * in_jm) * inv;
  r[2] = +(b * gp_ho - c * fp_hn + d * fo_gn) * inv;
  r[6] = -(a * gp_ho - c * ep_hm + d * eo_gm) * inv;
  r[10] = +(a * fp_hn - b * ep_hm + d * en_fm) * inv;
  r[14] = -(a * fo_gn - b * eo_gm + c * en_fm) * inv;
  r[3] = -(b * gl_hk - c * fl_hj + d * fk_gj) * inv;
  r[7] = +(a * gl_hk - c * el_hi + d * ek_gi) * inv;
  r[11] = -(a * fl_hj - b * el_hi + d * ej_fi) * inv;
  r[15] = +(a * fk_gj - b * ek_gi + c * ej_fi) * inv;
  return true;
}

struct AABB {
  V3 lo, hi;
  static AABB empty() { return {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; }
  V3 centroid() const { return 0.5f * (lo + hi); }
  V3 diagonal() const { return hi - lo; }
  float surface_area() const {
    V3 d = diagonal();
    return 2.f * (d.x * d.y + d.y * d.z + d.z * d.x);
  }
  // Util/BVH.fs:24-27
  int split_axis() const {
    V3 d = diagonal();
    if (d.x >= d.y && d.x >= d.z) return 0;
    return d.y >= d.z ? 1 : 2;
  }
};
inline AABB unite(const AABB& a, const AABB& b) { return {min_native(a.lo, b.lo), max_native(a.hi, b.hi)}; }
inline AABB unite(const AABB& a, V3 p) { return {min_native(a.lo, p), max_native(a.hi, p)}; }

// AxisAlignedBoundingBox.Transform (Util/BVH.fs:29-40): 8 corners
inline AABB transform_aabb(const AABB& box, const M4& M) {
  AABB r = AABB::empty();
  for (int i = 0; i < 8; ++i) {
    V3 c{(i & 1) == 0 ? box.lo.x : box.hi.x, (i & 2) == 0 ? box.lo.y : box.hi.y, (i & 4) == 0 ? box.lo.z : box.hi.z};
    V3 p = transform_point(c, M);
    r.lo = min_native(r.lo, p);
    r.hi = max_native(r.hi, p);
  }
  return r;
}

// ---------------------------------------------------------------------------
// Keyframe interpolation (Base/Scene.fs:9-20, Util/Transform.fs) — SURVEY §8(f) N4.
// KeyFrame.Interpolate calls into the .NET 9 BCL: Matrix4x4.Decompose, Vector3.Lerp,
// Quaternion.Slerp, Matrix4x4.CreateFromQuaternion/CreateTranslation/CreateScale.  The BCL is
// not in this image, so what follows restates those routines' published algorithms (the
// DirectXMath-style decomposition, the textbook slerp with a 1e-6 linear fallback); fp32 rounding
// against a real .NET host is UNPINNED — in the real integration the managed host evaluates
// Transform.Eval itself and hands over finished matrices.
// ---------------------------------------------------------------------------
struct Quat { float x, y, z, w; };

inline V3 normalize_v(V3 a) { float l = length(a); return {a.x / l, a.y / l, a.z / l}; }

// Vector3.Lerp: value1 * (1 - amount) + value2 * amount
inline V3 lerp(V3 a, V3 b, float t) {
  float s = 1.0f - t;
#if BN_NET9_FMA
  return {fmaf(a.x, s, b.x * t), fmaf(a.y, s, b.y * t), fmaf(a.z, s, b.z * t)};
#else
  return {a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t};
#endif
}

// Matrix4x4.GetDeterminant
inline float determinant(const M4& s) {
  const float* q = s.m;
  float a = q[0], b = q[1], c = q[2], d = q[3];
  float e = q[4], f = q[5], g = q[6], h = q[7];
  float i = q[8], j = q[9], k = q[10], l = q[11];
  float m = q[12], n = q[13], o = q[14], p = q[15];
  float kp_lo = k * p - l * o, jp_ln = j * p - l * n, jo_kn = j * o - k * n;
  float ip_lm = i * p - l * m, io_km = i * o - k * m, in_jm = i * n - j * m;
  return a * (f * kp_lo - g * jp_ln + h * jo_kn) - b * (e * kp_lo - g * ip_lm + h * io_km) +
         c * (e * jp_ln - f * ip_lm + h * in_jm) - d * (e * jo_kn - f * io_km + g * in_jm);
}

// Quaternion.CreateFromRotationMatrix
inline Quat quat_from_rotation_matrix(const M4& M) {
  const float m11 = M.m[0], m12 = M.m[1], m13 = M.m[2], m21 = M.m[4], m22 = M.m[5], m23 = M.m[6], m31 = M.m[8], m32 = M.m[9], m33 = M.m[10];
  float trace = m11 + m22 + m33;
  Quat q;
  if (trace > 0.0f) {
    float s = sqrtf(trace + 1.0f);
    q.w = s * 0.5f;
    s = 0.5f / s;
    q.x = (m23 - m32) * s; q.y = (m31 - m13) * s; q.z = (m12 - m21) * s;
  } else if (m11 >= m22 && m11 >= m33) {
    float s = sqrtf(1.0f + m11 - m22 - m33), inv = 0.5f / s;
    q.x = 0.5f * s; q.y = (m12 + m21) * inv; q.z = (m13 + m31) * inv; q.w = (m23 - m32) * inv;
  } else if (m22 > m33) {
    float s = sqrtf(1.0f + m22 - m11 - m33), inv = 0.5f / s;
    q.x = (m21 + m12) * inv; q.y = 0.5f * s; q.z = (m32 + m23) * inv; q.w = (m31 - m13) * inv;
  } else {
    float s = sqrtf(1.0f + m33 - m11 - m22), inv = 0.5f / s;
    q.x = (m31 + m13) * inv; q.y = (m32 + m23) * inv; q.z = 0.5f * s; q.w = (m12 - m21) * inv;
  }
  return q;
}

// Matrix4x4.Decompose: scale = row lengths, basis vectors re-orthonormalised in order of
// decreasing scale (degenerate rows replaced), handedness fixed on the largest axis, rotation
// = identity when the normalised basis is not a rotation (|det| - 1)^2 > 1e-4.  The result flag
// is ignored by Transform.Decompose (Util/Transform.fs:11-16), as here.
inline bool decompose(const M4& M, V3& scale, Quat& rotation, V3& translation) {
  const float eps = 0.0001f;
  translation = {M.m[12], M.m[13], M.m[14]};
  V3 basis[3] = {{M.m[0], M.m[1], M.m[2]}, {M.m[4], M.m[5], M.m[6]}, {M.m[8], M.m[9], M.m[10]}};
  const V3 canonical[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  float sc[3] = {length(basis[0]), length(basis[1]), length(basis[2])};
  int a, b, c;
  const float x = sc[0], y = sc[1], z = sc[2];
  if (x < y) {
    if (y < z) { a = 2; b = 1; c = 0; }
    else { a = 1; if (x < z) { b = 2; c = 0; } else { b = 0; c = 2; } }
  } else {
    if (x < z) { a = 2; b = 0; c = 1; }
    else { a = 0; if (y < z) { b = 2; c = 1; } else { b = 1; c = 2; } }
  }
  if (sc[a] < eps) basis[a] = canonical[a];
  basis[a] = normalize_v(basis[a]);
  if (sc[b] < eps) {
    const float ax = fabsf(basis[a].x), ay = fabsf(basis[a].y), az = fabsf(basis[a].z);
    int cc;
    if (ax < ay) { if (ay < az) cc = 0; else cc = ax < az ? 0 : 2; }
    else { if (ax < az) cc = 1; else cc = ay < az ? 1 : 2; }
    basis[b] = cross(basis[a], canonical[cc]);
  }
  basis[b] = normalize_v(basis[b]);
  if (sc[c] < eps) basis[c] = cross(basis[a], basis[b]);
  basis[c] = normalize_v(basis[c]);
  M4 T = M4::identity();
  for (int r = 0; r < 3; ++r) { T.m[4 * r] = basis[r].x; T.m[4 * r + 1] = basis[r].y; T.m[4 * r + 2] = basis[r].z; }
  float det = determinant(T);
  if (det < 0.0f) {
    sc[a] = -sc[a];
    basis[a] = {-basis[a].x, -basis[a].y, -basis[a].z};
    T.m[4 * a] = basis[a].x; T.m[4 * a + 1] = basis[a].y; T.m[4 * a + 2] = basis[a].z;
    det = -det;
  }
  det -= 1.0f;
  det *= det;
  scale = {sc[0], sc[1], sc[2]};
  if (eps < det) { rotation = {0.f, 0.f, 0.f, 1.f}; return false; }
  rotation = quat_from_rotation_matrix(T);
  return true;
}

// Quaternion.Slerp
inline Quat slerp(Quat q1, Quat q2, float t) {
  float cos_omega = q1.x * q2.x + q1.y * q2.y + q1.z * q2.z + q1.w * q2.w;
  bool flip = false;
  if (cos_omega < 0.0f) { flip = true; cos_omega = -cos_omega; }
  float s1, s2;
  if (cos_omega > 1.0f - 1e-6f) {
    s1 = 1.0f - t;
    s2 = flip ? -t : t;
  } else {
    float omega = acosf(cos_omega), inv_sin = 1.0f / sinf(omega);
    s1 = sinf((1.0f - t) * omega) * inv_sin;
    s2 = flip ? -sinf(t * omega) * inv_sin : sinf(t * omega) * inv_sin;
  }
  return {s1 * q1.x + s2 * q2.x, s1 * q1.y + s2 * q2.y, s1 * q1.z + s2 * q2.z, s1 * q1.w + s2 * q2.w};
}

// Matrix4x4.CreateFromQuaternion
inline M4 create_from_quaternion(Quat q) {
  float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  float xy = q.x * q.y, wz = q.z * q.w, xz = q.z * q.x, wy = q.y * q.w, yz = q.y * q.z, wx = q.x * q.w;
  M4 r = M4::identity();
  r.m[0] = 1.0f - 2.0f * (yy + zz); r.m[1] = 2.0f * (xy + wz); r.m[2] = 2.0f * (xz - wy);
  r.m[4] = 2.0f * (xy - wz); r.m[5] = 1.0f - 2.0f * (zz + xx); r.m[6] = 2.0f * (yz + wx);
  r.m[8] = 2.0f * (xz + wy); r.m[9] = 2.0f * (yz - wx); r.m[10] = 1.0f - 2.0f * (yy + xx);
  return r;
}

// Transform.Compose (Util/Transform.fs:6-9) AS WRITTEN: `Compose(s, r, t)` builds
// CreateTranslation(s) * CreateFromQuaternion(r) * CreateScale(t) — the scale vector is used as a
// translation and the translation vector as a scale (the parameter names and the factory calls are
// crossed in the reference).  Restated, not repaired: an interpolated transform is NOT the
// keyframe matrix even at ratio 0.
inline M4 compose_as_written(V3 s, Quat r, V3 t) {
  return mul(mul(create_translation(s.x, s.y, s.z), create_from_quaternion(r)), create_scale(t.x, t.y, t.z));
}

// KeyFrame.Interpolate (Base/Scene.fs:13-20)
inline M4 interpolate_keyframes(float time_a, const M4& A, float time_b, const M4& B, float t) {
  const float ratio = (t - time_a) / (time_b - time_a);
  V3 s1, t1, s2, t2;
  Quat r1, r2;
  decompose(A, s1, r1, t1);
  decompose(B, s2, r2, t2);
  const V3 translation = lerp(t1, t2, ratio);
  const Quat rotation = slerp(r1, r2, ratio);
  const V3 scale = lerp(s1, s2, ratio);
  return compose_as_written(scale, rotation, translation);
}

}  // namespace bnhost
