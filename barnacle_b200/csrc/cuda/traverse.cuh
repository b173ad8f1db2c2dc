// Two-level (TLAS + BLAS) BVH traversal: the sm_100a replacement of
//   BVHAggregate.Intersect/3,/2        Extensions/Aggregate/BVH.fs:37-58, 11-35
//   PrimitiveInstance.Intersect/3,/2   Base/Primitive.fs:118-129, 111-116
//   MeshPrimitive.Intersect/3,/2       Extensions/Primitive/Mesh.fs:217-242, 188-215
//   Triangle.Intersect/3,/2            Mesh.fs:50-82, 24-48
//   SpherePrimitive.Intersect/3,/2     Extensions/Primitive/Sphere.fs:35-77, 13-33
//
// Exactness contract: same tree topology, same
// front-to-back rule (left first iff dir[splitAxis] > 0), same slab / triangle
// arithmetic op for op, leaf items in slot order, strict `t' < t` acceptance —
// so the closest hit, including exact-t ties and NaN slab cases, is the one the
// reference returns.  What differs is the memory layout (child boxes in the
// parent, pre-gathered triangles, one unified stack) and that a deferred far
// child is re-checked with its stored entry distance instead of re-fetching
// its box; `tmin <= Min(t, thi)` == `(tmin <= t) && (tmin <= thi)`.
#pragma once
#include "device_scene.h"
#include "traverse_limits.h"
#include "vecmath.cuh"

namespace bn {

constexpr uint32_t kTlasBit = 0x40000000u;
constexpr uint32_t kIndexMask = 0x3FFFFFFFu;
constexpr uint32_t kFirstMask = 0x00FFFFFFu;

struct HitRec {
  float t;
  int inst;   // TLAS-order instance slot, -1 = miss
  int prim;   // BLAS-order triangle; spheres: 0 = near root (t0), 1 = far root (t1)
  float u, v; // barycentrics (triangles only)
};

BN_DEV uint32_t fbits(float f) { return __float_as_uint(f); }

// Triangle.Intersect — shared arithmetic of both overloads (Mesh.fs:24-82).
// Returns true and t' (+ u, v) iff the reference would accept against `t`.
BN_DEV bool tri_test(float3 p0, float3 p1, float3 p2, float3 o, float3 d, float t, float& tp, float& u, float& v) {
  float3 e0 = p1 - p0, e1 = p2 - p0;
  float3 rce1 = cross(d, e1);
  float det = dot(e0, rce1);
  if (fabsf(det) < kSingleEpsilon) return false;
  float inv = __frcp_rn(det);
  float3 s = o - p0;
  u = inv * dot(s, rce1);
  if (u < 0.f || u > 1.f) return false;
  float3 sce0 = cross(s, e0);
  v = inv * dot(d, sce0);
  if (v < 0.f || u + v > 1.f) return false;
  tp = inv * dot(e1, sce0);
  return tp > kSingleEpsilon && tp < t;
}

// SpherePrimitive.Intersect (Sphere.fs:13-77): returns 0 = miss, 1 = near root
// t0 accepted, 2 = far root t1 accepted.
BN_DEV int sphere_test(float radius, float3 o, float3 d, float t, float& tp) {
  const float eps = 1e-3f;
  float a = length_sq(d);
  float b = -dot(o, d);
  float r2 = radius * radius;
  float c = length_sq(o) - r2;
  float dd = r2 - length_sq(o + (b / a) * d);
  if (dd < 0.f) return 0;
  float q = b + copysignf(__fsqrt_rn(a * dd), b);
  float t0 = c / q;
  if (t0 > eps && t0 < t) { tp = t0; return 1; }
  float t1 = q / a;
  if (t1 > eps && t1 < t) { tp = t1; return 2; }
  return 0;
}

// t: in = tmax (closest: usually +inf), out = closest distance (closest only).
template <bool ANY>
BN_DEV bool trace(const DScene& sc, const float3 wo, const float3 wd, float& t, HitRec& hit) {
  uint32_t stk[kStackSize];
  float stkt[kStackSize];
  int sp = 0;

  const float3 winv = rcp3(wd);
  float3 o = wo, d = wd, inv = winv;  // ray in the CURRENT space (world or object)
  const GNode* nodes = sc.nodes;      // TLAS nodes start at 0
  const GTri* tris = nullptr;
  bool in_obj = false;
  int cur_inst = -1;
  bool found = false;
  hit.t = t; hit.inst = -1; hit.prim = -1; hit.u = 0.f; hit.v = 0.f;

  // BVHAggregate pops node 0 and tests its bounds first (BVH.fs:45-47)
  {
    Slab s = slab(f3(sc.tlas.bmin[0], sc.tlas.bmin[1], sc.tlas.bmin[2]), f3(sc.tlas.bmax[0], sc.tlas.bmax[1], sc.tlas.bmax[2]), o, inv);
    if (!slab_pass(s, t)) return false;
  }
  uint32_t cur = sc.tlas.root | kTlasBit;

  for (;;) {
    if ((cur & kTlasBit) && in_obj) {  // back from a BLAS: restore the world-space ray
      o = wo; d = wd; inv = winv;
      nodes = sc.nodes;
      in_obj = false;
    }
    if (!(cur & kLeafBit)) {
      // ---- interior node (TLAS or BLAS): both children's boxes in one 64-B record
      const float4* np = reinterpret_cast<const float4*>(nodes + (cur & kIndexMask));
      const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
      const Slab sl = slab(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), o, inv);
      const Slab sr = slab(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), o, inv);
      const bool pl = slab_pass(sl, t), pr = slab_pass(sr, t);
      const uint32_t level = cur & kTlasBit;
      const uint32_t left = fbits(n3.x) | level, right = fbits(n3.y) | level;
      const uint32_t axis = fbits(n3.z);
      const float dax = axis == 0 ? d.x : (axis == 1 ? d.y : d.z);
      const bool left_first = dax > 0.f;  // BVH.fs:51-56 / Mesh.fs:235-240
      const uint32_t nref = left_first ? left : right, fref = left_first ? right : left;
      const bool pn = left_first ? pl : pr, pf = left_first ? pr : pl;
      if (pn) {
        cur = nref;
        if (pf) { stk[sp] = fref; stkt[sp] = left_first ? sr.tmin : sl.tmin; ++sp; }
        continue;
      }
      if (pf) { cur = fref; continue; }
    } else if (cur & kTlasBit) {
      uint32_t count = (cur >> 24) & 63u;
      uint32_t first = cur & kFirstMask;
      if (count != 0) {
        // TLAS leaf: instances are visited in slot order (BVH.fs:49-50); defer all but the first
        for (uint32_t k = count - 1; k >= 1; --k) { stk[sp] = kLeafBit | kTlasBit | (first + k); stkt[sp] = -CUDART_INF_F; ++sp; }
      }
      // ---- PrimitiveInstance.Intersect (Primitive.fs:111-129) for slot `first`
      const float4* hp = reinterpret_cast<const float4*>(sc.inst_head + first);
      const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
      const Slab sb = slab(f3(h0.x, h0.y, h0.z), f3(h1.x, h1.y, h1.z), o, inv);
      if (slab_pass(sb, t)) {
        const Mat43 M = load_mat43(reinterpret_cast<const float4*>(sc.inst_w2o + first));
        const float3 oo = transform_point(wo, M);  // Ray.Transform (Ray.fs:19-22)
        const float3 od = transform_dir(wd, M);
        const uint32_t kind_prim = fbits(h0.w);
        if (kind_prim & 0x80000000u) {
          float tp;
          int root = sphere_test(__ldg(sc.sphere_radii + (kind_prim & 0x7FFFFFFFu)), oo, od, t, tp);
          if (root) {
            if (ANY) return true;
            t = tp; found = true;
            hit.inst = (int)first; hit.prim = root - 1; hit.u = 0.f; hit.v = 0.f;
          }
        } else {
          const GMesh* mesh = sc.meshes + kind_prim;
          const float4 m0 = __ldg(reinterpret_cast<const float4*>(mesh));
          const float4 m1 = __ldg(reinterpret_cast<const float4*>(mesh) + 1);
          const float4 m2 = __ldg(reinterpret_cast<const float4*>(mesh) + 2);
          o = oo; d = od; inv = rcp3(od);
          in_obj = true;
          cur_inst = (int)first;
          nodes = sc.nodes + fbits(m1.w);
          tris = sc.tris + fbits(m2.x);
          // MeshPrimitive pops BLAS node 0 and tests its bounds first (Mesh.fs:224-227)
          const Slab sm = slab(f3(m0.x, m0.y, m0.z), f3(m1.x, m1.y, m1.z), o, inv);
          if (slab_pass(sm, t)) { cur = fbits(m0.w); continue; }
        }
      }
    } else {
      // ---- BLAS leaf: triangles in slot order, each behind its own exact AABB
      // test (Mesh.fs:229-233 — load-bearing)
      const uint32_t count = (cur >> 24) & 63u;
      const uint32_t first = cur & kFirstMask;
      for (uint32_t k = 0; k < count; ++k) {
        const float4* tp4 = reinterpret_cast<const float4*>(tris + first + k);
        const float4 a = __ldg(tp4), b = __ldg(tp4 + 1), c = __ldg(tp4 + 2);
        const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(b.x, b.y, b.z), p2 = f3(c.x, c.y, c.z);
        const float3 lo = min_native(min_native(p0, p1), p2);  // Triangle.Bounds, Mesh.fs:19-22
        const float3 hi = max_native(max_native(p0, p1), p2);
        if (!slab_pass(slab(lo, hi, o, inv), t)) continue;
        float tp, u, v;
        if (tri_test(p0, p1, p2, o, d, t, tp, u, v)) {
          if (ANY) return true;
          t = tp; found = true;
          hit.inst = cur_inst; hit.prim = (int)(first + k); hit.u = u; hit.v = v;
        }
      }
    }
    // ---- pop; a deferred child is re-checked against the CURRENT t
    for (;;) {
      if (sp == 0) { hit.t = t; return found; }
      --sp;
      cur = stk[sp];
      if (ANY || stkt[sp] <= t) break;
    }
  }
}

}  // namespace bn
