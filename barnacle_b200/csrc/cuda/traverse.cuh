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
//
// Execution model: a persistent, warp-synchronous loop.  Every lane owns one ray;
// the warp alternates between three phases so that lanes doing the same kind of
// work do it together (SIMT efficiency is what bounds this kernel, not DRAM):
//   N  node steps (64-B GNode, two slab tests) until every lane holds a leaf
//   E  enter instance: ray -> object space, BLAS root / sphere test
//   T  triangles of the held BLAS leaf, one triangle per step across the lanes
// Lanes whose ray is finished are refilled from the global queue (one
// warp-aggregated atomic) as soon as enough of them are idle.
#pragma once
#include "device_scene.h"
#include "traverse_limits.h"
#include "vecmath.cuh"

namespace bn {

constexpr uint32_t kTlasBit = 0x40000000u;
constexpr uint32_t kIndexMask = 0x3FFFFFFFu;
constexpr uint32_t kFirstMask = 0x00FFFFFFu;
constexpr uint32_t kNone = 0xFFFFFFFFu;  // "stack empty": a leaf ref that no scene can produce
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kRefillMin = 8;            // refill when at least this many lanes are idle

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

// Per-lane traversal state of the persistent loop.
struct Lane {
  float3 wo, wd;          // world-space ray
  float3 o, d, inv;       // ray in the CURRENT space
  float t;                // closest distance so far (ANY: the fixed tmax)
  int h_inst, h_prim;
  float h_u, h_v;
  uint32_t cur;           // ref being processed (kNone: nothing left)
  int sp;
  int cur_inst;
  int index;              // queue slot of this ray
  const GNode* nodes;
  const GTri* tris;
  bool in_obj, fast, wfast, active;
};

// IO concept:  int count() ; int* cursor() ;
//              void load(int i, float3& o, float3& d, float& tmax) ;
//              void store(int i, bool hit, float t, int inst, int prim, float u, float v)
template <bool ANY, class IO>
BN_DEV void traverse_persistent(const DScene& sc, IO& io) {
  uint2 stk[kStackSize];  // (ref, entry distance bits)
  Lane L;
  L.active = false;
  L.cur = kNone;
  L.sp = 0;
  const int n = io.count();
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  bool exhausted = false;
  const bool scene_fast = sc.all_finite != 0u;

  // pops the next entry whose stored entry distance still passes; kNone if empty
  auto pop = [&]() {
    for (;;) {
      if (L.sp == 0) { L.cur = kNone; return; }
      --L.sp;
      const uint2 e = stk[L.sp];
      if (ANY || __uint_as_float(e.y) <= L.t) { L.cur = e.x; break; }
    }
    if ((L.cur & kTlasBit) && L.in_obj) {  // back from a BLAS: restore the world-space ray
      L.o = L.wo; L.d = L.wd; L.inv = rcp3(L.wd);
      L.nodes = sc.nodes;
      L.in_obj = false;
      L.fast = L.wfast;
    }
  };
  auto finish = [&](bool hit) {
    io.store(L.index, hit, L.t, L.h_inst, L.h_prim, L.h_u, L.h_v);
    L.active = false;
    L.cur = kNone;
    L.sp = 0;
  };

  for (;;) {
    // ---- refill idle lanes from the queue
    const unsigned idle = __ballot_sync(kFull, !L.active);
    if (idle != 0u && !exhausted && (__popc(idle) >= kRefillMin || idle == kFull)) {
      const int cnt = __popc(idle);
      int base = 0;
      if (lane == 0) base = atomicAdd(io.cursor(), cnt);
      base = __shfl_sync(kFull, base, 0);
      if (base + cnt >= n) exhausted = true;
      const int mine = base + __popc(idle & lt_mask);
      if (!L.active && mine < n) {
        L.index = mine;
        io.load(mine, L.wo, L.wd, L.t);
        L.o = L.wo; L.d = L.wd; L.inv = rcp3(L.wd);
        L.wfast = scene_fast && slab_fast_ok(L.wo, L.inv);
        L.fast = L.wfast;
        L.nodes = sc.nodes; L.tris = nullptr;
        L.in_obj = false; L.cur_inst = -1; L.sp = 0;
        L.h_inst = -1; L.h_prim = -1; L.h_u = 0.f; L.h_v = 0.f;
        L.active = true;
        // BVHAggregate pops node 0 and tests its bounds first (BVH.fs:45-47)
        const float3 bmin = f3(sc.tlas.bmin[0], sc.tlas.bmin[1], sc.tlas.bmin[2]), bmax = f3(sc.tlas.bmax[0], sc.tlas.bmax[1], sc.tlas.bmax[2]);
        const bool pass = L.fast ? slab_pass<true>(slab<true>(bmin, bmax, L.o, L.inv), L.t) : slab_pass<false>(slab<false>(bmin, bmax, L.o, L.inv), L.t);
        if (pass) L.cur = sc.tlas.root | kTlasBit;
        else finish(false);
      }
    }
    if (!__any_sync(kFull, L.active)) {
      if (exhausted) break;
      continue;
    }

    // ---- phase N: node steps until every active lane holds a leaf (or is done)
    for (;;) {
      const bool want = L.active && !(L.cur & kLeafBit);
      if (!__any_sync(kFull, want)) break;
      if (want) {
        const float4* np = reinterpret_cast<const float4*>(L.nodes + (L.cur & kIndexMask));
        const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
        Slab sl, sr;
        bool pl, pr;
        if (L.fast) {
          sl = slab<true>(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), L.o, L.inv);
          sr = slab<true>(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), L.o, L.inv);
          pl = slab_pass<true>(sl, L.t); pr = slab_pass<true>(sr, L.t);
        } else {
          sl = slab<false>(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), L.o, L.inv);
          sr = slab<false>(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), L.o, L.inv);
          pl = slab_pass<false>(sl, L.t); pr = slab_pass<false>(sr, L.t);
        }
        const uint32_t level = L.cur & kTlasBit;
        const uint32_t left = fbits(n3.x) | level, right = fbits(n3.y) | level;
        const uint32_t axis = fbits(n3.z);
        const float dax = axis == 0 ? L.d.x : (axis == 1 ? L.d.y : (axis == 2 ? L.d.z : 1.f));
        const bool left_first = dax > 0.f;  // BVH.fs:51-56 / Mesh.fs:235-240
        const uint32_t nref = left_first ? left : right, fref = left_first ? right : left;
        const bool pn = left_first ? pl : pr, pf = left_first ? pr : pl;
        if (pn) {
          L.cur = nref;
          if (pf) { stk[L.sp] = make_uint2(fref, __float_as_uint(left_first ? sr.tmin : sl.tmin)); ++L.sp; }
        } else if (pf) {
          L.cur = fref;
        } else {
          pop();
        }
        if (L.cur == kNone) finish(L.h_inst >= 0);
      }
    }

    // ---- phase E: PrimitiveInstance.Intersect (Primitive.fs:111-129); the world AABB
    // test already happened in the parent node
    if (L.active && (L.cur & (kLeafBit | kTlasBit)) == (kLeafBit | kTlasBit)) {
      const uint32_t slot = L.cur & kIndexMask;
      const float4* ip = reinterpret_cast<const float4*>(sc.inst_trav + slot);
      const Mat43 M = load_mat43(ip);
      const float4 m0 = __ldg(ip + 3), m1 = __ldg(ip + 4), m2 = __ldg(ip + 5);
      const float3 oo = transform_point(L.wo, M);  // Ray.Transform (Ray.fs:19-22)
      const float3 od = transform_dir(L.wd, M);
      bool descended = false;
      if (fbits(m2.y)) {
        float tp;
        const int root = sphere_test(m2.z, oo, od, L.t, tp);
        if (root) {
          L.h_inst = (int)slot; L.h_prim = root - 1; L.h_u = 0.f; L.h_v = 0.f;
          if (ANY) { finish(true); descended = true; }  // lane is done: nothing to pop
          else L.t = tp;
        }
      } else {
        L.o = oo; L.d = od; L.inv = rcp3(od);
        L.in_obj = true;
        L.fast = scene_fast && slab_fast_ok(oo, L.inv);
        L.cur_inst = (int)slot;
        L.nodes = sc.nodes + fbits(m1.w);
        L.tris = sc.tris + fbits(m2.x);
        // MeshPrimitive pops BLAS node 0 and tests its bounds first (Mesh.fs:224-227)
        const float3 bmin = f3(m0.x, m0.y, m0.z), bmax = f3(m1.x, m1.y, m1.z);
        const bool pass = L.fast ? slab_pass<true>(slab<true>(bmin, bmax, L.o, L.inv), L.t) : slab_pass<false>(slab<false>(bmin, bmax, L.o, L.inv), L.t);
        if (pass) { L.cur = fbits(m0.w); descended = true; }
      }
      if (!descended) {
        pop();
        if (L.cur == kNone) finish(L.h_inst >= 0);
      }
    }

    // ---- phase T: triangles of the held BLAS leaf in slot order, each behind its own
    // exact AABB test (Mesh.fs:229-233 — load-bearing)
    {
      bool has = L.active && (L.cur & (kLeafBit | kTlasBit)) == kLeafBit;
      const uint32_t count = has ? ((L.cur >> 24) & 63u) : 0u;
      const uint32_t first = L.cur & kFirstMask;
      for (uint32_t k = 0;; ++k) {
        const bool more = has && k < count;
        if (!__any_sync(kFull, more)) break;
        if (more) {
          const float4* tp4 = reinterpret_cast<const float4*>(L.tris + first + k);
          const float4 a = __ldg(tp4), b = __ldg(tp4 + 1), c = __ldg(tp4 + 2);
          const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(b.x, b.y, b.z), p2 = f3(c.x, c.y, c.z);
          bool box;
          if (L.fast) {  // Triangle.Bounds (Mesh.fs:19-22): finite vertices => fmin/fmax == minps/maxps
            const float3 lo = f3(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z));
            const float3 hi = f3(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z));
            box = slab_pass<true>(slab<true>(lo, hi, L.o, L.inv), L.t);
          } else {
            const float3 lo = min_native(min_native(p0, p1), p2);
            const float3 hi = max_native(max_native(p0, p1), p2);
            box = slab_pass<false>(slab<false>(lo, hi, L.o, L.inv), L.t);
          }
          float tp, u, v;
          if (box && tri_test(p0, p1, p2, L.o, L.d, L.t, tp, u, v)) {
            L.h_inst = L.cur_inst; L.h_prim = (int)(first + k); L.h_u = u; L.h_v = v;
            if (ANY) { finish(true); has = false; }
            else L.t = tp;
          }
        }
      }
      if (has) {
        pop();
        if (L.cur == kNone) finish(L.h_inst >= 0);
      }
    }
  }
}

}  // namespace bn
