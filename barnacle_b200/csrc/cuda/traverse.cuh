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
// Execution model: a persistent, warp-synchronous loop.  Every lane owns one ray
// and is in one of four states; each iteration the warp votes and runs the ONE
// phase most lanes are ready for, so that lanes doing the same kind of work do it
// together (SIMT efficiency is what bounds this kernel, not DRAM):
//   N  one node step (64-B GNode, two slab tests)
//   E  enter instance: ray -> object space, BLAS root / sphere test
//   T  the next triangle of the held BLAS leaf
//   S  small TLAS only: the next candidate of the octant-ordered instance list
// Phases N and T, once chosen, repeat on a single ballot while enough lanes still
// have that kind of work (kStayMin / kStayT).  Lanes whose ray is finished are
// refilled from the global queue (one warp-aggregated atomic) as soon as enough of
// them are idle (kRefillMin / kRefillMinAny).
//
// The phases run the FAST slab form (vecmath.cuh), which is bit-identical to the
// reference's for every ray whose 1/d has only finite non-zero components.  A ray
// that is not (zero / denormal direction component, in world or in some object
// space) is appended to a deferred list and re-traced from scratch by the
// companion fix-up kernel with trace_exact(), a plain per-lane loop that spells
// out the minps/maxps + NaN-propagating Math.Max/Min sequence.
#pragma once
#include <type_traits>

#include "device_scene.h"
#include "traverse_limits.h"
#include "vecmath.cuh"

namespace bn {

constexpr uint32_t kTlasBit = 0x40000000u;
constexpr uint32_t kIndexMask = 0x3FFFFFFFu;
constexpr uint32_t kFirstMask = 0x07FFFFFFu;
constexpr uint32_t kNone = 0xFFFFFFFFu;  // lane idle (a leaf ref that no scene can produce)
constexpr uint32_t kScan = 0xFFFFFFFEu;  // lane is at the small-TLAS ordered scan (ditto)
constexpr unsigned kFull = 0xFFFFFFFFu;
#ifndef BN_REFILL_MIN
#define BN_REFILL_MIN 20   // (14 before the paths were ordered between bounces; re-swept on ordered rays: profiles/r02_ab_session18_*.log)
#endif
#ifndef BN_REFILL_MIN_ANY
#define BN_REFILL_MIN_ANY 20
#endif
constexpr int kRefillMin = BN_REFILL_MIN;  // refill when at least this many lanes are idle
constexpr int kRefillMinAny = BN_REFILL_MIN_ANY;  // ... for any-hit (shadow) rays
#ifndef BN_STAY_MIN
#define BN_STAY_MIN 6
#endif
#ifndef BN_STAY_T
#define BN_STAY_T 4
#endif
#ifdef BN_EXP_ANY_UNORDERED   // (round-1 name of the switch)
#define BN_ANY_UNORDERED 1
#endif
#ifndef BN_ANY_UNORDERED
#define BN_ANY_UNORDERED 1
#endif
#ifndef BN_WIDE_LDG256
#define BN_WIDE_LDG256 0   // 1: a 4-wide node is fetched with four 256-bit loads instead of eight 128-bit ones
#endif
#ifndef BN_SPLIT_REFILL
#define BN_SPLIT_REFILL 0   // 1: split-phase refill — the cursor's atomic is issued at one vote and its result used at the next, with ONE
                            // phase step of the lanes that still hold a ray in between (the claim stays exact: the lanes idle at the
                            // first vote get the slots)
#endif
#ifndef BN_POP2
#define BN_POP2 0   // 1: closest-hit pops look at the two topmost entries at once (their loads in flight together): a stale entry — its
                    // entry distance beyond the t found since it was pushed — then costs no second local-memory round trip
#endif
#ifndef BN_STACK_TOP_REG
#define BN_STACK_TOP_REG 0   // 1: the top entry of the traversal stack lives in registers; a pop hands it out and starts the load of
                            // the entry below without waiting for it (the pop's local-memory load was 8-11 % of the ordered
                            // kernel's stall samples, executed by 2-3 lanes at a time)
#endif
#ifndef BN_PREFETCH_AHEAD
#define BN_PREFETCH_AHEAD 16384
#endif
constexpr int kPrefetchAhead = BN_PREFETCH_AHEAD;  // queue entries between a refill's loads and its L2 prefetches
constexpr int kStayT = BN_STAY_T;        // same for phase T (33: one triangle per vote)
constexpr int kStayMin = BN_STAY_MIN;      // phase N repeats without a vote while at least this many lanes are at a node (33: never)

#ifdef BN_TRAV_STATS
// debug builds only (python -m barnacle_b200.build --stats): phase executions and ready lanes
__device__ unsigned long long g_trav_stats[24];
#define BN_STAT(slot, lanes) do { const int n_ready_ = (lanes); if ((threadIdx.x & 31) == 0) { st_cnt[slot] += 1; st_sum[slot] += n_ready_; } } while (0)
#else
#define BN_STAT(slot, lanes) do { } while (0)
#endif
// Host warp emulator only (tests/hostsim, tools/warp_stats.py): per-lane work in SASS-instruction units; the emulator
// charges a warp the MAXIMUM over its lanes between two rendezvous, i.e. what a SIMT machine pays for a divergent step.
#if defined(BN_TRAV_STATS) && defined(BN_HOSTSIM_WARP)
#define BN_WORK(units) ::hostsim_work(units)  /* declared by tests/hostsim/device_shim.h */
#else
#define BN_WORK(units) do { } while (0)
#endif

BN_DEV uint32_t fbits(float f) { return __float_as_uint(f); }

// 32-B load (sm_100: LDG.E.ENL2.256): one L1 sector per lane in ONE request
struct F8 { float v[8]; };
#ifndef BN_HOSTSIM
BN_DEV F8 ldg256(uintptr_t p) {
  F8 r;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
#else
BN_DEV F8 ldg256(uintptr_t p) { F8 r; for (int k = 0; k < 8; ++k) r.v[k] = reinterpret_cast<const float*>(p)[k]; return r; }
#endif

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

// bit a set iff d[a] > 0; bit 3 always set (GNode.axis == 3: "always left first")
BN_DEV uint32_t dir_signs(float3 d) { return (d.x > 0.f ? 1u : 0u) | (d.y > 0.f ? 2u : 0u) | (d.z > 0.f ? 4u : 0u) | 8u; }

// slab<true> + slab_pass<true> on a GFlatInst, whose planes are already sorted into near (a) / far (b) for the ray's octant
BN_DEV bool flat_pass(const float4 a, const float4 b, const float3 o, const float3 inv, const float t) {
  const float tn = fmaxf(fmaxf(1e-3f, (a.x - o.x) * inv.x), fmaxf((a.y - o.y) * inv.y, (a.z - o.z) * inv.z));
  const float tf = fminf((b.x - o.x) * inv.x, fminf((b.y - o.y) * inv.y, (b.z - o.z) * inv.z));
  return tn <= fminf(t, tf);
}

struct TraceResult {
  bool hit;
  float t;
  int inst, prim;  // prim: ABSOLUTE triangle index; spheres: 0 = near root (t0), 1 = far root (t1)
  float u, v;
};

// ---------------------------------------------------------------------------------
// Exact per-lane traversal (slow path, fix-up kernel).  Same visiting order, the
// reference's slab operations spelled out (slab<false>).
// ---------------------------------------------------------------------------------
// FAST = true runs the same loop with the fast slab form (see slab<> in vecmath.cuh) and returns false
// — result unusable — as soon as the ray turns out not to qualify in some object space; the caller
// (trace_lane) then repeats it with FAST = false.  FAST = false always returns true.
template <bool ANY, bool FAST>
BN_DEV bool trace_lane_impl(const DScene& sc, const float3 wo, const float3 wd, float t, TraceResult& res) {
  uint2 stk[kStackSize];
  int sp = 0;
  const float3 winv = rcp3(wd);
  float3 o = wo, d = wd, inv = winv;
  uint32_t signs = dir_signs(wd);
  bool in_obj = false;
  int cur_inst = -1;
  res.hit = false; res.t = t; res.inst = -1; res.prim = -1; res.u = 0.f; res.v = 0.f;
  if (!slab_pass<FAST>(slab<FAST>(f3(sc.tlas.bmin[0], sc.tlas.bmin[1], sc.tlas.bmin[2]), f3(sc.tlas.bmax[0], sc.tlas.bmax[1], sc.tlas.bmax[2]), o, inv), t)) return true;
  uint32_t cur = sc.tlas.root | kTlasBit;
  for (;;) {
    if ((cur & kTlasBit) && in_obj) { o = wo; d = wd; inv = winv; signs = dir_signs(wd); in_obj = false; }
    if (!(cur & kLeafBit)) {
      const float4* np = reinterpret_cast<const float4*>(sc.nodes + (cur & kIndexMask));
      const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
      const Slab sl = slab<FAST>(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), o, inv);
      const Slab sr = slab<FAST>(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), o, inv);
      const bool pl = slab_pass<FAST>(sl, t), pr = slab_pass<FAST>(sr, t);
      const uint32_t level = cur & kTlasBit;
      const uint32_t left = fbits(n3.x) | level, right = fbits(n3.y) | level;
      const bool left_first = ((signs >> fbits(n3.z)) & 1u) != 0u;
      const uint32_t nref = left_first ? left : right, fref = left_first ? right : left;
      const bool pn = left_first ? pl : pr, pf = left_first ? pr : pl;
      if (pn) {
        cur = nref;
        if (pf) { stk[sp] = make_uint2(fref, __float_as_uint(left_first ? sr.tmin : sl.tmin)); ++sp; }
        continue;
      }
      if (pf) { cur = fref; continue; }
    } else if (cur & kTlasBit) {
      const uint32_t slot = cur & kIndexMask;
      const float4* ip = reinterpret_cast<const float4*>(sc.inst_trav + slot);
      const Mat43 M = load_mat43(ip);
      const float4 m0 = __ldg(ip + 3), m1 = __ldg(ip + 4), m2 = __ldg(ip + 5);
      const float3 oo = transform_point(wo, M), od = transform_dir(wd, M);
      if (fbits(m2.y)) {
        float tp;
        const int root = sphere_test(m2.z, oo, od, t, tp);
        if (root) {
          res.hit = true; res.inst = (int)slot; res.prim = root - 1; res.u = 0.f; res.v = 0.f;
          if (ANY) return true;
          t = tp; res.t = tp;
        }
      } else {
        o = oo; d = od; inv = rcp3(od); signs = dir_signs(od);
        if (FAST && !slab_fast_ok(oo, inv)) return false;
        in_obj = true;
        cur_inst = (int)slot;
        if (slab_pass<FAST>(slab<FAST>(f3(m0.x, m0.y, m0.z), f3(m1.x, m1.y, m1.z), o, inv), t)) { cur = fbits(m0.w); continue; }
      }
    } else {
      const uint32_t count = (cur >> 27) & 7u, first = cur & kFirstMask;
      for (uint32_t k = 0; k < count; ++k) {
        const float4* tp4 = reinterpret_cast<const float4*>(sc.tris + first + k);
        const float4 a = __ldg(tp4), b = __ldg(tp4 + 1), c = __ldg(tp4 + 2);
        const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(b.x, b.y, b.z), p2 = f3(c.x, c.y, c.z);
        // Triangle.Bounds (Mesh.fs:19-22): finite vertices => fmin/fmax == minps/maxps
        const float3 lo = FAST ? f3(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z)) : min_native(min_native(p0, p1), p2);
        const float3 hi = FAST ? f3(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z)) : max_native(max_native(p0, p1), p2);
        if (!slab_pass<FAST>(slab<FAST>(lo, hi, o, inv), t)) continue;
        float tp, u, v;
        if (tri_test(p0, p1, p2, o, d, t, tp, u, v)) {
          res.hit = true; res.inst = cur_inst; res.prim = (int)(first + k); res.u = u; res.v = v;
          if (ANY) return true;
          t = tp; res.t = tp;
        }
      }
    }
    for (;;) {
      if (sp == 0) return true;
      --sp;
      cur = stk[sp].x;
      if (ANY || __uint_as_float(stk[sp].y) <= t) break;
    }
  }
}

// The reference's operations one for one (fix-up kernel, bn_trace's exact mode).
template <bool ANY>
BN_DEV void trace_exact(const DScene& sc, const float3 wo, const float3 wd, float t, TraceResult& res) {
  trace_lane_impl<ANY, false>(sc, wo, wd, t, res);
}
// Per-lane trace for callers that are not warp-synchronous (the PSSMLT megakernels).  With
// BN_MLT_FAST_TRACE: the fast slab form whenever the ray qualifies (bit-identical — the PSSMLT parity
// tests pass with it — same argument as for the persistent loop), else, or when some object space
// disqualifies it half way, the exact form from scratch.  Measured on C5 and NOT the default: the
// bootstrap kernel gains 9 % (35.9 -> 32.8 ms) but the chain kernel, which sits at its 128-register
// cap with both forms inlined twice, loses 13 % (134 -> 152 ms).
template <bool ANY>
BN_DEV void trace_lane(const DScene& sc, const float3 wo, const float3 wd, float t, TraceResult& res) {
#ifdef BN_MLT_FAST_TRACE
  if (sc.all_finite != 0u && slab_fast_ok(wo, rcp3(wd)) && trace_lane_impl<ANY, true>(sc, wo, wd, t, res)) return;
#endif
  trace_lane_impl<ANY, false>(sc, wo, wd, t, res);
}

// ---------------------------------------------------------------------------------
// Persistent warp-synchronous traversal (fast path).
// IO concept:  int count() ; int* cursor() ;
//              void load(int i, float3& o, float3& d, float& tmax) ;
//              void store(int i, const TraceResult&) ;
//              void defer(int i)      // ray i needs the exact path (fix-up kernel)
//              void prefetch(int i)   // hint: ray i will be loaded soon (L2 prefetch)
// optional:    static constexpr bool kHasCand = true ; uint32_t cand(int i)
//              the small-TLAS candidate word of ray i, computed beforehand by candidate_word() in a full-width
//              pre-pass (kernels.cu: k_candidates) instead of by the refilled lanes of this loop
// ---------------------------------------------------------------------------------
constexpr uint32_t kCandDefer = 0x80000000u;  // candidate word: the ray does not qualify for the fast slab form
template <class IO, class = void> struct io_has_cand : std::false_type {};
template <class IO> struct io_has_cand<IO, std::void_t<decltype(IO::kHasCand)>> : std::bool_constant<IO::kHasCand> {};

// Small TLAS: ONE uniform pass over the octant-ordered instance boxes with the initial t gives the candidate set
// (bit k = k-th instance in visiting order).  t only shrinks, so an instance that fails now fails later too; candidates
// are re-checked against the then-current t when their turn comes (phase S), which is the reference's test.
BN_DEV uint32_t candidate_mask(const DScene& sc, const float3 wo, const float3 winv, const uint32_t wsigns, const float t) {
  const uint32_t n_inst = sc.n_inst;
  const float4* fp = reinterpret_cast<const float4*>(sc.flat_tlas + (size_t)(wsigns & 7u) * n_inst);
  uint32_t mask = 0u;
  for (uint32_t k = 0; k < n_inst; ++k) {
    BN_WORK(25);
    const float4 a = __ldg(fp + 2u * k), b = __ldg(fp + 2u * k + 1u);
    if (flat_pass(a, b, wo, winv, t)) mask |= 1u << k;
  }
  return mask;
}
// The word the pre-pass stores per ray: kCandDefer, or the candidate mask (n_inst <= 16: bits 0..15).
BN_DEV uint32_t candidate_word(const DScene& sc, const float3 wo, const float3 wd, const float t) {
  const float3 winv = rcp3(wd);
  if (!(sc.all_finite != 0u && slab_fast_ok(wo, winv))) return kCandDefer;
  return candidate_mask(sc, wo, winv, dir_signs(wd), t);
}
// `cold`: this thread's column of a [kTravColdWords][blockDim.x] shared-memory array.  The state a
// ray touches a few times in its life (committed hit, queue slot, world direction, candidate mask)
// lives there instead of in registers; that is what lets the kernel fit 9 CTAs per SM.
constexpr int kTravColdWords = 9;
// Traversal stack entry.  Closest hit: (ref, entry distance), re-checked against the current t at pop.
// Any hit: t is fixed, the entry distance is never read again, so shadow rays keep refs only (half the
// local-memory traffic of their pushes and pops).
template <bool ANY> struct TravStack;
template <> struct TravStack<false> {
  using type = uint2;
  static BN_DEV void push(uint2& e, uint32_t ref, float tmin) { e = make_uint2(ref, __float_as_uint(tmin)); }
  static BN_DEV bool pop(const uint2& e, float t, uint32_t& cur) { if (__uint_as_float(e.y) <= t) { cur = e.x; return true; } return false; }
};
template <> struct TravStack<true> {
  using type = uint32_t;
  static BN_DEV void push(uint32_t& e, uint32_t ref, float) { e = ref; }
  static BN_DEV bool pop(const uint32_t& e, float, uint32_t& cur) { cur = e; return true; }
};
// WIDE = true: interior refs index the 4-wide nodes (sc.wide, device_scene.h: GWide) and phase N is one step of such a
// node — four slab tests, the passing children taken in the reference's visiting order; WIDE = false: the binary
// two-box nodes (sc.nodes).  Leaves, phases T / E / S, the stack discipline and the results are the same.
template <bool ANY, bool WIDE, class IO>
BN_DEV void traverse_persistent(const DScene& sc, const IO& io, uint32_t* __restrict__ cold, const int cold_stride) {
  typename TravStack<ANY>::type stk[kStackSize];
#if BN_STACK_TOP_REG
  typename TravStack<ANY>::type top{};  // entry sp - 1 (entries 0 .. sp - 2 are in stk[])
#endif
  // per-lane state
  float3 wo, winv;        // world-space ray (origin, 1/direction)
  float3 o, d, inv;       // ray in the CURRENT space
  float t = 0.f;          // closest distance so far (ANY: the fixed tmax)
  int& h_inst = reinterpret_cast<int*>(cold)[0];
  int& h_prim = reinterpret_cast<int*>(cold)[cold_stride];
  float& h_u = reinterpret_cast<float*>(cold)[2 * cold_stride];
  float& h_v = reinterpret_cast<float*>(cold)[3 * cold_stride];
  float& wdx = reinterpret_cast<float*>(cold)[4 * cold_stride];
  float& wdy = reinterpret_cast<float*>(cold)[5 * cold_stride];
  float& wdz = reinterpret_cast<float*>(cold)[6 * cold_stride];
  h_inst = -1; h_prim = -1; h_u = 0.f; h_v = 0.f;
  uint32_t cur = kNone;   // ref being processed; kScan: next instance of the flat TLAS; kNone: lane idle
  uint32_t signs = 8u;    // dir_signs of the current-space direction
  // (the signs of the world-space direction are read off winv when needed — same signs for every ray the fast phases see, whose
  // direction components are all non-zero — instead of living in a register)
  auto wsigns_of = [&]() { return dir_signs(winv); };
  int sp = 0;
  auto push_entry = [&](uint32_t ref, float tmin) {
#if BN_STACK_TOP_REG
    if (sp > 0) stk[sp - 1] = top;
    TravStack<ANY>::push(top, ref, tmin);
#else
    TravStack<ANY>::push(stk[sp], ref, tmin);
#endif
    ++sp;
  };
  int cur_inst = -1;
  int& index = reinterpret_cast<int*>(cold)[7 * cold_stride];  // queue slot of this ray
  uint32_t tri_k = 0;     // next triangle of the held BLAS leaf
  uint32_t& tl_pos = cold[8 * cold_stride];  // small TLAS: bit mask of the candidate instances still to visit (visiting order)
  bool in_obj = false;
  index = 0; tl_pos = 0u;
  wdx = wdy = wdz = 0.f;
  wo = winv = o = d = inv = splat(0.f);

#ifdef BN_TRAV_STATS
  unsigned long long st_cnt[5] = {0, 0, 0, 0, 0}, st_sum[5] = {0, 0, 0, 0, 0};
#endif
  const int n = io.count();
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  bool exhausted = false;
#if BN_SPLIT_REFILL
  unsigned pend_mask = 0u;  // lanes that get the slots of the claim in flight (0: none)
  int pend_base = 0;        // lane 0: the atomic's result
#endif
  const int refill_min = ANY ? kRefillMinAny : (sc.refill_min != 0u ? (int)sc.refill_min : kRefillMin);  // (per scene: DScene.refill_min)
  const bool scene_fast = sc.all_finite != 0u;
  const bool flat = sc.flat_tlas != nullptr;
  const uint32_t n_inst = sc.n_inst;
  const float4* const node_base = reinterpret_cast<const float4*>(sc.nodes);
  const float4* const tri_base = reinterpret_cast<const float4*>(sc.tris);

  auto finish = [&]() {
    TraceResult r;
    r.hit = h_inst >= 0; r.t = t; r.inst = h_inst; r.prim = h_prim; r.u = h_u; r.v = h_v;
    io.store(index, r);
    cur = kNone;
  };
  // pops the next entry whose stored entry distance still passes
  auto pop = [&]() {
    for (;;) {
      if (sp == 0) {
        if (flat) cur = kScan;  // BLAS exhausted: back to the ordered instance scan (world ray is kept in wo/winv)
        else finish();
        return;
      }
#if BN_POP2
      if (!ANY && sp >= 2) {
        const typename TravStack<ANY>::type e1 = stk[sp - 1], e0 = stk[sp - 2];
        if (TravStack<ANY>::pop(e1, t, cur)) { sp -= 1; break; }
        sp -= 2;
        if (TravStack<ANY>::pop(e0, t, cur)) break;
        continue;
      }
#endif
      --sp;
#if BN_STACK_TOP_REG
      const typename TravStack<ANY>::type e = top;
      if (sp > 0) top = stk[sp - 1];  // issued now, needed at the next pop
      if (TravStack<ANY>::pop(e, t, cur)) break;
#else
      if (TravStack<ANY>::pop(stk[sp], t, cur)) break;
#endif
    }
    if ((cur & kTlasBit) && in_obj) {  // tree TLAS: back from a BLAS, restore the world-space ray (d is only read inside a BLAS)
      o = wo; inv = winv;
      signs = wsigns_of();
      in_obj = false;
    }
  };

  for (;;) {
    // ---- one vote: how many lanes are ready for each phase (a single REDUX)
    const bool isN = !(cur & kLeafBit);
    const bool isT = (cur >> 30) == 2u;
    const bool isS = cur == kScan;
    const bool isE = (cur >> 30) == 3u && cur < kScan;
    const uint32_t votes = __reduce_add_sync(kFull, isN ? 1u : (isT ? (1u << 8) : (isE ? (1u << 16) : (isS ? (1u << 24) : 0u))));
    const int nN = (int)(votes & 255u), nT = (int)((votes >> 8) & 255u), nE = (int)((votes >> 16) & 255u), nS = (int)(votes >> 24);
    const int n_idle = 32 - (nN + nT + nE + nS);

    // ---- refill idle lanes from the queue
    // Each refill claims exactly the slots it fills, at the cursor's position of that moment.  (Measured and dropped,
    // profiles/r02_ab_session22_*.log, r02_ab_session23_*.log: holding claimed blocks / chunks of 64-256 consecutive slots per
    // warp with the next atomic in flight, so that no refill waits for the cursor's round trip — ncu's top long-scoreboard
    // stall, 10 % of the samples.  C1 +3.5 %, but C2 -7 %, C3 -5 %, C4 -7 %, and worse the more a warp claims ahead: with
    // exact claims all warps of the GPU walk ONE narrow front through the ordered queue, i.e. through the same cell of the
    // scene, and that is what keeps the node records in L1.)
#if BN_SPLIT_REFILL
    bool claimed_now = false;
    if (pend_mask == 0u && n_idle != 0 && !exhausted && (n_idle >= refill_min || n_idle == 32)) {
      pend_mask = __ballot_sync(kFull, cur == kNone);
      if (lane == 0) pend_base = atomicAdd(io.cursor(), n_idle);
      claimed_now = n_idle != 32;  // lanes with a ray left: one phase step while the atomic is in flight
    }
    if (pend_mask != 0u && !claimed_now) {
      const unsigned idle = pend_mask;
      const int claimed = __popc(idle);
      pend_mask = 0u;
      const int base = __shfl_sync(kFull, pend_base, 0);
      if (base + claimed >= n) exhausted = true;
      const int mine = base + __popc(idle & lt_mask);
      if (((idle >> lane) & 1u) && mine < n) {
#else
    if (n_idle != 0 && !exhausted && (n_idle >= refill_min || n_idle == 32)) {
      const unsigned idle = __ballot_sync(kFull, cur == kNone);
      int base = 0;
      if (lane == 0) base = atomicAdd(io.cursor(), n_idle);
      base = __shfl_sync(kFull, base, 0);
      if (base + n_idle >= n) exhausted = true;
      const int mine = base + __popc(idle & lt_mask);
      if (cur == kNone && mine < n) {
#endif
        index = mine;
        // Claims tile the queue in increasing order, so "my slot + kPrefetchAhead" is a ray some
        // warp will claim about a DRAM latency from now: every refill pulls its share of that
        // future window into L2.
        if (mine + kPrefetchAhead < n) io.prefetch(mine + kPrefetchAhead);
        float3 wd;
        BN_WORK(60);
        io.load(mine, wo, wd, t);
        wdx = wd.x; wdy = wd.y; wdz = wd.z;
        winv = rcp3(wd);
        const uint32_t wsigns = dir_signs(wd);
        o = wo; d = wd; inv = winv; signs = wsigns;
        in_obj = false; cur_inst = -1; sp = 0; tri_k = 0; tl_pos = 0;
        h_inst = -1; h_prim = -1; h_u = 0.f; h_v = 0.f;
        // what the refilled lane does when nobody has looked at its ray yet (flat_scan: the scene has a small TLAS)
        auto start_ray = [&](const bool flat_scan) {
          if (!(scene_fast && slab_fast_ok(wo, winv))) {
            io.defer(index);
          } else if (flat_scan) {
            const uint32_t mask = candidate_mask(sc, wo, winv, wsigns, t);
            tl_pos = mask;
            if (mask) cur = kScan;
            else finish();
          } else {
            // BVHAggregate pops node 0 and tests its bounds first (BVH.fs:45-47)
            const Slab s = slab<true>(f3(sc.tlas.bmin[0], sc.tlas.bmin[1], sc.tlas.bmin[2]), f3(sc.tlas.bmax[0], sc.tlas.bmax[1], sc.tlas.bmax[2]), o, inv);
            if (slab_pass<true>(s, t)) cur = (WIDE ? sc.tlas_wroot : sc.tlas.root) | kTlasBit;
            else finish();
          }
        };
        if constexpr (io_has_cand<IO>::value) {
          if (flat) {
            // the pre-pass has run the fast-form check and the candidate pass for this ray (k_candidates: all 32 lanes
            // of its warps busy, where the refilled lanes of this loop are 14-20 of 32)
            const uint32_t m = io.cand(mine);
            if (m & kCandDefer) {
              io.defer(index);
            } else {
              tl_pos = m;
              if (m) cur = kScan;
              else finish();
            }
          } else {
            start_ray(false);
          }
        } else {
          start_ray(flat);
        }
      }
      continue;  // re-vote with the new rays
    }
#if BN_SPLIT_REFILL
    if (n_idle == 32 && pend_mask == 0u) break;  // nothing in flight, nothing claimed, and the queue is exhausted
    if (n_idle == 32) continue;
#else
    if (n_idle == 32) break;  // nothing in flight and the queue is exhausted
#endif
    BN_STAT(3, 32 - n_idle);

    if (nN >= nT && nN >= nE && nN >= nS) {
      // ---- phase N: one node step (64-B GNode, two slab tests)
      // keep stepping while enough lanes still have a node step to do (one ballot instead of a vote).
      // "This lane is at a node" is read off `cur` (sign bit = kLeafBit) each time: a loop-carried flag
      // costs five instructions per step to keep in a register.
      for (;;) {
        BN_STAT(0, __popc(__ballot_sync(kFull, (int)cur >= 0)));
        if (WIDE) {
          if ((int)cur >= 0) {
            BN_WORK(135);
            // ---- one 4-wide node (128 B): near / far planes of the four slots picked by the direction signs at the
            // address level, so each box is 6 FADD + 6 FMUL and the two min / max chains of slab<true>
            const uintptr_t nb = reinterpret_cast<uintptr_t>(sc.wide) + (size_t)(cur & kIndexMask) * 128u;
#if BN_WIDE_LDG256
            // the whole node in four 32-B loads (one L1 sector each; sm_100's 256-bit LDG), near / far picked by selects
            // (swizzled chunks: a 32-B pair {2k, 2k+1} stays a pair under the XOR, with its halves exchanged when bit 0 of the swizzle is set)
            const uint32_t swz32 = BN_WIDE_SWIZZLE ? (cur & 6u) << 4 : 0u;
            const bool swap_halves = BN_WIDE_SWIZZLE && (cur & 1u) != 0u;
            F8 c0 = ldg256(nb | swz32), c1 = ldg256(nb | (swz32 ^ 32u)), c2 = ldg256(nb | (swz32 ^ 64u)), c3 = ldg256(nb | (swz32 ^ 96u));
            if (swap_halves) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float tq;
                tq = c0.v[q]; c0.v[q] = c0.v[q + 4]; c0.v[q + 4] = tq;
                tq = c1.v[q]; c1.v[q] = c1.v[q + 4]; c1.v[q + 4] = tq;
                tq = c2.v[q]; c2.v[q] = c2.v[q + 4]; c2.v[q + 4] = tq;
                tq = c3.v[q]; c3.v[q] = c3.v[q + 4]; c3.v[q + 4] = tq;
              }
            }
            const bool sx = (signs & 1u) != 0u, sy = (signs & 2u) != 0u, sz = (signs & 4u) != 0u;
            const float4 lx = make_float4(c0.v[0], c0.v[1], c0.v[2], c0.v[3]), ly = make_float4(c0.v[4], c0.v[5], c0.v[6], c0.v[7]);
            const float4 lz = make_float4(c1.v[0], c1.v[1], c1.v[2], c1.v[3]);
            const float4 hx = make_float4(c2.v[0], c2.v[1], c2.v[2], c2.v[3]), hy = make_float4(c2.v[4], c2.v[5], c2.v[6], c2.v[7]);
            const float4 hz = make_float4(c3.v[0], c3.v[1], c3.v[2], c3.v[3]);
            const float4 nx = sx ? lx : hx, fx = sx ? hx : lx, ny = sy ? ly : hy, fy = sy ? hy : ly, nz = sz ? lz : hz, fz = sz ? hz : lz;
            const uint4 rf = make_uint4(__float_as_uint(c1.v[4]), __float_as_uint(c1.v[5]), __float_as_uint(c1.v[6]), __float_as_uint(c1.v[7]));
            const uint32_t flips_word = __float_as_uint(c3.v[4]);
#else
            // near plane of an axis = lo if dir > 0, else hi (64 B further: bit 6 of the offset); far = the other one.  The
            // node's 16-B chunks are XOR-swizzled by its index (device_scene.h), so that the lanes of a warp, which all
            // read the same logical chunk of different nodes, hit different L1 bank groups: offset = logical ^ swz
            const uint32_t swz = BN_WIDE_SWIZZLE ? (cur & 7u) << 4 : 0u;
            const uint32_t mx = swz ^ ((signs & 1u) ? 0u : 64u), my = swz ^ ((signs & 2u) ? 16u : 80u), mz = swz ^ ((signs & 4u) ? 32u : 96u);
            const float4 nx = __ldg(reinterpret_cast<const float4*>(nb | mx)), fx = __ldg(reinterpret_cast<const float4*>(nb | (mx ^ 64u)));
            const float4 ny = __ldg(reinterpret_cast<const float4*>(nb | my)), fy = __ldg(reinterpret_cast<const float4*>(nb | (my ^ 64u)));
            const float4 nz = __ldg(reinterpret_cast<const float4*>(nb | mz)), fz = __ldg(reinterpret_cast<const float4*>(nb | (mz ^ 64u)));
            const uint4 rf = __ldg(reinterpret_cast<const uint4*>(nb | (swz ^ 48u)));
#endif
            // slab<true> + slab_pass<true> per slot; key = entry distance if the slot passes, else -1 (an entry distance is >= 1e-3)
#define BN_WIDE_SLOT(c)                                                                                                         \
            fmaxf(fmaxf(1e-3f, (nx.c - o.x) * inv.x), fmaxf((ny.c - o.y) * inv.y, (nz.c - o.z) * inv.z)) <=                      \
                    fminf(t, fminf((fx.c - o.x) * inv.x, fminf((fy.c - o.y) * inv.y, (fz.c - o.z) * inv.z)))                     \
                ? fmaxf(fmaxf(1e-3f, (nx.c - o.x) * inv.x), fmaxf((ny.c - o.y) * inv.y, (nz.c - o.z) * inv.z))                   \
                : -1.f
            float k0 = BN_WIDE_SLOT(x), k1 = BN_WIDE_SLOT(y), k2 = BN_WIDE_SLOT(z), k3 = BN_WIDE_SLOT(w);
#undef BN_WIDE_SLOT
            uint32_t r0 = rf.x, r1 = rf.y, r2 = rf.z, r3 = rf.w;
            // the reference's order: group L (slots 0, 1) before group R (2, 3) iff dir[axis_P] > 0, slot 0 before 1 iff
            // dir[axis_L] > 0, slot 2 before 3 iff dir[axis_R] > 0 (BVH.fs:51-56 / Mesh.fs:235-240 applied twice) — tabulated
            // per direction octant in the node.  An any-hit query returns the same boolean whatever the order.
            if (!ANY || !BN_ANY_UNORDERED) {
#if BN_WIDE_LDG256
              const uint32_t fl = flips_word >> ((signs & 7u) * 3u);
#else
              const uint32_t fl = __ldg(reinterpret_cast<const uint32_t*>(nb | (swz ^ 112u))) >> ((signs & 7u) * 3u);
#endif
              if (fl & 1u) { const uint32_t r = r0; r0 = r1; r1 = r; const float k = k0; k0 = k1; k1 = k; }
              if (fl & 2u) { const uint32_t r = r2; r2 = r3; r3 = r; const float k = k2; k2 = k3; k3 = k; }
              if (fl & 4u) {
                uint32_t r = r0; r0 = r2; r2 = r; r = r1; r1 = r3; r3 = r;
                float k = k0; k0 = k2; k2 = k; k = k1; k1 = k3; k3 = k;
              }
            }
            const uint32_t level = cur & kTlasBit;
            const bool q0 = k0 > 0.f, q1 = k1 > 0.f, q2 = k2 > 0.f, q3 = k3 > 0.f;
            if (!(q0 || q1 || q2 || q3)) {
              pop();
            } else {
              // the first passing child in visiting order is next; the others wait on the stack, nearest on top
              if (q3 && (q0 || q1 || q2)) push_entry(r3 | level, k3);
              if (q2 && (q0 || q1)) push_entry(r2 | level, k2);
              if (q1 && q0) push_entry(r1 | level, k1);
              cur = (q0 ? r0 : (q1 ? r1 : (q2 ? r2 : r3))) | level;
            }
          }
        } else if ((int)cur >= 0) {
          BN_WORK(75);
          const float4* np = node_base + (size_t)(cur & kIndexMask) * 4u;
          const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
          const Slab sl = slab<true>(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), o, inv);
          const Slab sr = slab<true>(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), o, inv);
          const bool pl = slab_pass<true>(sl, t), pr = slab_pass<true>(sr, t);
          const uint32_t level = cur & kTlasBit;
          // left first iff dir[splitAxis] > 0 (BVH.fs:51-56 / Mesh.fs:235-240).  An any-hit query returns the same boolean
          // whatever the visiting order (every box that passes is visited unless a hit ends the walk first), so shadow rays
          // skip the front-to-back bookkeeping (BN_ANY_UNORDERED; measured on the B200: shadow rays +2..3 %)
          const bool lf = (ANY && BN_ANY_UNORDERED) ? true : ((signs >> fbits(n3.z)) & 1u) != 0u;
          const uint32_t left = fbits(n3.x) | level, right = fbits(n3.y) | level;
          if (!pl) {  // branches on the two predicates as they are (no pl | pr to materialise)
            if (pr) cur = right;
            else pop();
          } else if (!pr) {
            cur = left;
          } else {
            cur = lf ? left : right;
            push_entry(lf ? right : left, lf ? sr.tmin : sl.tmin);
          }
        }
        if (__popc(__ballot_sync(kFull, (int)cur >= 0)) < kStayMin) break;
#if BN_SPLIT_REFILL
        if (pend_mask != 0u) break;
#endif
#ifdef BN_EXP_STAY_REFILL
        // experiment queued for the next GPU session (default off): closest-hit rays in tree-TLAS scenes leave
        // the stay loop as soon as a refill is due (BN_EXP_STAY_REFILL idle lanes), instead of stepping on with those lanes empty
        if (!ANY && !flat && !exhausted && __popc(__ballot_sync(kFull, cur == kNone)) >= BN_EXP_STAY_REFILL) break;
#endif
      }
    } else if (nT >= nE && nT >= nS) {
      // ---- phase T: next triangle of the held BLAS leaf (slot order), behind its own
      // AABB test (Mesh.fs:229-233 — load-bearing)
      // keep going while enough lanes still hold a triangle to test (same idea as phase N's repeat)
      for (;;) {
        BN_STAT(1, __popc(__ballot_sync(kFull, (cur >> 30) == 2u)));
        if ((cur >> 30) == 2u) {
          const uint32_t count = (cur >> 27) & 7u;
          const uint32_t tri = (cur & kFirstMask) + tri_k;
          float3 p0, p1, p2;
          if (BN_TRI64 && !ANY) {
            // two 32-B loads (device_scene.h: GTri; a measured switch, off)
            const uintptr_t tri_at = reinterpret_cast<uintptr_t>(tri_base) + (size_t)tri * sizeof(GTri);
            const F8 c0 = ldg256(tri_at), c1 = ldg256(tri_at + 32u);
            p0 = f3(c0.v[0], c0.v[1], c0.v[2]); p1 = f3(c0.v[4], c0.v[5], c0.v[6]); p2 = f3(c1.v[0], c1.v[1], c1.v[2]);
          } else {
            const float4* tp4 = tri_base + (size_t)tri * (sizeof(GTri) / 16u);
            const float4 a = __ldg(tp4), b = __ldg(tp4 + 1), c = __ldg(tp4 + 2);
            p0 = f3(a.x, a.y, a.z); p1 = f3(b.x, b.y, b.z); p2 = f3(c.x, c.y, c.z);
          }
          // Triangle.Bounds (Mesh.fs:19-22): finite vertices => fmin/fmax == minps/maxps
          const float3 lo = f3(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z));
          const float3 hi = f3(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z));
          float tp, u, v;
          bool done = false;
          BN_WORK(slab_pass<true>(slab<true>(lo, hi, o, inv), t) ? 95 : 45);
          if (slab_pass<true>(slab<true>(lo, hi, o, inv), t) && tri_test(p0, p1, p2, o, d, t, tp, u, v)) {
            h_inst = ANY ? 0 : cur_inst; h_prim = (int)tri; h_u = u; h_v = v;  // (an any-hit query reports a boolean: no instance index to keep)
            if (ANY) { finish(); done = true; }
            else t = tp;
          }
          if (!done) {
            ++tri_k;
            if (tri_k >= count) { tri_k = 0; pop(); }
          }
        }
        if (__popc(__ballot_sync(kFull, (cur >> 30) == 2u)) < kStayT) break;
#if BN_SPLIT_REFILL
        if (pend_mask != 0u) break;
#endif
#ifdef BN_EXP_STAY_REFILL
        if (!ANY && !flat && !exhausted && __popc(__ballot_sync(kFull, cur == kNone)) >= BN_EXP_STAY_REFILL) break;
#endif
      }
    } else if (nE >= nS) {
      // ---- phase E: PrimitiveInstance.Intersect (Primitive.fs:111-129); the world AABB
      // test already happened (parent node / ordered scan)
      BN_STAT(2, nE);
      if (isE) {
        BN_WORK(70);
        const uint32_t slot = cur & kIndexMask;
        const float4* ip = reinterpret_cast<const float4*>(sc.inst_trav + slot);
        const float4 m0 = __ldg(ip + 3), m1 = __ldg(ip + 4), m2 = __ldg(ip + 5);
        const uint32_t blas_root = WIDE ? fbits(m1.w) : fbits(m0.w);  // GInstTrav.wroot | root
        tri_k = 0;
        if (fbits(m2.w)) {
          // identity mesh instance: object space == world space, root box == instance box (passed)
          o = wo; d = f3(wdx, wdy, wdz); inv = winv; signs = wsigns_of();
          in_obj = true;
          cur_inst = (int)slot;
          cur = blas_root;
          continue;
        }
        const Mat43 M = load_mat43(ip);
        const float3 oo = transform_point(wo, M);  // Ray.Transform (Ray.fs:19-22)
        const float3 od = transform_dir(f3(wdx, wdy, wdz), M);
        if (fbits(m2.y)) {
          float tp;
          const int root = sphere_test(m2.z, oo, od, t, tp);
          bool done = false;
          if (root) {
            h_inst = (int)slot; h_prim = root - 1; h_u = 0.f; h_v = 0.f;
            if (ANY) { finish(); done = true; }
            else t = tp;
          }
          if (!done) pop();
        } else {
          o = oo; d = od; inv = rcp3(od);
          signs = dir_signs(od);
          in_obj = true;
          cur_inst = (int)slot;
          if (!slab_fast_ok(oo, inv)) {
            io.defer(index);  // re-traced from scratch by the fix-up kernel
            cur = kNone;
          } else {
            // MeshPrimitive pops BLAS node 0 and tests its bounds first (Mesh.fs:224-227)
            const Slab s = slab<true>(f3(m0.x, m0.y, m0.z), f3(m1.x, m1.y, m1.z), o, inv);
            if (slab_pass<true>(s, t)) cur = blas_root;
            else pop();
          }
        }
      }
    } else {
      // ---- phase S (small TLAS): scan the octant-ordered instance list for the next world
      // AABB the ray passes with the CURRENT t.  In the fast path a passing instance box
      // implies passing boxes for all its TLAS ancestors (they contain it, were tested
      // earlier against a larger t, and no lane can be NaN), so skipping the interior TLAS
      // nodes changes neither the set nor the order of instances entered.
      BN_STAT(4, nS);
      if (isS) {
        const float4* fp = reinterpret_cast<const float4*>(sc.flat_tlas + (size_t)(wsigns_of() & 7u) * n_inst);
        bool found = false;
        while (tl_pos != 0u) {
          const uint32_t k = (uint32_t)__ffs((int)tl_pos) - 1u;
          tl_pos &= tl_pos - 1u;
          BN_WORK(30);
          const float4 a = __ldg(fp + 2u * k), b = __ldg(fp + 2u * k + 1u);
          if (ANY || flat_pass(a, b, wo, winv, t)) {
#ifdef BN_EXP_SCAN_LEAF
            // experiment queued for the next GPU session (default off; found on the warp emulator): an identity
            // mesh instance whose whole BLAS is ONE leaf (the Cornell-box walls: two triangles) is tested right here, in slot
            // order behind each triangle's own box, instead of sending the lane through a phase-T round trip per wall
            if (fbits(b.w) != kNone && (fbits(b.w) & kLeafBit)) {
              const uint32_t leaf = fbits(b.w);
              const uint32_t count = (leaf >> 27) & 7u, first = leaf & kFirstMask;
              const float3 wd = f3(wdx, wdy, wdz);
              bool occluded = false;
              for (uint32_t j = 0; j < count; ++j) {
                const float4* tp4 = tri_base + (size_t)(first + j) * (sizeof(GTri) / 16u);
                const float4 ta = __ldg(tp4), tb = __ldg(tp4 + 1), tc = __ldg(tp4 + 2);
                const float3 p0 = f3(ta.x, ta.y, ta.z), p1 = f3(tb.x, tb.y, tb.z), p2 = f3(tc.x, tc.y, tc.z);
                const float3 lo = f3(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z));
                const float3 hi = f3(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z));
                float tp, u, v;
                BN_WORK(slab_pass<true>(slab<true>(lo, hi, wo, winv), t) ? 95 : 45);
                if (slab_pass<true>(slab<true>(lo, hi, wo, winv), t) && tri_test(p0, p1, p2, wo, wd, t, tp, u, v)) {
                  h_inst = (int)fbits(a.w); h_prim = (int)(first + j); h_u = u; h_v = v;
                  if (ANY) { occluded = true; break; }
                  t = tp;
                }
              }
              if (ANY && occluded) { finish(); found = true; break; }
              continue;  // next candidate of the scan, against the (possibly shorter) t
            }
#endif
            if (fbits(b.w) != kNone) {
              // identity mesh instance: what phase E would do (object space == world space, the BLAS
              // root box is this box), done here so that the ray goes straight to its N / T phase
              o = wo; d = f3(wdx, wdy, wdz); inv = winv; signs = wsigns_of();
              in_obj = true;
              cur_inst = (int)fbits(a.w);
              tri_k = 0;
              cur = fbits(b.w);
            } else {
              cur = kLeafBit | kTlasBit | fbits(a.w);
            }
            found = true;
            break;
          }
        }
        if (!found) finish();
      }
    }
  }
#ifdef BN_TRAV_STATS
  if (lane == 0)
    for (int k = 0; k < 5; ++k) { atomicAdd(&g_trav_stats[(ANY ? 10 : 0) + k], st_cnt[k]); atomicAdd(&g_trav_stats[(ANY ? 10 : 0) + 5 + k], st_sum[k]); }
#endif
}

// Fix-up: the deferred rays, one per thread, exact path.
template <bool ANY, class IO>
BN_DEV void traverse_deferred(const DScene& sc, IO& io, const int* __restrict__ list, int n_deferred) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_deferred; k += gridDim.x * blockDim.x) {
    const int i = list[k];
    float3 o, d;
    float t;
    io.load(i, o, d, t);
    TraceResult r;
    trace_exact<ANY>(sc, o, d, t, r);
    io.store(i, r);
  }
}

}  // namespace bn
