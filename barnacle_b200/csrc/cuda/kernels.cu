// Wavefront path tracer for sm_100a: the kernels and the C-ABI entry points
// that replace ProgressiveIntegrator.Render + PathTracingIntegrator.Li
// (Base/Integrator.fs:22-55, Extensions/Integrator/PathTracing.fs:14-81).
//
// One "wave" = (a range of 8x4 pixel blocks) x (a range of sample ids), at most
// `wave_capacity` camera paths.  Per wave:
//   raygen                        -> compacted path state (bounce 0)
//   for bounce in 0..maxDepth-1:  extend (closest hit) -> shade -> shadow (any hit + connect)
//   accumulate                    -> film[pix] = fma(1/spp, L, film[pix]) in sample order
// Kernels are persistent (grid = SMs x resident CTAs); warps fetch 32 queue
// entries at a time from a global cursor and append to the next queue with
// warp-aggregated atomics.  All paths of a wave are at the same depth, so the
// reference's `depth` is the bounce index.
//
// Compiled with -fmad=false (see vecmath.cuh).  No tensor cores: nothing here is
// a dense contraction.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/barnacle_b200.h"
#include "device_scene.h"
#include "scene_convert.h"
#include "ray_sort.cuh"
#include "shade.cuh"
#include "traverse.cuh"
#include "vecmath.cuh"

namespace bnhost {
void set_error(const std::string& msg);
}

namespace bn {

constexpr int kBlock = 128;
#ifndef BN_TRAV_MIN_BLOCKS
#define BN_TRAV_MIN_BLOCKS 9   // resident CTAs per SM the traversal kernels are compiled for (56 registers: the cold per-ray state is in shared memory)
#endif
#ifndef BN_SHADE_MIN_BLOCKS
#define BN_SHADE_MIN_BLOCKS 7
#endif
#ifndef BN_COUNTER_STRIDE
#define BN_COUNTER_STRIDE 64   // ints between two queue counters / cursors (256 B)
#endif
#ifndef BN_INKERNEL_DRAIN
#define BN_INKERNEL_DRAIN 0    // 1: the last CTA of a traversal launch drains the deferred rays itself (no fix-up launch). Measured on the B200
                               // (profiles/r02_ab_session1.log): 26 instead of 42 launches per wave, but the drain's code costs the hot loop
                               // three spilled registers — C1 -0.6 %, C2 -0.7 %, C3 -1.9 %, C4 -2.4 % — while the fix-up launches themselves
                               // cost nothing measurable (same build with BN_SEPARATE_FIXUP: +0.1 %).  Off.
#endif
#ifndef BN_CAND_PREPASS
#define BN_CAND_PREPASS 0      // 1: small-TLAS scenes get their candidate masks from a full-width pre-pass kernel (k_candidates) before each
                               // traversal launch instead of from the refilled lanes of the persistent loop.  Measured on the B200
                               // (profiles/r02_ab_session12_*.log) and OFF: C1 -11.7 %, C2 -5.3 %: the pass is cheap where it is (other warps
                               // hide its latency) and the pre-pass adds a kernel that runs at 40 % of the issue rate
#endif
#ifndef BN_TRAV_GRID_MULT
#define BN_TRAV_GRID_MULT 9    // persistent grid = SMs x this
#endif

struct WaveParams {
  int width, height;
  int spp, max_depth, rr_depth, frame_id;
  int x0, y0, x1, y1;
  int nbx;            // 8x4 pixel blocks per row of the window
  int block_begin;    // first pixel block of this wave
  int n_blocks;       // pixel blocks in this wave
  int sample_begin;   // first sampleId of this wave
  int n_samples;      // samples per pixel in this wave
  uint32_t flags;
  float inv_spp;
  int il_count, il_index;  // tile-row interleave (BnRenderParams.interleave_*)
  int integrator;          // BN_INTEGRATOR_*
};

// ---- warp helpers --------------------------------------------------------------
BN_DEV int lane_id() { return threadIdx.x & 31; }

// Each warp claims 32 consecutive queue entries.
BN_DEV int warp_fetch(int* cursor) {
  int base = 0;
  if (lane_id() == 0) base = atomicAdd(cursor, 32);
  return __shfl_sync(0xffffffffu, base, 0);
}
// Warp-aggregated append: one atomicAdd per warp.  Must be called by all 32 lanes.
BN_DEV int warp_append(int* counter, bool pred) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  int base = 0;
  if (lane_id() == 0 && m) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + __popc(m & ((1u << lane_id()) - 1u));
}

BN_DEV void wave_pixel(const WaveParams& wp, int pl, int& x, int& y) {
  const int block = wp.block_begin + (pl >> 5);
  const int lane = pl & 31;
  x = wp.x0 + (block % wp.nbx) * 8 + (lane & 7);
  const int brow = block / wp.nbx;  // 4-pixel block row among the rows this call owns
  const int tile_row = (brow >> 2) * wp.il_count + wp.il_index;
  y = wp.y0 + tile_row * 16 + (brow & 3) * 4 + (lane >> 3);
}

// ---- raygen: RenderTile's sample loop head (Integrator.fs:34-39) ------------------
__global__ void __launch_bounds__(kBlock) k_raygen(DScene sc, WaveParams wp, float4* __restrict__ s0, float4* __restrict__ s1,
                                                   float4* __restrict__ s2, float4* __restrict__ rad, int* n_active) {
  const int npw = wp.n_blocks * 32;
  const int total = npw * wp.n_samples;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < total; base += gridDim.x * blockDim.x) {
    const int pid = base + lane_id();
    bool alive = false;
    float3 o = splat(0.f), d = splat(0.f);
    uint32_t rng = 0;
    if (pid < total) {
      rad[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int sl = pid / npw, pl = pid - sl * npw;
      int x, y;
      wave_pixel(wp, pl, x, y);
      if (x < wp.x1 && y < wp.y1) {
        alive = true;
        rng = xxhash32_three((uint32_t)x, (uint32_t)y, (uint32_t)(wp.frame_id * wp.spp + wp.sample_begin + sl));
        const float upx = lcg(rng), upy = lcg(rng);  // Next2D: X then Y (Sampler.fs:16)
        const float ulx = lcg(rng), uly = lcg(rng);
        primary_ray(sc.cam, wp.width, wp.height, x, y, upx, upy, ulx, uly, o, d);
      }
    }
    const int pos = warp_append(n_active, alive);
    if (alive) {
      s0[pos] = make_float4(o.x, o.y, o.z, d.x);
      s1[pos] = make_float4(d.y, d.z, 1.f, 1.f);  // beta = 1
      s2[pos] = make_float4(1.f, 0.f, __uint_as_float(rng), __int_as_float(pid));
    }
  }
}

BN_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- extend: closest hit for every live path ----------------------------------------
struct DeferList {  // rays that need the exact traversal (fix-up kernel)
  int* count;
  int* list;
  BN_DEV void push(int i) const { list[atomicAdd(count, 1)] = i; }
};
struct ExtendIO {
  const float4* __restrict__ s0;
  const float4* __restrict__ s1;
  float4* __restrict__ hits;
  const int* __restrict__ n_ptr;
  int* cur;
  DeferList deferred;
  const uint32_t* __restrict__ cand_words;  // k_candidates' output (small-TLAS scenes)
  // ray_sort.cuh: slot i of the ordered queue is path perm[i] of the state planes (nullptr: i itself).  The refill that
  // claims slot i gathers the two planes it needs anyway (origin, direction) and leaves them at slot i of the planes t0, t1,
  // so that shade reads them coalesced (two coalesced stores per ray here); shade fetches the third plane through perm.
  const uint32_t* __restrict__ perm;
  float4* __restrict__ t0;
  float4* __restrict__ t1;
  static constexpr bool kHasCand = BN_CAND_PREPASS != 0;
  BN_DEV uint32_t cand(int i) const { return cand_words[i]; }
  BN_DEV int count() const { return *n_ptr; }
  BN_DEV int* cursor() const { return cur; }
  BN_DEV void load(int i, float3& o, float3& d, float& t) const {
    if (perm) {
      // past L1 (ld.global.cg) and streaming stores: neither the gathered records nor the ordered copies are read again by
      // this kernel (neutral to +1 %, profiles/r02_ab_session21_*.log)
      const int src = (int)__ldcg(perm + i);
      const float4 a = __ldcg(s0 + src), b = __ldcg(s1 + src);
      __stcs(t0 + i, a); __stcs(t1 + i, b);
      o = f3(a.x, a.y, a.z); d = f3(a.w, b.x, b.y);
    } else {
      const float4 a = s0[i], b = s1[i];
      o = f3(a.x, a.y, a.z); d = f3(a.w, b.x, b.y);
    }
    t = CUDART_INF_F;  // PathTracing.fs:25
  }
  BN_DEV void load_ray(int i, float3& o, float3& d, float& t) const { load(i, o, d, t); }
  BN_DEV void store(int i, const TraceResult& r) const {
    hits[i] = make_float4(r.t, __int_as_float(r.inst), __int_as_float(r.prim), 0.f);
  }
  BN_DEV void defer(int i) const { deferred.push(i); }
  BN_DEV void prefetch(int i) const {
    if (perm) {
      // two stages: the perm line one more window ahead (32 entries per 128-B line), and through the entry that an earlier
      // refill prefetched, the state it points at
      if ((i & 31) == 0 && i + kPrefetchAhead < *n_ptr) prefetch_l2(perm + i + kPrefetchAhead);
      const int src = (int)__ldcg(perm + i);
      prefetch_l2(s0 + src); prefetch_l2(s1 + src);
    } else {
      prefetch_l2(s0 + i); prefetch_l2(s1 + i);
    }
    if (kHasCand && (i & 31) == 0) prefetch_l2(cand_words + i);  // 32 words per 128-B line
  }
};
// The deferred rays (zero / denormal direction component: normally none, a handful at most) re-traced with the exact form.
// Out of line on purpose: the hot loop's register allocation must not see this code.
template <bool ANY, class IO>
BN_DEV void drain_deferred(const DScene& sc, const IO& io) {
  const int n_deferred = *reinterpret_cast<volatile int*>(io.deferred.count);
  for (int k = threadIdx.x; k < n_deferred; k += blockDim.x) {
    const int i = reinterpret_cast<volatile int*>(io.deferred.list)[k];
    float3 o, d;
    float t;
    io.load(i, o, d, t);
    TraceResult r;
    trace_exact<ANY>(sc, o, d, t, r);
    io.store(i, r);
  }
}
// `done` != nullptr: the CTA that finishes LAST (ticket counter) drains the deferred list itself, so no fix-up launch
// follows (render path: 2 launches per bounce saved, whether or not a ray was deferred).  `done` == nullptr: the caller
// launches k_traverse_fixup afterwards (bn_trace and BN_RENDER_FORCE_EXACT, where deferral is the rule, not the exception).
template <bool ANY, bool WIDE, class IO>
__global__ void __launch_bounds__(kBlock, BN_TRAV_MIN_BLOCKS) k_traverse(const __grid_constant__ DScene sc, const __grid_constant__ IO io, int* done) {
  __shared__ uint32_t s_cold[kTravColdWords * kBlock];
  traverse_persistent<ANY, WIDE>(sc, io, s_cold + threadIdx.x, kBlock);
#if BN_INKERNEL_DRAIN
  if (done != nullptr) {
    __shared__ int s_last;
    __threadfence();  // this thread's deferred-list appends and result stores, before the ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1) == (int)gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (s_last) {
      __threadfence();
      drain_deferred<ANY>(sc, io);
    }
  }
#endif
}
template <bool ANY, class IO>
__global__ void __launch_bounds__(kBlock) k_traverse_fixup(DScene sc, IO io) {
  traverse_deferred<ANY>(sc, io, io.deferred.list, *io.deferred.count);
}

// Small-TLAS scenes: the candidate word of every queued ray (traverse.cuh: candidate_word — the fast-form check and the pass
// over the <= 16 octant-ordered instance boxes with the ray's initial t), one thread per ray with all 32 lanes of a warp
// busy.  In the persistent loop this pass ran on the 14-20 lanes a refill fills and was 16 % of the extend kernel's
// instructions on C2; here it streams (32 B in, 4 B out per ray).  Same device function, same bits.
template <class IO>
__global__ void __launch_bounds__(256) k_candidates(const __grid_constant__ DScene sc, const __grid_constant__ IO io, uint32_t* __restrict__ cand) {
  const int n = io.count();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float3 o, d;
    float t;
    io.load_ray(i, o, d, t);
    cand[i] = candidate_word(sc, o, d, t);
  }
}

// host side: the 4-wide-node kernel when the scene has such nodes (sc.wide), else the binary-node kernel
template <bool ANY, class IO>
void launch_traverse(int grid, cudaStream_t stream, const DScene& sc, const IO& io, int* done) {
  if constexpr (io_has_cand<IO>::value) {
    if (sc.flat_tlas != nullptr) k_candidates<IO><<<grid, 256, 0, stream>>>(sc, io, const_cast<uint32_t*>(io.cand_words));
  }
  if (sc.wide != nullptr) k_traverse<ANY, true, IO><<<grid, kBlock, 0, stream>>>(sc, io, done);
  else k_traverse<ANY, false, IO><<<grid, kBlock, 0, stream>>>(sc, io, done);
}

// ---- shade: one iteration of Li's loop body (PathTracing.fs:30-79) --------------------
__global__ void __launch_bounds__(kBlock, BN_SHADE_MIN_BLOCKS) k_shade(DScene sc, WaveParams wp, int bounce, const float4* __restrict__ s0, const float4* __restrict__ s1,
                                                  const float4* __restrict__ s2, const float4* __restrict__ hits, float4* __restrict__ o0,
                                                  float4* __restrict__ o1, float4* __restrict__ o2, float4* __restrict__ q0, float4* __restrict__ q1,
                                                  float4* __restrict__ q2, float4* __restrict__ q3, float4* __restrict__ rad, const int* __restrict__ n_ptr,
                                                  int* n_out, int* cursor, unsigned long long* shadow_ref, uint16_t* __restrict__ key_out, const uint32_t* __restrict__ perm) {
  // n_out: the next bounce's path counter; n_out + 1: this bounce's shadow-queue counter (one 8-byte word, see the append below)
  const int n = *n_ptr;
  unsigned ref_total = 0;  // lane 0: reference-equivalent shadow rays of this warp's chunks
  // Claim pipeline, three chunks of 32 paths deep, so that nothing on this loop's critical path waits for a round trip it
  // could have started earlier (ncu on the ordered kernel: the shuffle behind the cursor's atomic and the permutation entry
  // in front of the third plane's gather were 12 % of the stall samples):
  //   chunk c0: shaded now | chunk c1: its inputs prefetched into L2 now, through the permutation entries fetched an
  //   iteration ago | chunk c2: its base comes out of the atomic issued an iteration ago, its permutation entries are
  //   fetched now | the atomic for the chunk after that is issued now.
  const int lane = lane_id();
  auto order = [&](int j) { return (perm && j < n) ? (int)__ldcg(perm + j) : j; };
  int raw = 0;
  if (lane == 0) raw = atomicAdd(cursor, 96);
  int c0 = __shfl_sync(0xffffffffu, raw, 0), c1 = c0 + 32;
  raw = c0 + 64;
  int src0 = order(c0 + lane), src1 = order(c1 + lane);
  for (;;) {
    const int base = c0;
    if (base >= n) break;
    const int c2 = __shfl_sync(0xffffffffu, raw, 0);
    if (lane == 0) raw = atomicAdd(cursor, 32);
    const int src2 = order(c2 + lane);
    if (c1 + lane < n) {
      const int j = c1 + lane;
      prefetch_l2(s0 + j); prefetch_l2(s1 + j); prefetch_l2(s2 + src1); prefetch_l2(hits + j);
    }
    const int src = src0;
    c0 = c1; c1 = c2; src0 = src1; src1 = src2;
    const int i = base + lane;
    bool alive = false, has_shadow = false, ref_shadow = false;
    float3 P = splat(0.f), nd = splat(0.f), beta = splat(0.f);
    float bs_pdf = 0.f;
    uint32_t rng = 0;
    int pid = 0;
    float3 sh_wi = splat(0.f), sh_a = splat(0.f), sh_b = splat(0.f);
    float sh_tmax = 0.f;
    if (i < n) {
      const float4 a = s0[i], b = s1[i], c = perm ? __ldcg(s2 + src) : s2[i], h = hits[i];  // third plane: where the previous shade left it (ray_sort.cuh)
      shade_lane(sc, wp.integrator, wp.rr_depth, wp.max_depth, wp.flags, bounce, a, b, c, h, rad, alive, has_shadow, ref_shadow, P, nd, beta, bs_pdf, rng, pid,
                 sh_wi, sh_a, sh_b, sh_tmax);
    }
    // both appends with ONE atomic: the shadow-queue counter is the int right after the path counter (render_waves), so a
    // 64-bit add claims the slots of both queues in one round trip (two round trips were 7 % of this kernel's stall samples)
    const unsigned m_alive = __ballot_sync(0xffffffffu, alive), m_shadow = __ballot_sync(0xffffffffu, has_shadow);
    unsigned long long claimed = 0ull;
    if (lane == 0 && (m_alive | m_shadow))
      claimed = atomicAdd(reinterpret_cast<unsigned long long*>(n_out), (unsigned long long)__popc(m_alive) | ((unsigned long long)__popc(m_shadow) << 32));
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    const unsigned lt = (1u << lane) - 1u;
    const int pos = (int)(unsigned)(claimed & 0xffffffffull) + __popc(m_alive & lt);
    const int spos = (int)(unsigned)(claimed >> 32) + __popc(m_shadow & lt);
    if (alive) {
      o0[pos] = make_float4(P.x, P.y, P.z, nd.x);
      o1[pos] = make_float4(nd.y, nd.z, beta.x, beta.y);
      o2[pos] = make_float4(beta.z, bs_pdf, __uint_as_float(rng), __int_as_float(pid));
      if (key_out) key_out[pos] = (uint16_t)sort_key(sc.sort_grid, P, nd);  // for the re-ordering before the next extend (ray_sort.cuh)
    }
    if (has_shadow) {
      q0[spos] = make_float4(P.x, P.y, P.z, sh_wi.x);
      q1[spos] = make_float4(sh_wi.y, sh_wi.z, sh_tmax, __int_as_float(pid));
      q2[spos] = make_float4(sh_a.x, sh_a.y, sh_a.z, sh_b.x);
      q3[spos] = make_float4(sh_b.y, sh_b.z, 0.f, 0.f);
    }
    ref_total += (unsigned)__popc(__ballot_sync(0xffffffffu, ref_shadow));
  }
  if (lane_id() == 0 && ref_total) atomicAdd(shadow_ref, (unsigned long long)ref_total);
}

// ---- shadow: any hit + connect (PathTracing.fs:47-59) ----------------------------------
struct ShadowIO {
  const float4* __restrict__ q0;
  const float4* __restrict__ q1;
  const float4* __restrict__ q2;
  const float4* __restrict__ q3;
  float4* __restrict__ rad;
  const int* __restrict__ n_ptr;
  int* cur;
  DeferList deferred;
  const uint32_t* __restrict__ cand_words;  // k_candidates' output (small-TLAS scenes)
  static constexpr bool kHasCand = BN_CAND_PREPASS != 0;
  BN_DEV uint32_t cand(int i) const { return cand_words[i]; }
  BN_DEV int count() const { return *n_ptr; }
  BN_DEV int* cursor() const { return cur; }
  BN_DEV void load_ray(int i, float3& o, float3& d, float& t) const {
    const float4 a = q0[i], b = q1[i];
    o = f3(a.x, a.y, a.z); d = f3(a.w, b.x, b.y);
    t = b.z;
  }
  BN_DEV void load(int i, float3& o, float3& d, float& t) const {
    const float4 a = q0[i], b = q1[i];
    o = f3(a.x, a.y, a.z); d = f3(a.w, b.x, b.y);
    t = b.z;
    // what store() reads if the ray turns out unoccluded: in L2 by the time the traversal is done
    prefetch_l2(q2 + i); prefetch_l2(q3 + i); prefetch_l2(rad + __float_as_int(b.w));
  }
  BN_DEV void prefetch(int i) const {
    prefetch_l2(q0 + i); prefetch_l2(q1 + i);
    if (kHasCand && (i & 31) == 0) prefetch_l2(cand_words + i);
  }
  BN_DEV void store(int i, const TraceResult& r) const {
    if (r.hit) return;  // occluded
    const float4 b = q1[i], c = q2[i], e = q3[i];
    const int pid = __float_as_int(b.w);
    const float4 L4 = rad[pid];
    const float3 L = vfma(f3(c.x, c.y, c.z), f3(c.w, e.x, e.y), f3(L4.x, L4.y, L4.z));
    rad[pid] = make_float4(L.x, L.y, L.z, 0.f);
  }
  BN_DEV void defer(int i) const { deferred.push(i); }
};

// ---- accumulate: accum = fma(1/spp, radiance, accum); Film.SetPixel -------------------
// (Integrator.fs:41-44, Film.fs:41-46).  One thread per pixel walks its samples in
// sampleId order, continuing the chain left in the film by earlier waves.
__global__ void __launch_bounds__(kBlock) k_accumulate(WaveParams wp, const float4* __restrict__ rad, float* __restrict__ film) {
  const int npw = wp.n_blocks * 32;
  for (int pl = blockIdx.x * blockDim.x + threadIdx.x; pl < npw; pl += gridDim.x * blockDim.x) {
    int x, y;
    wave_pixel(wp, pl, x, y);
    if (x >= wp.x1 || y >= wp.y1) continue;
    float* px = film + ((size_t)(wp.height - y - 1) * wp.width + x) * 3;
    float3 acc = f3(px[0], px[1], px[2]);
    for (int s = 0; s < wp.n_samples; ++s) {
      const float4 L = rad[(size_t)s * npw + pl];
      const float3 radiance = f3(L.x, L.y, L.z) * (1.f / 1.f);  // * rcp(camera pdf), pdf == 1 (Pinhole.fs:27)
      acc = vfma(splat(wp.inv_spp), radiance, acc);
    }
    px[0] = acc.x; px[1] = acc.y; px[2] = acc.z;
  }
}

// Film.PostProcess + Rgba32 (Film.fs:21-30,64): ACES / gamma / identity, clamp, 8-bit RGBA
__global__ void __launch_bounds__(256) k_film_to_rgba8(const float* __restrict__ film, size_t npix, int tone, uchar4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float x = film[i * 3 + k];
      if (tone == 1) x = x * (2.51f * x + 0.03f) / (x * (2.43f * x + 0.59f) + 0.14f);
      else if (tone == 2) x = powf(x, 1.f / 2.2f);
      c[k] = !(x > 0.f) ? 0.f : (x > 1.f ? 1.f : x);  // Vector3.Clamp; NaN -> 0
    }
    out[i] = make_uchar4((unsigned char)(c[0] * 255.f + 0.5f), (unsigned char)(c[1] * 255.f + 0.5f), (unsigned char)(c[2] * 255.f + 0.5f), 255);
  }
}

// per-path radiance export (bn_render_radiance): [sample][(y-y0)*rw + (x-x0)][3]
__global__ void __launch_bounds__(kBlock) k_export_radiance(WaveParams wp, const float4* __restrict__ rad, float* __restrict__ out, int sample_origin) {
  const int npw = wp.n_blocks * 32;
  const int total = npw * wp.n_samples;
  const int rw = wp.x1 - wp.x0, rh = wp.y1 - wp.y0;
  for (int pid = blockIdx.x * blockDim.x + threadIdx.x; pid < total; pid += gridDim.x * blockDim.x) {
    const int sl = pid / npw, pl = pid - sl * npw;
    int x, y;
    wave_pixel(wp, pl, x, y);
    if (x >= wp.x1 || y >= wp.y1) continue;
    const float4 L = rad[pid];
    float* o = out + (((size_t)(wp.sample_begin + sl - sample_origin) * rh + (y - wp.y0)) * rw + (x - wp.x0)) * 3;
    o[0] = L.x; o[1] = L.y; o[2] = L.z;
  }
}

// ---- fixed-batch traversal (bn_trace) -------------------------------------------------
template <bool ANY>
struct TraceIO {
  DScene sc;
  const BnRay* __restrict__ rays;
  BnHit* __restrict__ hits;
  int n;
  int* cur;
  DeferList deferred;
  BN_DEV int count() const { return n; }
  BN_DEV int* cursor() const { return cur; }
  BN_DEV void defer(int i) const { deferred.push(i); }
  BN_DEV void prefetch(int i) const { prefetch_l2(rays + i); }
  BN_DEV void load(int i, float3& o, float3& d, float& t) const {
    const BnRay r = rays[i];
    o = f3(r.origin[0], r.origin[1], r.origin[2]); d = f3(r.direction[0], r.direction[1], r.direction[2]);
    t = r.tmax;
  }
  BN_DEV void store(int i, const TraceResult& res) const {
    BnHit out;
    const bool hit = res.hit;
    const float t = res.t;
    const int inst = res.inst;
    if (ANY) {
      out.t = 0.f; out.u = 0.f; out.v = 0.f; out.instance = hit ? 1 : 0; out.primitive = 0;
    } else {
      out.t = t; out.u = res.u; out.v = res.v; out.instance = inst; out.primitive = res.prim;
      if (hit) {
        // BnHit.primitive is the BLAS-order index WITHIN the mesh; the kernel carries scene-wide indices
        out.primitive = res.prim - (int)sc.inst_trav[inst].tri_base;
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(sc.inst_head + inst));
        if (__float_as_uint(h0.w) & 0x80000000u) {  // sphere uv (Sphere.fs:55-56), libdevice atan2/acos
          const BnRay r = rays[i];
          const float3 o = f3(r.origin[0], r.origin[1], r.origin[2]), d = f3(r.direction[0], r.direction[1], r.direction[2]);
          const Mat43 M = load_mat43(reinterpret_cast<const float4*>(sc.inst_w2o + inst));
          const float3 nn = normalize(point_at(transform_point(o, M), transform_dir(d, M), t));
          out.u = atan2f(nn.z, nn.x) / (2.f * kPi) + 0.5f;
          out.v = acosf(nn.y) / kPi;
          out.primitive = 0;
        }
      }
    }
    hits[i] = out;
  }
};
// ---- L2 read-bandwidth probe (the traversal's scene data is L2-resident, so the L2
// rate is the second roofline denominator).  Every thread sweeps a buffer that fits L2 with
// 16-B ld.global.cg loads (L1 bypassed), `iters` times; returns bytes read / device time.
__global__ void __launch_bounds__(256) k_l2_sweep(const uint4* __restrict__ buf, size_t n_vec, int iters, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint4 v = __ldcg(buf + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0xDEADBEEFu) *sink = acc;  // keeps the loads alive
}

}  // namespace bn

// =================================================================================
// host side of the C ABI
// =================================================================================
using namespace bn;

#include "scene_internal.h"

namespace {

// Wave buffers are large (hundreds of MB) and identical from scene to scene; a managed host
// that re-creates the scene every frame must not pay cudaMalloc/cudaFree for
// them each time.  One parked set per device.
struct WaveBuffers {
  int device = -1;
  size_t cap = 0;
  float4* state[3] = {nullptr, nullptr, nullptr};
  uint32_t* sort_perm = nullptr;
  uint16_t* sort_key = nullptr;
  uint32_t* sort_bins = nullptr;
  float4* hits = nullptr;
  float4* shq = nullptr;
  float4* rad = nullptr;
  int* defer_list = nullptr;
  uint32_t* cand = nullptr;
  // the small per-render buffers travel with the bundle so that creating / destroying a scene
  // (the e2e path does both per frame) allocates nothing once a device is warm
  int* counters = nullptr;
  size_t counters_len = 0;
  unsigned long long* shadow_ref = nullptr;
  float* film = nullptr;
  size_t film_len = 0;
  void* arena = nullptr;   // the dead scene's array allocation, reused by the next scene if it is large enough
  size_t arena_bytes = 0;
};
std::mutex g_pool_mutex;
std::vector<WaveBuffers> g_pool;

bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  bnhost::set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}
#define BN_CUDA(call)                                  \
  do {                                                 \
    if (!cuda_ok((call), #call)) return BN_ERR_CUDA;   \
  } while (0)

void adopt_parked_buffers(BnScene* s);

// The flattened device layout of one scene description, ready to upload: every array in ONE host image (256-B aligned
// slices) so that bn_scene_create costs one cudaMalloc and one host -> device copy whatever the number of arrays.
// Built by stage_scene(); immutable afterwards and shared by every upload of a multi-device scene.
struct StagedSlices {  // byte offsets into the image
  size_t nodes, inst_trav, inst_head, inst_w2o, inst_o2w, meshes, tris, alias, sphere_radii, materials, lights, light_inst, flat_tlas, wide;
};
constexpr size_t kNoSlice = ~(size_t)0;

}  // namespace

struct bnint::Staged {
  std::vector<unsigned char> image;
  StagedSlices at{};
  bn::GTree tlas{};
  uint32_t tlas_wroot = 0, n_inst = 0, n_light_inst = 0, all_finite = 0;
  bool pinned = false;
  std::vector<unsigned char> key;  // the input arrays this image was made from (exact-comparison cache key; camera excluded)
  ~Staged() { if (pinned) cudaHostUnregister(image.data()); }
};

namespace {

template <class T>
size_t stage_add(std::vector<unsigned char>& image, const std::vector<T>& v) {
  const size_t off = (image.size() + 255) & ~(size_t)255;
  const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  image.resize(off + bytes);
  if (!v.empty()) std::memcpy(image.data() + off, v.data(), v.size() * sizeof(T));
  return off;
}

// Everything of the description that the device layout depends on, as one byte string (the camera is applied per scene).
void scene_key(const BnSceneDesc& d, uint32_t env_flags, std::vector<unsigned char>& key) {
  key.clear();
  auto put = [&](const void* p, size_t count, size_t size) {
    const uint64_t n = count;
    const size_t at = key.size();
    key.resize(at + 8 + count * size);
    std::memcpy(key.data() + at, &n, 8);
    if (count) std::memcpy(key.data() + at + 8, p, count * size);
  };
  put(&env_flags, 1, 4);
  put(d.tlas_nodes, d.tlas_node_count, sizeof(BnBVHNode)); put(d.instances, d.instance_count, sizeof(BnInstance));
  put(d.light_instances, d.light_instance_count, 4); put(d.meshes, d.mesh_count, sizeof(BnMesh)); put(d.vertices, d.vertex_count, 12);
  put(d.triangles, d.triangle_count, 12); put(d.blas_nodes, d.blas_node_count, sizeof(BnBVHNode)); put(d.alias, d.alias_count, sizeof(BnAliasEntry));
  put(d.sphere_radii, d.sphere_count, 4); put(d.materials, d.material_count, sizeof(BnMaterial)); put(d.lights, d.light_count, sizeof(BnLight));
}
bool scene_key_matches(const BnSceneDesc& d, uint32_t env_flags, const std::vector<unsigned char>& key) {
  size_t at = 0;
  auto same = [&](const void* p, size_t count, size_t size) {
    if (at + 8 > key.size()) return false;
    uint64_t n = 0;
    std::memcpy(&n, key.data() + at, 8);
    if (n != count || at + 8 + count * size > key.size()) return false;
    if (count && std::memcmp(key.data() + at + 8, p, count * size) != 0) return false;
    at += 8 + count * size;
    return true;
  };
  return same(&env_flags, 1, 4) && same(d.tlas_nodes, d.tlas_node_count, sizeof(BnBVHNode)) && same(d.instances, d.instance_count, sizeof(BnInstance)) &&
         same(d.light_instances, d.light_instance_count, 4) && same(d.meshes, d.mesh_count, sizeof(BnMesh)) && same(d.vertices, d.vertex_count, 12) &&
         same(d.triangles, d.triangle_count, 12) && same(d.blas_nodes, d.blas_node_count, sizeof(BnBVHNode)) && same(d.alias, d.alias_count, sizeof(BnAliasEntry)) &&
         same(d.sphere_radii, d.sphere_count, 4) && same(d.materials, d.material_count, sizeof(BnMaterial)) && same(d.lights, d.light_count, sizeof(BnLight)) &&
         at == key.size();
}

std::mutex g_stage_mutex;
std::shared_ptr<const bnint::Staged> g_last_staged;  // the most recent scene: a host that re-creates its scene every frame hits it

size_t wave_capacity_paths() {
  const char* e = std::getenv("BN_WAVE_PATHS");
  // 64 Mi paths (13 GB of queues on a 180 GB device): every launch of a wave, the deep bounces with
  // few live paths above all, stays large enough that the drain at the end of each persistent kernel
  // and the per-launch costs are a small share (measured on C2: 4 Mi -> 16 Mi +12%, 16 -> 64 Mi +5%,
  // 128 Mi +0.6% more)
  size_t v = e ? (size_t)std::strtoull(e, nullptr, 10) : (size_t)64 << 20;
  v = std::max<size_t>(v, 1024);
  return (v + 31) & ~(size_t)31;
}

void free_wave_buffers(WaveBuffers& w) {
  for (void* p : {(void*)w.state[0], (void*)w.state[1], (void*)w.state[2], (void*)w.sort_perm, (void*)w.sort_key, (void*)w.sort_bins, (void*)w.hits, (void*)w.shq, (void*)w.rad, (void*)w.defer_list, (void*)w.cand, (void*)w.counters, (void*)w.shadow_ref, (void*)w.film, w.arena})
    if (p) cudaFree(p);
  w = WaveBuffers();
}

void release_wave_buffers(BnScene* s) {  // park the scene's buffers for the next scene on this device
  if (s->cap == 0 && !s->counters && !s->shadow_ref && !s->film && !s->arena) return;
  WaveBuffers w;
  const bool park = std::getenv("BN_NO_BUFFER_CACHE") == nullptr;  // a host that shares the GPU can opt out of the retained footprint
  w.device = s->device; w.cap = s->cap;
  w.state[0] = s->state[0]; w.state[1] = s->state[1]; w.state[2] = s->state[2]; w.sort_perm = s->sort_perm; w.sort_key = s->sort_key; w.sort_bins = s->sort_bins; w.hits = s->hits; w.shq = s->shq; w.rad = s->rad; w.defer_list = s->defer_list; w.cand = s->cand;
  w.counters = s->counters; w.counters_len = s->counters_len; w.shadow_ref = s->shadow_ref; w.film = s->film; w.film_len = s->film_len;
  w.arena = s->arena; w.arena_bytes = s->arena_bytes;
  s->arena = nullptr; s->arena_bytes = 0;
  s->cap = 0;
  s->state[0] = s->state[1] = s->state[2] = nullptr; s->sort_perm = nullptr; s->sort_key = nullptr; s->sort_bins = nullptr; s->hits = s->shq = s->rad = nullptr; s->defer_list = nullptr; s->cand = nullptr;
  s->counters = nullptr; s->counters_len = 0; s->shadow_ref = nullptr; s->film = nullptr; s->film_len = 0;
  if (!park) { free_wave_buffers(w); return; }
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (WaveBuffers& p : g_pool)
    if (p.device == w.device) {
      if (p.cap >= w.cap) { free_wave_buffers(w); return; }
      free_wave_buffers(p);
      p = w;
      return;
    }
  g_pool.push_back(w);
}

// Takes this device's parked bundle, if there is one (whatever its size: the caller grows what is too small).
void adopt_parked_buffers(BnScene* s) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (size_t k = 0; k < g_pool.size(); ++k)
    if (g_pool[k].device == s->device) {
      const WaveBuffers w = g_pool[k];
      g_pool.erase(g_pool.begin() + (long)k);
      s->cap = w.cap;
      s->state[0] = w.state[0]; s->state[1] = w.state[1]; s->state[2] = w.state[2]; s->sort_perm = w.sort_perm; s->sort_key = w.sort_key; s->sort_bins = w.sort_bins; s->hits = w.hits; s->shq = w.shq; s->rad = w.rad; s->defer_list = w.defer_list; s->cand = w.cand;
      s->counters = w.counters; s->counters_len = w.counters_len; s->shadow_ref = w.shadow_ref; s->film = w.film; s->film_len = w.film_len;
      s->arena = w.arena; s->arena_bytes = w.arena_bytes;
      return;
    }
}

constexpr int BN_ERR_NOMEM_RETRY = -100;  // internal: cudaErrorMemoryAllocation while growing the wave buffers (never leaves this file)

int ensure_wave_buffers(BnScene* s, size_t cap) {
  if (s->cap >= cap) return BN_OK;
  auto drop = [&]() {
    for (void* p : {(void*)s->state[0], (void*)s->state[1], (void*)s->state[2], (void*)s->sort_perm, (void*)s->sort_key, (void*)s->sort_bins, (void*)s->hits, (void*)s->shq, (void*)s->rad, (void*)s->defer_list, (void*)s->cand})
      if (p) cudaFree(p);
    s->cap = 0;
    s->state[0] = s->state[1] = s->state[2] = nullptr; s->sort_perm = nullptr; s->sort_key = nullptr; s->sort_bins = nullptr; s->hits = s->shq = s->rad = nullptr; s->defer_list = nullptr; s->cand = nullptr;
  };
  drop();
  const struct { void** p; size_t bytes; } want[] = {
      {(void**)&s->state[0], cap * 3 * sizeof(float4)}, {(void**)&s->state[1], cap * 3 * sizeof(float4)}, {(void**)&s->state[2], cap * 2 * sizeof(float4)}, {(void**)&s->sort_perm, cap * sizeof(uint32_t)},
      {(void**)&s->sort_key, cap * sizeof(uint16_t)}, {(void**)&s->sort_bins, 2 * kSortBins * sizeof(uint32_t)}, {(void**)&s->hits, cap * sizeof(float4)},
      {(void**)&s->shq, cap * 4 * sizeof(float4)},      {(void**)&s->rad, cap * sizeof(float4)},          {(void**)&s->defer_list, cap * sizeof(int)},
      {(void**)&s->cand, BN_CAND_PREPASS ? cap * sizeof(uint32_t) : (size_t)16}};
  for (const auto& w : want) {
    const cudaError_t e = cudaMalloc(w.p, w.bytes);
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();  // not sticky
      drop();
      bnhost::set_error("out of device memory for the wave buffers");
      return BN_ERR_NOMEM_RETRY;
    }
    if (!cuda_ok(e, "cudaMalloc(wave buffers)")) { drop(); return BN_ERR_CUDA; }
  }
  if (!cuda_ok(cudaMemset(s->sort_bins, 0, 2 * kSortBins * sizeof(uint32_t)), "cudaMemset(sort bins)")) { drop(); return BN_ERR_CUDA; }
  // k_sort_hist reads the keys eight at a time and masks what lies beyond the queue's end: give those lanes defined bytes
  if (!cuda_ok(cudaMemset(s->sort_key, 0, cap * sizeof(uint16_t)), "cudaMemset(sort keys)")) { drop(); return BN_ERR_CUDA; }
  s->cap = cap;
  return BN_OK;
}

int validate_params(const BnRenderParams* p) {
  if (!p || p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->max_depth < 0 || p->sample_begin < 0 || p->sample_end > p->spp ||
      p->sample_begin > p->sample_end || p->x0 < 0 || p->y0 < 0 || p->x1 > p->width || p->y1 > p->height || p->x0 > p->x1 || p->y0 > p->y1 ||
      (p->interleave_count > 1 && (p->interleave_index < 0 || p->interleave_index >= p->interleave_count)) ||
      p->integrator < BN_INTEGRATOR_PATH_TRACING || p->integrator > BN_INTEGRATOR_NORMAL) {
    bnhost::set_error("bn_render: invalid BnRenderParams");
    return BN_ERR_INVALID;
  }
  return BN_OK;
}

// Runs every wave of the window.  If d_radiance != nullptr the per-path radiance is
// exported instead of accumulated into the film.
int render_waves(BnScene* s, const BnRenderParams* p, float* d_film, float* d_radiance, cudaStream_t stream, BnStats* stats) {
  if (s->poisoned) { bnhost::set_error("scene is unusable after an earlier CUDA error"); return BN_ERR_CUDA; }
  if (s->d.n_light_inst == 0) { bnhost::set_error("LightSamplerBase: No light primitives found."); return BN_ERR_NO_LIGHT; }
  BN_CUDA(cudaSetDevice(s->device));
  const int rw = p->x1 - p->x0, rh = p->y1 - p->y0, ns = p->sample_end - p->sample_begin;
  const int il_count = p->interleave_count > 1 ? p->interleave_count : 1, il_index = il_count > 1 ? p->interleave_index : 0;
  const int tile_rows = (rh + 15) / 16;
  const int owned_rows = tile_rows > il_index ? (tile_rows - il_index + il_count - 1) / il_count : 0;
  const int nbx = (rw + 7) / 8, nby = il_count > 1 ? owned_rows * 4 : (rh + 3) / 4;
  const long long total_blocks = (long long)nbx * nby;
  uint64_t launches = 0;
  if (!s->ev_begin) BN_CUDA(cudaEventCreate(&s->ev_begin));  // owned by the scene: nothing to leak on the error returns below
  if (!s->ev_end) BN_CUDA(cudaEventCreate(&s->ev_end));
  const cudaEvent_t ev0 = s->ev_begin, ev1 = s->ev_end;
  if (d_film) BN_CUDA(cudaMemsetAsync(d_film, 0, sizeof(float) * 3 * (size_t)p->width * p->height, stream));
  BN_CUDA(cudaEventRecord(ev0, stream));
  uint64_t n_paths = 0;
  std::vector<int> h_counters;
  // bounces per path: PathTracing maxDepth | Direct: the hit and one BSDF-sampled ray | Normal: the hit
  const int D_bounces = p->integrator == BN_INTEGRATOR_DIRECT ? 2 : (p->integrator == BN_INTEGRATOR_NORMAL ? 1 : p->max_depth);
  if (total_blocks > 0 && ns > 0 && D_bounces > 0) {
    size_t cap_target = wave_capacity_paths();
    long long blocks_per_wave = 0;
    int samples_per_wave = 0;
    for (;;) {
      blocks_per_wave = std::min<long long>(total_blocks, std::max<long long>(1, (long long)(cap_target / 32)));
      samples_per_wave = (int)std::max<long long>(1, std::min<long long>(ns, (long long)cap_target / (blocks_per_wave * 32)));
      const size_t cap = (size_t)blocks_per_wave * 32 * samples_per_wave;
      const int rc = ensure_wave_buffers(s, cap);
      if (rc == BN_OK) break;
      // a smaller or shared GPU: 64 Mi paths want 13 GB of queues — halve the wave instead of failing the render
      if (rc != BN_ERR_NOMEM_RETRY || cap_target <= ((size_t)1 << 18)) return rc == BN_ERR_NOMEM_RETRY ? BN_ERR_CUDA : rc;
      cap_target /= 2;
    }
    const long long n_block_chunks = (total_blocks + blocks_per_wave - 1) / blocks_per_wave;
    const long long n_sample_chunks = (ns + samples_per_wave - 1) / samples_per_wave;
    const long long n_waves = n_block_chunks * n_sample_chunks;
    // per wave: n_active(bounce 0..maxDepth) | n_shadow(bounce) | cursors 3 per bounce | deferred counts 2 per bounce
    const int D = D_bounces;
    // Every counter sits in its own 256-B slot: the kernels of one bounce hammer three or four of them
    // with one atomic per warp per 32 paths, and atomics that share a cache line are serialised by
    // the L2 slice that owns the line (BN_COUNTER_STRIDE ints apart = different lines / slices).
    constexpr size_t CS = BN_COUNTER_STRIDE;
    const size_t per_wave = ((size_t)(D + 1) + D + 3 * (size_t)D + 2 * (size_t)D + 2 * (size_t)D) * CS;  // ... + finished-CTA tickets 2 per bounce
    const size_t need = per_wave * (size_t)n_waves;
    if (s->counters_len < need) {
      if (s->counters) cudaFree(s->counters);
      s->counters = nullptr; s->counters_len = 0;
      BN_CUDA(cudaMalloc((void**)&s->counters, need * sizeof(int)));
      s->counters_len = need;
    }
    if (!s->shadow_ref) BN_CUDA(cudaMalloc((void**)&s->shadow_ref, sizeof(unsigned long long)));
    BN_CUDA(cudaMemsetAsync(s->counters, 0, need * sizeof(int), stream));
    BN_CUDA(cudaMemsetAsync(s->shadow_ref, 0, sizeof(unsigned long long), stream));
    const int grid = s->num_sms * 8;
    const int tgrid = s->num_sms * BN_TRAV_GRID_MULT;
    DScene dsc = s->d;
    if (p->flags & BN_RENDER_FORCE_EXACT) dsc.all_finite = 0u;  // every ray is deferred to the exact fix-up kernel
    // paths ordered before every extend launch but the first (ray_sort.cuh); BN_SORT=0 switches it off (A/B, parity tests)
    const bool order_paths = !(std::getenv("BN_SORT") && std::atoi(std::getenv("BN_SORT")) == 0);
    const int order_from = std::getenv("BN_SORT_FROM") ? std::max(1, std::atoi(std::getenv("BN_SORT_FROM"))) : 1;  // first ordered bounce (primary rays are coherent as generated)
    const bool separate_fixup = !BN_INKERNEL_DRAIN || (p->flags & BN_RENDER_FORCE_EXACT) != 0 || std::getenv("BN_SEPARATE_FIXUP") != nullptr;
    // BN_RENDER_PROFILE: bracket every launch with events on the launching stream
    const bool profile = (p->flags & BN_RENDER_PROFILE) != 0 && stats != nullptr;
    std::vector<int> ev_class;  // 0 extend, 1 shade, 2 shadow, 3 other
    size_t ev_used = 0;
    auto prof_begin = [&](int cls) {
      if (!profile) return;
      if (s->events.size() < ev_used + 2) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        s->events.push_back(a);
        s->events.push_back(b);
      }
      ev_class.push_back(cls);
      cudaEventRecord(s->events[ev_used], stream);
    };
    auto prof_end = [&]() {
      if (!profile) return;
      cudaEventRecord(s->events[ev_used + 1], stream);
      ev_used += 2;
    };
    long long wave = 0;
    for (long long bc = 0; bc < n_block_chunks; ++bc) {
      for (long long scn = 0; scn < n_sample_chunks; ++scn, ++wave) {
        WaveParams wp{};
        wp.width = p->width; wp.height = p->height; wp.spp = p->spp; wp.max_depth = D; wp.rr_depth = p->rr_depth;
        wp.integrator = p->integrator;
        wp.frame_id = p->frame_id; wp.x0 = p->x0; wp.y0 = p->y0; wp.x1 = p->x1; wp.y1 = p->y1; wp.nbx = nbx;
        wp.block_begin = (int)(bc * blocks_per_wave);
        wp.n_blocks = (int)std::min<long long>(blocks_per_wave, total_blocks - bc * blocks_per_wave);
        wp.sample_begin = p->sample_begin + (int)(scn * samples_per_wave);
        wp.n_samples = std::min(samples_per_wave, p->sample_end - wp.sample_begin);
        wp.flags = p->flags;
        wp.inv_spp = 1.0f / (float)p->spp;  // MathF.ReciprocalEstimate restated as IEEE 1/x
        wp.il_count = il_count; wp.il_index = il_index;
        int* base = s->counters + per_wave * (size_t)wave;
        int* n_active = base;                               // [b] at n_active + b * CS, and so on
        int* n_shadow = n_active + CS + 1;  // [b]: the int right after n_active[b + 1] — k_shade appends to both queues with one 64-bit atomic
        int* cursors = base + (size_t)((D + 1) + D) * CS;
        int* n_defer = cursors + (size_t)(3 * D) * CS;
        int* n_done = n_defer + (size_t)(2 * D) * CS;
        const size_t cp = s->cap;
        float4* A = s->state[0];
        float4* B = s->state[1];
        prof_begin(3);
        k_raygen<<<grid, kBlock, 0, stream>>>(s->d, wp, A, A + cp, A + 2 * cp, s->rad, n_active);
        prof_end();
        ++launches;
        float4* const T = s->state[2];  // the bounce's paths in its ordering (written by extend's refills, read by shade)
        for (int b = 0; b < D; ++b) {
          const uint32_t* perm = nullptr;
          if (order_paths && b >= order_from) {
            // the paths that survived bounce b-1, ranked by (direction octant, origin cell): ray_sort.cuh
            prof_begin(3);
            k_sort_hist<<<s->num_sms * 4, kSortThreads, 0, stream>>>(s->sort_key, n_active + b * CS, s->sort_bins);
            k_sort_scan<<<1, kSortScanThreads, 0, stream>>>(s->sort_bins, s->sort_bins + kSortBins);
            k_sort_rank<<<s->num_sms * 4, kSortThreads, 0, stream>>>(s->sort_key, n_active + b * CS, s->sort_bins + kSortBins, s->sort_perm);
            prof_end();
            launches += 3;
            perm = s->sort_perm;
          }
          // ordered bounce: extend reads planes 0, 1 of A through perm and leaves them in order in T, where shade reads them;
          // plane 2 stays in A and shade reads it through perm
          const float4* const S = perm ? T : A;
          prof_begin(0);
          const ExtendIO eio{A, A + cp, s->hits, n_active + b * CS, cursors + (3 * b) * CS, DeferList{n_defer + (2 * b) * CS, s->defer_list}, s->cand, perm, T, T + cp};
          // the last CTA to finish drains the deferred rays; with BN_RENDER_FORCE_EXACT every ray is deferred and a
          // full-width fix-up launch does the work instead
          launch_traverse<false>(tgrid, stream, dsc, eio, separate_fixup ? nullptr : n_done + (2 * b) * CS);
          if (separate_fixup) k_traverse_fixup<false, ExtendIO><<<grid, kBlock, 0, stream>>>(dsc, eio);
          prof_end();
          prof_begin(1);
          k_shade<<<s->num_sms * BN_SHADE_MIN_BLOCKS, kBlock, 0, stream>>>(s->d, wp, b, S, S + cp, A + 2 * cp, s->hits, B, B + cp, B + 2 * cp, s->shq, s->shq + cp,
                                               s->shq + 2 * cp, s->shq + 3 * cp, s->rad, n_active + b * CS, n_active + (b + 1) * CS,
                                               cursors + (3 * b + 1) * CS, s->shadow_ref, (order_paths && b + 1 < D && b + 1 >= order_from) ? s->sort_key : nullptr, perm);
          prof_end();
          prof_begin(2);
          const ShadowIO sio{s->shq, s->shq + cp, s->shq + 2 * cp, s->shq + 3 * cp, s->rad, n_shadow + b * CS, cursors + (3 * b + 2) * CS, DeferList{n_defer + (2 * b + 1) * CS, s->defer_list}, s->cand};
          launch_traverse<true>(tgrid, stream, dsc, sio, separate_fixup ? nullptr : n_done + (2 * b + 1) * CS);
          if (separate_fixup) k_traverse_fixup<true, ShadowIO><<<grid, kBlock, 0, stream>>>(dsc, sio);
          prof_end();
          launches += (separate_fixup ? 5 : 3) + ((BN_CAND_PREPASS && dsc.flat_tlas != nullptr) ? 2 : 0);
          std::swap(A, B);
        }
        prof_begin(3);
        if (d_radiance) k_export_radiance<<<grid, kBlock, 0, stream>>>(wp, s->rad, d_radiance, p->sample_begin);
        else k_accumulate<<<grid, kBlock, 0, stream>>>(wp, s->rad, d_film);
        prof_end();
        ++launches;
      }
    }
    {
      const cudaError_t le = cudaGetLastError();  // a failed launch leaves the queues half-built: the scene is unusable from here
      if (le != cudaSuccess) { s->poisoned = true; cuda_ok(le, "kernel launch"); return BN_ERR_CUDA; }
    }
    h_counters.resize(need);
    BN_CUDA(cudaEventRecord(ev1, stream));
    BN_CUDA(cudaMemcpyAsync(h_counters.data(), s->counters, need * sizeof(int), cudaMemcpyDeviceToHost, stream));
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { s->poisoned = true; cuda_ok(e, "render waves"); return BN_ERR_CUDA; }
    if (std::getenv("BN_DEBUG_COUNTS")) {  // debug aid: rays per launch (extend / shadow of every bounce of every wave), for the profiles
      for (long long w = 0; w < n_waves; ++w) {
        const int* base = h_counters.data() + per_wave * (size_t)w;
        std::fprintf(stderr, "bn_counts wave %lld:", w);
        for (int b = 0; b < D; ++b) std::fprintf(stderr, " b%d extend=%d shadow=%d", b, base[(size_t)b * CS], base[(size_t)(b + 1) * CS + 1]);
        std::fprintf(stderr, "\n");
      }
    }
    if (stats) {
      uint64_t ext = 0, sh = 0;
      for (long long w = 0; w < n_waves; ++w) {
        const int* base = h_counters.data() + per_wave * (size_t)w;
        n_paths += (uint64_t)base[0];
        for (int b = 0; b < D; ++b) { ext += (uint64_t)base[(size_t)b * CS]; sh += (uint64_t)base[(size_t)(b + 1) * CS + 1]; }
      }
      unsigned long long ref = 0;
      BN_CUDA(cudaMemcpy(&ref, s->shadow_ref, sizeof ref, cudaMemcpyDeviceToHost));
      stats->paths = n_paths; stats->extend_rays = ext; stats->shadow_rays = sh; stats->shadow_rays_ref = ref;
      stats->extend_ms = stats->shade_ms = stats->shadow_ms = stats->other_ms = 0.0;
      for (size_t k = 0; k < ev_class.size(); ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->events[2 * k], s->events[2 * k + 1]);
        double* dst = ev_class[k] == 0 ? &stats->extend_ms : (ev_class[k] == 1 ? &stats->shade_ms : (ev_class[k] == 2 ? &stats->shadow_ms : &stats->other_ms));
        *dst += ms;
      }
    }
  } else {
    BN_CUDA(cudaEventRecord(ev1, stream));
    BN_CUDA(cudaStreamSynchronize(stream));
    if (stats) { stats->paths = 0; stats->extend_rays = 0; stats->shadow_rays = 0; stats->shadow_rays_ref = 0; stats->extend_ms = stats->shade_ms = stats->shadow_ms = stats->other_ms = 0.0; }
  }
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    stats->gpu_ms = ms;
    stats->kernel_launches = launches;
  }
  return BN_OK;
}

}  // namespace

extern "C" {

#ifdef BN_TRAV_STATS
// debug builds only: [closest: cntN cntT cntE cntAll sumN sumT sumE sumAll | any: same]
__attribute__((visibility("default"))) int bn_debug_trav_stats(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, bn::g_trav_stats, sizeof(unsigned long long) * 24);
  if (reset) { unsigned long long z[24] = {0}; cudaMemcpyToSymbol(bn::g_trav_stats, z, sizeof z); }
  return 0;
}
#endif

int bn_measure_l2_read_gbs(int device, uint64_t bytes, int iters, double* gbs) {
  if (!gbs || bytes < 4096 || iters < 1) { bnhost::set_error("bn_measure_l2_read_gbs: bad argument"); return BN_ERR_INVALID; }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); bnhost::set_error("no CUDA device"); return BN_ERR_NO_DEVICE; }
  cudaSetDevice(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  uint4* buf = nullptr;
  unsigned* sink = nullptr;
  if (cudaMalloc((void**)&buf, bytes) != cudaSuccess || cudaMalloc((void**)&sink, 4) != cudaSuccess) { cudaGetLastError(); bnhost::set_error("cudaMalloc failed"); return BN_ERR_CUDA; }
  cudaMemset(buf, 1, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t n_vec = bytes / 16;
  const int grid = prop.multiProcessorCount * 8;
  bn::k_l2_sweep<<<grid, 256>>>(buf, n_vec, 2, sink);  // warm the L2
  cudaEventRecord(e0);
  bn::k_l2_sweep<<<grid, 256>>>(buf, n_vec, iters, sink);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  if (e != cudaSuccess || ms <= 0.f) { bnhost::set_error("L2 probe failed"); return BN_ERR_CUDA; }
  *gbs = (double)n_vec * 16.0 * iters / (ms * 1e-3) / 1e9;
  return BN_OK;
}

// parity-test entry (declared in the header): the ordering's three kernels on a key array of the caller
int bn_debug_order_keys(int device, const uint16_t* keys, uint32_t n, uint32_t* perm) {
  if ((n && (!keys || !perm)) || n > (1u << 30)) { bnhost::set_error("bn_debug_order_keys: bad argument"); return BN_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); bnhost::set_error("no CUDA device"); return BN_ERR_NO_DEVICE; }
  if (n == 0) return BN_OK;
  BN_CUDA(cudaSetDevice(device));
  int sms = 0;
  BN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const size_t padded = ((size_t)n + 31) & ~(size_t)31;  // k_sort_hist reads the keys eight at a time
  uint16_t* d_key = nullptr;
  uint32_t *d_perm = nullptr, *d_bins = nullptr;
  int* d_n = nullptr;
  int rc = BN_OK;
  const int n_int = (int)n;
  if (!cuda_ok(cudaMalloc((void**)&d_key, padded * sizeof(uint16_t)), "cudaMalloc") || !cuda_ok(cudaMalloc((void**)&d_perm, (size_t)n * sizeof(uint32_t)), "cudaMalloc") ||
      !cuda_ok(cudaMalloc((void**)&d_bins, 2 * kSortBins * sizeof(uint32_t)), "cudaMalloc") || !cuda_ok(cudaMalloc((void**)&d_n, sizeof(int)), "cudaMalloc") ||
      !cuda_ok(cudaMemset(d_key, 0, padded * sizeof(uint16_t)), "cudaMemset") || !cuda_ok(cudaMemset(d_bins, 0, 2 * kSortBins * sizeof(uint32_t)), "cudaMemset") ||
      !cuda_ok(cudaMemset(d_perm, 0xFF, (size_t)n * sizeof(uint32_t)), "cudaMemset") ||
      !cuda_ok(cudaMemcpy(d_key, keys, (size_t)n * sizeof(uint16_t), cudaMemcpyHostToDevice), "copy keys") ||
      !cuda_ok(cudaMemcpy(d_n, &n_int, sizeof(int), cudaMemcpyHostToDevice), "copy n")) {
    rc = BN_ERR_CUDA;
  } else {
    k_sort_hist<<<sms * 4, kSortThreads>>>(d_key, d_n, d_bins);
    k_sort_scan<<<1, kSortScanThreads>>>(d_bins, d_bins + kSortBins);
    k_sort_rank<<<sms * 4, kSortThreads>>>(d_key, d_n, d_bins + kSortBins, d_perm);
    if (!cuda_ok(cudaGetLastError(), "ordering kernels") || !cuda_ok(cudaMemcpy(perm, d_perm, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost), "copy perm")) rc = BN_ERR_CUDA;
  }
  for (void* p : {(void*)d_key, (void*)d_perm, (void*)d_bins, (void*)d_n})
    if (p) cudaFree(p);
  return rc;
}

int bn_release_cached_buffers(int device) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (size_t k = 0; k < g_pool.size();) {
    if (device < 0 || g_pool[k].device == device) {
      cudaSetDevice(g_pool[k].device);
      free_wave_buffers(g_pool[k]);
      g_pool.erase(g_pool.begin() + (long)k);
    } else {
      ++k;
    }
  }
  return BN_OK;
}

int bn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int bn_scene_create(const BnSceneDesc* desc, int device, BnScene** out) {
  if (!desc || !out) { bnhost::set_error("bn_scene_create: NULL argument"); return BN_ERR_INVALID; }
  *out = nullptr;
  std::shared_ptr<const bnint::Staged> st;
  int rc = bnint::stage_scene(desc, st);
  if (rc != BN_OK) return rc;
  return bnint::scene_from_staged(*st, desc->camera, device, out);
}

}  // extern "C"

// Flattening (host) and upload (device) are separate steps so that a multi-device render flattens ONCE (multi.cu), and so
// that a host which re-creates its scene every frame from unchanged geometry (the F# binding does: INTEGRATION.md) does not
// re-derive the device layout: the last staged scene is kept and reused when the new description's arrays are byte-identical
// (exact comparison, ~0.3 ms for C2's 3.4 MB against ~8 ms of conversion; BN_NO_SCENE_CACHE switches it off).  The upload
// itself — the step's host -> device copy — happens every time.
int bnint::stage_scene(const BnSceneDesc* desc, std::shared_ptr<const bnint::Staged>& out) {
  if (bn_device_count() <= 0) { bnhost::set_error("no CUDA device available (the hot path has no CPU fallback)"); return BN_ERR_NO_DEVICE; }
  const uint32_t env_flags = (std::getenv("BN_BINARY_NODES") ? 1u : 0u) | (std::getenv("BN_NO_FLAT_TLAS") ? 2u : 0u);
  const bool cache = std::getenv("BN_NO_SCENE_CACHE") == nullptr;
  if (cache) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (g_last_staged && scene_key_matches(*desc, env_flags, g_last_staged->key)) { out = g_last_staged; return BN_OK; }
  }
  bnconv::ConvertedScene cs;
  std::string err;
  if (!bnconv::convert_scene(*desc, cs, err)) { bnhost::set_error(err); return BN_ERR_INVALID; }
  if (env_flags & 1u) bnconv::use_binary_nodes(cs);  // A/B switch: the binary-node fast path
  auto st = std::make_shared<bnint::Staged>();
  std::vector<unsigned char>& im = st->image;
  im.reserve(cs.nodes.size() * sizeof(bn::GNode) + cs.wide.size() * sizeof(bn::GWide) + cs.tris.size() * sizeof(bn::GTri) + cs.alias.size() * sizeof(bn::GAlias) +
             cs.inst_trav.size() * 512 + (64 << 10));
  st->at.nodes = stage_add(im, cs.nodes); st->at.inst_trav = stage_add(im, cs.inst_trav); st->at.inst_head = stage_add(im, cs.inst_head);
  st->at.inst_w2o = stage_add(im, cs.inst_w2o); st->at.inst_o2w = stage_add(im, cs.inst_o2w); st->at.meshes = stage_add(im, cs.meshes);
  st->at.tris = stage_add(im, cs.tris); st->at.alias = stage_add(im, cs.alias); st->at.sphere_radii = stage_add(im, cs.sphere_radii);
  st->at.materials = stage_add(im, cs.materials); st->at.lights = stage_add(im, cs.lights); st->at.light_inst = stage_add(im, cs.light_inst);
  st->at.flat_tlas = (!cs.flat_tlas.empty() && !(env_flags & 2u)) ? stage_add(im, cs.flat_tlas) : kNoSlice;
  st->at.wide = !cs.wide.empty() ? stage_add(im, cs.wide) : kNoSlice;
  st->tlas = cs.tlas; st->tlas_wroot = cs.tlas_wroot;
  st->n_inst = (uint32_t)cs.inst_head.size(); st->n_light_inst = (uint32_t)cs.light_inst.size(); st->all_finite = cs.all_finite ? 1u : 0u;
  if (cache) {
    scene_key(*desc, env_flags, st->key);
    // page-locked: the upload is then a plain DMA (the image lives as long as the cache entry or a scene being created)
    if (cudaHostRegister(im.data(), im.size(), cudaHostRegisterPortable) == cudaSuccess) st->pinned = true;
    else cudaGetLastError();
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    g_last_staged = st;
  }
  out = st;
  return BN_OK;
}

int bnint::render_on_stream(BnScene* s, const BnRenderParams* p, float* d_film, cudaStream_t stream, BnStats* stats) {
  int rc = validate_params(p);
  if (rc != BN_OK) return rc;
  return render_waves(s, p, d_film, nullptr, stream, stats);
}

int bnint::scene_from_staged(const bnint::Staged& st, const BnCamera& c, int device, BnScene** out) {
  *out = nullptr;
  const int ndev = bn_device_count();
  if (ndev <= 0) { bnhost::set_error("no CUDA device available (the hot path has no CPU fallback)"); return BN_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) { bnhost::set_error("bn_scene_create: device ordinal out of range"); return BN_ERR_INVALID; }
  if (c.type > BN_CAM_THIN_LENS) { bnhost::set_error("unknown camera type"); return BN_ERR_INVALID; }
  BN_CUDA(cudaSetDevice(device));
  int cc_major = 0, sm_count = 0;  // (cudaGetDeviceProperties costs milliseconds per call; two attributes do not)
  BN_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  BN_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  if (cc_major < 10) { bnhost::set_error("device is not sm_100-class (this library is built for sm_100a only)"); return BN_ERR_NO_DEVICE; }
  auto* s = new BnScene();
  s->device = device;
  s->num_sms = sm_count;
  DScene& d = s->d;
  // every scene array in ONE device allocation, filled by ONE host -> device copy
  adopt_parked_buffers(s);  // a warm device hands over the previous scene's allocation (and its wave buffers)
  if (s->arena_bytes < st.image.size()) {
    if (s->arena) cudaFree(s->arena);
    s->arena = nullptr; s->arena_bytes = 0;
    if (!cuda_ok(cudaMalloc(&s->arena, st.image.size()), "cudaMalloc(scene arena)")) { bn_scene_destroy(s); return BN_ERR_CUDA; }
    s->arena_bytes = st.image.size();
  }
  if (!cuda_ok(cudaMemcpy(s->arena, st.image.data(), st.image.size(), cudaMemcpyHostToDevice), "scene upload")) { bn_scene_destroy(s); return BN_ERR_CUDA; }
  const unsigned char* base = static_cast<const unsigned char*>(s->arena);
  auto at = [&](size_t off) -> const void* { return off == kNoSlice ? nullptr : base + off; };
  d.nodes = static_cast<const GNode*>(at(st.at.nodes)); d.inst_trav = static_cast<const GInstTrav*>(at(st.at.inst_trav));
  d.inst_head = static_cast<const GInstHead*>(at(st.at.inst_head)); d.inst_w2o = static_cast<const GMat43*>(at(st.at.inst_w2o));
  d.inst_o2w = static_cast<const GMat43*>(at(st.at.inst_o2w)); d.meshes = static_cast<const GMesh*>(at(st.at.meshes));
  d.tris = static_cast<const GTri*>(at(st.at.tris)); d.alias = static_cast<const GAlias*>(at(st.at.alias));
  d.sphere_radii = static_cast<const float*>(at(st.at.sphere_radii)); d.materials = static_cast<const GMaterial*>(at(st.at.materials));
  d.lights = static_cast<const GLight*>(at(st.at.lights)); d.light_inst = static_cast<const uint32_t*>(at(st.at.light_inst));
  d.flat_tlas = static_cast<const GFlatInst*>(at(st.at.flat_tlas));
  d.wide = static_cast<const GWide*>(at(st.at.wide));
  d.tlas_wroot = st.tlas_wroot;
  // a TLAS of thousands of instances (C4) likes its closest-hit rays refilled earlier: 16 idle lanes instead of 20 gives C4 +1.4 %,
  // and costs the scenes of a few dozen instances 1 % (profiles/r02_ab_session35_*.log)
  d.refill_min = st.n_inst >= 256u ? 16u : 0u;
  d.tlas = st.tlas;
  d.n_inst = st.n_inst;
  d.n_light_inst = st.n_light_inst;
  d.all_finite = st.all_finite;
  bnconv::convert_camera(c, d.cam);
  d.sort_grid.cell_major = d.flat_tlas != nullptr ? 1u : 0u;
  d.sort_grid.pad = 0u;
  for (int a = 0; a < 3; ++a) {  // ray_sort.cuh: 2^m cells per axis over the TLAS root box (any finite grid is valid: the key only orders work)
    const float lo = st.tlas.bmin[a], ext = st.tlas.bmax[a] - st.tlas.bmin[a];
    const bool ok = std::isfinite(lo) && std::isfinite(ext) && ext > 0.f;
    d.sort_grid.lo[a] = ok ? lo : 0.f;
    d.sort_grid.scale[a] = ok ? (float)(1 << kSortMBits) / ext : 0.f;
  }
  *out = s;
  return BN_OK;
}

// The scene's own device film (grown on demand, parked with the other buffers when the scene dies).
int bnint::scene_film(BnScene* s, size_t len, float** out) {
  BN_CUDA(cudaSetDevice(s->device));
  if (s->film_len < len) {
    if (s->film) cudaFree(s->film);
    s->film = nullptr; s->film_len = 0;
    BN_CUDA(cudaMalloc((void**)&s->film, len * sizeof(float)));
    s->film_len = len;
  }
  *out = s->film;
  return BN_OK;
}

// ---- the traversal kernels for other wavefronts (mlt.cu): same queues, same kernels, fix-up launch included ------------
int bnint::ensure_wave(BnScene* s, size_t cap) { return ensure_wave_buffers(s, cap); }

void bnint::launch_extend(BnScene* s, cudaStream_t stream, const float4* s0, const float4* s1, float4* hits, const int* n_ptr, int* cursor, int* n_defer) {
  const ExtendIO io{s0, s1, hits, n_ptr, cursor, DeferList{n_defer, s->defer_list}, s->cand, nullptr, nullptr, nullptr};
  launch_traverse<false>(s->num_sms * BN_TRAV_GRID_MULT, stream, s->d, io, nullptr);
  k_traverse_fixup<false, ExtendIO><<<s->num_sms, kBlock, 0, stream>>>(s->d, io);
}

void bnint::launch_shadow(BnScene* s, cudaStream_t stream, const float4* q0, const float4* q1, const float4* q2, const float4* q3, float4* rad, const int* n_ptr,
                          int* cursor, int* n_defer) {
  const ShadowIO io{q0, q1, q2, q3, rad, n_ptr, cursor, DeferList{n_defer, s->defer_list}, s->cand};
  launch_traverse<true>(s->num_sms * BN_TRAV_GRID_MULT, stream, s->d, io, nullptr);
  k_traverse_fixup<true, ShadowIO><<<s->num_sms, kBlock, 0, stream>>>(s->d, io);
}

extern "C" {

void bn_scene_destroy(BnScene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (cudaEvent_t e : s->events) cudaEventDestroy(e);
  if (s->ev_begin) cudaEventDestroy(s->ev_begin);
  if (s->ev_end) cudaEventDestroy(s->ev_end);
  if (s->trace_ctr) cudaFree(s->trace_ctr);
  if (s->trace_dlist) cudaFree(s->trace_dlist);
  release_wave_buffers(s);  // parks the wave, counter and film buffers for the next scene on this device
  for (void* p : {(void*)s->mlt_f, (void*)s->mlt_i, (void*)s->mlt_w, (void*)s->mlt_cnt, (void*)s->mlt_acc, (void*)s->mlt_wave_i, (void*)s->mlt_wave_f, (void*)s->mlt_counters})
    if (p) cudaFree(p);
  if (s->mlt_w_host) cudaFreeHost(s->mlt_w_host);
  delete s;
}

int bn_render_device(BnScene* s, const BnRenderParams* p, void* d_film, void* stream, BnStats* stats) {
  if (!s || !d_film) { bnhost::set_error("bn_render_device: NULL argument"); return BN_ERR_INVALID; }
  int rc = validate_params(p);
  if (rc != BN_OK) return rc;
  return render_waves(s, p, static_cast<float*>(d_film), nullptr, static_cast<cudaStream_t>(stream), stats);
}

int bn_render(BnScene* s, const BnRenderParams* p, float* film, BnStats* stats) {
  if (!s || !film) { bnhost::set_error("bn_render: NULL argument"); return BN_ERR_INVALID; }
  int rc = validate_params(p);
  if (rc != BN_OK) return rc;
  const size_t len = (size_t)p->width * p->height * 3;
  float* d_film = nullptr;
  if ((rc = bnint::scene_film(s, len, &d_film)) != BN_OK) return rc;
  rc = render_waves(s, p, s->film, nullptr, nullptr, stats);
  if (rc != BN_OK) return rc;
  BN_CUDA(cudaMemcpy(film, s->film, len * sizeof(float), cudaMemcpyDeviceToHost));
  return BN_OK;
}

int bn_render_radiance(BnScene* s, const BnRenderParams* p, float* radiance) {
  if (!s || !radiance) { bnhost::set_error("bn_render_radiance: NULL argument"); return BN_ERR_INVALID; }
  int rc = validate_params(p);
  if (rc != BN_OK) return rc;
  BN_CUDA(cudaSetDevice(s->device));
  const size_t len = (size_t)(p->x1 - p->x0) * (p->y1 - p->y0) * (p->sample_end - p->sample_begin) * 3;
  if (len == 0) return BN_OK;
  float* d = nullptr;  // the scene's own film buffer doubles as the export buffer (grown on demand, parked with the scene's buffers)
  if ((rc = bnint::scene_film(s, len, &d)) != BN_OK) return rc;
  BN_CUDA(cudaMemset(d, 0, len * sizeof(float)));
  rc = render_waves(s, p, nullptr, d, nullptr, nullptr);
  if (rc == BN_OK && !cuda_ok(cudaMemcpy(radiance, d, len * sizeof(float), cudaMemcpyDeviceToHost), "copy radiance")) rc = BN_ERR_CUDA;
  return rc;
}

int bn_film_to_rgba8_device(BnScene* s, const void* d_film, int32_t w, int32_t h, int32_t tone, void* d_rgba8, void* stream_v) {
  if (!s || !d_film || !d_rgba8 || w <= 0 || h <= 0 || tone < 0 || tone > 2) { bnhost::set_error("bn_film_to_rgba8_device: bad arguments"); return BN_ERR_INVALID; }
  BN_CUDA(cudaSetDevice(s->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  k_film_to_rgba8<<<s->num_sms * 4, 256, 0, stream>>>(static_cast<const float*>(d_film), (size_t)w * h, tone, static_cast<uchar4*>(d_rgba8));
  BN_CUDA(cudaGetLastError());
  BN_CUDA(cudaStreamSynchronize(stream));
  return BN_OK;
}

int bn_trace_device(BnScene* s, const void* d_rays, uint64_t n, int any_hit, void* d_hits, void* stream_v, float* ms) {
  if (!s || (n && (!d_rays || !d_hits))) { bnhost::set_error("bn_trace_device: NULL argument"); return BN_ERR_INVALID; }
  if (s->poisoned) { bnhost::set_error("scene is unusable after an earlier CUDA error"); return BN_ERR_CUDA; }
  BN_CUDA(cudaSetDevice(s->device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const uint64_t chunk = 1ull << 28;
  const int n_chunks = (int)((n + chunk - 1) / chunk);
  // per chunk: cursor, deferred count; one deferred list shared by the (stream-ordered) chunks
  // scratch kept with the scene and grown on demand (no driver allocation in a warm call, nothing to leak on an error return)
  const size_t want_ctr = 2 * (size_t)std::max(n_chunks, 1), want_dlist = (size_t)std::max<uint64_t>(std::min<uint64_t>(chunk, n), 1);
  if (s->trace_ctr_len < want_ctr) {
    if (s->trace_ctr) cudaFree(s->trace_ctr);
    s->trace_ctr = nullptr; s->trace_ctr_len = 0;
    BN_CUDA(cudaMalloc((void**)&s->trace_ctr, sizeof(int) * want_ctr));
    s->trace_ctr_len = want_ctr;
  }
  if (s->trace_dlist_len < want_dlist) {
    if (s->trace_dlist) cudaFree(s->trace_dlist);
    s->trace_dlist = nullptr; s->trace_dlist_len = 0;
    BN_CUDA(cudaMalloc((void**)&s->trace_dlist, sizeof(int) * want_dlist));
    s->trace_dlist_len = want_dlist;
  }
  int* const ctr = s->trace_ctr;
  int* const dlist = s->trace_dlist;
  BN_CUDA(cudaMemsetAsync(ctr, 0, sizeof(int) * want_ctr, stream));
  if (!s->ev_begin) BN_CUDA(cudaEventCreate(&s->ev_begin));
  if (!s->ev_end) BN_CUDA(cudaEventCreate(&s->ev_end));
  const cudaEvent_t e0 = s->ev_begin, e1 = s->ev_end;
  BN_CUDA(cudaEventRecord(e0, stream));
  const int grid = s->num_sms * 8;
  for (int c = 0; c < n_chunks; ++c) {
    const BnRay* r = static_cast<const BnRay*>(d_rays) + (uint64_t)c * chunk;
    BnHit* h = static_cast<BnHit*>(d_hits) + (uint64_t)c * chunk;
    const int m = (int)std::min<uint64_t>(chunk, n - (uint64_t)c * chunk);
    if (any_hit) {
      const TraceIO<true> io{s->d, r, h, m, ctr + 2 * c, DeferList{ctr + 2 * c + 1, dlist}};
      launch_traverse<true>(grid, stream, s->d, io, nullptr);
      k_traverse_fixup<true, TraceIO<true>><<<s->num_sms, kBlock, 0, stream>>>(s->d, io);
    } else {
      const TraceIO<false> io{s->d, r, h, m, ctr + 2 * c, DeferList{ctr + 2 * c + 1, dlist}};
      launch_traverse<false>(grid, stream, s->d, io, nullptr);
      k_traverse_fixup<false, TraceIO<false>><<<s->num_sms, kBlock, 0, stream>>>(s->d, io);
    }
  }
  BN_CUDA(cudaEventRecord(e1, stream));
  cudaError_t e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  float t = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
  if (e != cudaSuccess) { s->poisoned = true; cuda_ok(e, "bn_trace"); return BN_ERR_CUDA; }
  if (ms) *ms = t;
  return BN_OK;
}

int bn_trace(BnScene* s, const BnRay* rays, uint64_t n, int any_hit, BnHit* hits) {
  if (!s || (n && (!rays || !hits))) { bnhost::set_error("bn_trace: NULL argument"); return BN_ERR_INVALID; }
  if (n == 0) return BN_OK;
  BN_CUDA(cudaSetDevice(s->device));
  BnRay* d_rays = nullptr;
  BnHit* d_hits = nullptr;
  BN_CUDA(cudaMalloc((void**)&d_rays, n * sizeof(BnRay)));
  if (!cuda_ok(cudaMalloc((void**)&d_hits, n * sizeof(BnHit)), "cudaMalloc hits")) { cudaFree(d_rays); return BN_ERR_CUDA; }
  int rc = BN_OK;
  if (!cuda_ok(cudaMemcpy(d_rays, rays, n * sizeof(BnRay), cudaMemcpyHostToDevice), "copy rays")) rc = BN_ERR_CUDA;
  if (rc == BN_OK) rc = bn_trace_device(s, d_rays, n, any_hit, d_hits, nullptr, nullptr);
  if (rc == BN_OK && !cuda_ok(cudaMemcpy(hits, d_hits, n * sizeof(BnHit), cudaMemcpyDeviceToHost), "copy hits")) rc = BN_ERR_CUDA;
  cudaFree(d_rays);
  cudaFree(d_hits);
  return rc;
}

}  // extern "C"
