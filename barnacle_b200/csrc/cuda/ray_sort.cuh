// Ordering of the live paths between bounces (no counterpart in the reference: its paths are independent loop iterations,
// Integrator.fs:34-44 — the order in which a wavefront processes them is free and the film does not depend on it: every
// path owns rad[pathId], and the additions to it keep their order).
//
// Why: after the first diffuse bounce the compacted queue holds rays whose origins and directions have nothing to do with
// their neighbours', and the traversal loop (traverse.cuh) runs a warp's 32 rays in lock step — 13.7 of 32 lanes per
// instruction and 143 warp instructions per ray on C2's deep bounces against 16.8 / 94 on the coherent primary rays.
// Measured on the B200 with a library sort standing in (profiles/r02_ab_session12_*.log): paths ordered by (direction
// octant, Morton cell of the origin) before every extend launch: extend -21 %, shade -23 %, shadow -5 % on C2 (C3: -20 / -21 /
// -11 %); 64 cells (2 bits per axis) give 90 % of what 32 768 cells give, the octant alone a quarter.
//
// How: a counting sort on a 12-bit key (8 octants x 8 x 8 x 8 cells of the scene's bounding box), hand-written because the
// problem is much smaller than a general sort's — the producer (k_shade) holds origin and direction when it appends a path,
// so the key costs a 2-byte store; 4096 bins fit a CTA's shared memory, so ranking needs no radix passes:
//   k_sort_hist   keys -> global histogram (shared-memory histogram per CTA, one flush)
//   k_sort_scan   4096 counts -> bin cursors (one CTA)
//   k_sort_rank   per tile of 4096 paths: shared-memory ranks, ONE global atomic per non-empty bin of the tile claims the
//                 tile's slice of that bin; writes perm[ordered slot] = queue index
// The sort proper moves no path: 2 + 2 + 2 B of keys and 4 B of perm per path, 0.2 ms for 33 M paths.  Who pays for reading
// 48-B path records in a scattered order was the question; measured on C2 at 32 spp (63 ms per frame unordered;
// profiles/r02_ab_session13..17_*.log):
//   a copy kernel moving the state (all of the queue, or inside 256 Ki-path segments so that the writes stay in L2):
//       runs at 2.9-3.3 TB/s, 7-8 ms per frame for the 9.5 ms the order gains                              -> 61.5 ms
//   nothing moves, extend and shade both read through perm: shade 13.5 -> 15.6 ms instead of -> 11.0      -> 61.0 ms
//   extend's refill gathers all three planes and leaves them in order for shade: the gathers are L1 wavefronts, the
//       traversal kernel's scarcest resource (a 16-B gather costs a lane-cycle): extend 31.3 -> 30.2 only  -> 58.8 ms
//   extend gathers the two planes it needs anyway and leaves them in order, shade gathers the third       -> 56.8 ms  (adopted)
// (ExtendIO::load and k_shade in kernels.cu).  Ordering the shadow queue the same way loses (shadow 15.2 -> 19.4 ms: the
// connect's reads become gathers too).  Full-size frames with the ordering and what it made worth doing (refill thresholds, key
// layout per scene, swizzled node chunks, k_shade's claim pipeline): C1 +9 %, C2 +19 %, C3 +16 %, C4 +10 % (profiles/README.md).
#pragma once
#include <stdint.h>

#include "device_scene.h"
#include "vecmath.cuh"

namespace bn {

#ifndef BN_SORT_MBITS
#define BN_SORT_MBITS 3   // Morton bits per axis of the origin cell (2: 64 cells, 512 bins; 3: 512 cells, 4096 bins — C3 / C4 1.3 % faster, C1 / C2 the same)
#endif
constexpr int kSortMBits = BN_SORT_MBITS;
constexpr int kSortBins = 8 << (3 * kSortMBits);
constexpr int kSortThreads = 256;
#ifndef BN_SORT_PER_THREAD
#define BN_SORT_PER_THREAD 16
#endif
constexpr int kSortPerThread = BN_SORT_PER_THREAD;
constexpr int kSortTile = kSortThreads * kSortPerThread;  // 4096 paths: a rank fits 16 bits next to a 12-bit key
static_assert(kSortTile <= 65536, "a rank within a tile must fit 16 bits");
static_assert(kSortBins <= 4096 && kSortMBits >= 0, "the sort key must fit 12 bits");

// key = octant of the direction (the small-TLAS scan order and the child order of every node depend on it) and the
// Morton code of the origin's cell.  Any value is a valid key: a NaN origin lands in cell 0.
BN_DEV uint32_t sort_key(const SortGrid& g, const float3 p, const float3 d) {
  const int m = (1 << kSortMBits) - 1;
  const int qx = min(max(__float2int_rz((p.x - g.lo[0]) * g.scale[0]), 0), m);
  const int qy = min(max(__float2int_rz((p.y - g.lo[1]) * g.scale[1]), 0), m);
  const int qz = min(max(__float2int_rz((p.z - g.lo[2]) * g.scale[2]), 0), m);
  uint32_t code = 0;
#pragma unroll
  for (int b = 0; b < kSortMBits; ++b)
    code |= (((uint32_t)qx >> b) & 1u) << (3 * b) | (((uint32_t)qy >> b) & 1u) << (3 * b + 1) | (((uint32_t)qz >> b) & 1u) << (3 * b + 2);
  const uint32_t oct = (d.x > 0.f ? 1u : 0u) | (d.y > 0.f ? 2u : 0u) | (d.z > 0.f ? 4u : 0u);
  // octant-major: neighbours in the queue share the octant and lie in neighbouring cells; cell-major: they share the cell.
  // Measured (profiles/r02_ab_session24_*.log): cell-major C1 -2.8 %, C2 -1.8 % per frame (scenes behind a small TLAS: every
  // octant scans the same <= 16 instance boxes, a cell's rays share the BLAS regions they enter), C3 +1.3 %, C4 -0.2 % (tree
  // TLAS: the octant decides which half of every node comes first) — so the layout follows the scene (SortGrid.cell_major).
  return g.cell_major ? (code << 3) | oct : (oct << (3 * kSortMBits)) | code;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint16_t* __restrict__ key, const int* __restrict__ n_ptr, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kSortBins];
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) sh[b] = 0u;
  __syncthreads();
  const int n = *n_ptr;
  for (int i = (blockIdx.x * kSortThreads + threadIdx.x) * 8; i < n; i += gridDim.x * kSortThreads * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(key + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i + j < n) atomicAdd(&sh[(w[j >> 1] >> ((j & 1) * 16)) & (uint32_t)(kSortBins - 1)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads)
    if (sh[b]) atomicAdd(&hist[b], sh[b]);
}

// exclusive prefix sum of the histogram -> cursor (first slot of every bin); the histogram is zeroed for the next sort
constexpr int kSortScanThreads = kSortBins < 1024 ? kSortBins : 1024;
__global__ void __launch_bounds__(kSortScanThreads) k_sort_scan(uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor) {
  constexpr int per = kSortBins / kSortScanThreads;
  __shared__ uint32_t warp_sum[32];
  uint32_t local[per];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < per; ++k) {
    local[k] = hist[threadIdx.x * per + k];
    hist[threadIdx.x * per + k] = 0u;
    sum += local[k];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += v;
  }
  if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t w = threadIdx.x < kSortScanThreads / 32 ? warp_sum[threadIdx.x] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
      if (threadIdx.x >= o) wi += v;
    }
    warp_sum[threadIdx.x] = wi - w;  // exclusive
  }
  __syncthreads();
  uint32_t run = warp_sum[threadIdx.x >> 5] + incl - sum;
#pragma unroll
  for (int k = 0; k < per; ++k) {
    cursor[threadIdx.x * per + k] = run;
    run += local[k];
  }
}

// perm[slot in sorted order] = queue index of the path.  The order inside a bin is whatever the atomics make it.
__global__ void __launch_bounds__(kSortThreads) k_sort_rank(const uint16_t* __restrict__ key, const int* __restrict__ n_ptr, uint32_t* __restrict__ cursor,
                                                           uint32_t* __restrict__ perm) {
  __shared__ uint32_t sh[kSortBins];  // per tile: count per bin, then the tile's first slot in the bin
  const int n = *n_ptr;
  const int n_tiles = (n + kSortTile - 1) / kSortTile;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) sh[b] = 0u;
    __syncthreads();
    const int first = tile * kSortTile + threadIdx.x;
    uint32_t kr[kSortPerThread];  // key | rank within the tile's share of the bin << 16
    // all of a thread's key loads first (unconditional, index clamped: "load, rank, load, rank" would pay sixteen memory
    // round trips per tile one after the other), then the ranks
#pragma unroll
    for (int j = 0; j < kSortPerThread; ++j) kr[j] = key[min(first + j * kSortThreads, n - 1)] & (uint32_t)(kSortBins - 1);
#pragma unroll
    for (int j = 0; j < kSortPerThread; ++j)
      if (first + j * kSortThreads < n) kr[j] |= atomicAdd(&sh[kr[j]], 1u) << 16;
    __syncthreads();
    {
      // all of a thread's claims in flight together: a loop of "atomic, then store its result" issues them one round trip
      // after the other (ncu: the kernel sat at 14 % issue, 33 long-scoreboard stalls per instruction)
      constexpr int kPer = (kSortBins + kSortThreads - 1) / kSortThreads;
      uint32_t cnt[kPer], got[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int b = threadIdx.x + k * kSortThreads;
        cnt[k] = b < kSortBins ? sh[b] : 0u;
      }
#pragma unroll
      for (int k = 0; k < kPer; ++k) got[k] = cnt[k] ? atomicAdd(&cursor[threadIdx.x + k * kSortThreads], cnt[k]) : 0u;
#pragma unroll
      for (int k = 0; k < kPer; ++k)
        if (cnt[k]) sh[threadIdx.x + k * kSortThreads] = got[k];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortPerThread; ++j) {
      const int i = first + j * kSortThreads;
      if (i < n) perm[sh[kr[j] & 0xFFFFu] + (kr[j] >> 16)] = (uint32_t)i;
    }
    __syncthreads();
  }
}

}  // namespace bn
