// BVHNode.Build on the GPU ("next" row N2): the binned-SAH builder of
// Util/BVH.fs:109-247, restated level by level so that the node array AND the
// item permutation are the reference's, byte for byte.
//
// The reference recurses; every step of one recursion level is independent of
// the others, so all subtrees ("segments" of the item array) of one depth are
// processed together, item-parallel:
//   A  bounds + centroid bounds per segment           fold of MinNative/MaxNative  (:133-136, :141-145)
//   B  leaf test (:138), split axis / extent (:147-148), median split when extent = 0 (:150-156)
//   C  bin index per item (:163-169), bin bounds + counts (:171-177)
//   D  SAH sweep over the 11 split planes (:179-203), one thread per segment, the reference's
//      float operations in the reference's order; child segments
//   E  partition (:205-217): the left part keeps the item order, the right part is filled from
//      the end, i.e. reversed — positions follow from one prefix sum of the "goes left" flags
// and, when no segment is left, the nodes are numbered in preorder (Flatten, :224-237: left
// child = i+1, RightChild = i+1+|left subtree|) from subtree sizes computed bottom-up.
//
// Exactness: MinNative/MaxNative folds return, among equal extremes (only +0 / -0 can differ),
// the LAST operand.  The accumulators therefore hold (order-preserving key of the value with
// -0 folded onto +0, position in the segment) packed in 64 bits and are combined with 64-bit
// atomicMin/atomicMax, and the final bits are read back from the item that won.  Everything
// else is scalar fp32 evaluated once per segment exactly as the host builder does
// (scene_host.cpp, build_rec).  This file is compiled with -fmad=false.  Inputs must be finite
// (the reference's NaN propagation through minps/maxps is order-dependent); non-finite boxes
// are rejected with BN_ERR_INVALID.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/barnacle_b200.h"

namespace bnhost {
void set_error(const std::string& msg);
}

namespace bnbuild {

constexpr int kBins = 12;      // BVHBuildConfig.SAHBinCount (Util/BVH.fs:101)
constexpr int kMaxLeaf = 4;    // MaxLeafSize (:104)
constexpr int kMaxDepth = 64;  // MaxDepth (:107)
constexpr int kThreads = 256;
typedef unsigned long long u64;

struct Seg {
  int first, last;  // item range
  int depth;
  int node;   // BFS id of this segment's node
  int big;    // accumulator slot when last - first > kMaxLeaf and depth < kMaxDepth, else -1
  int mode;   // 0 leaf | 1 median split | 2 SAH split
  int axis;
  float cmin, extent;
  int best;   // best split plane
  int nl;     // items that go left
  int child;  // id of the left child segment in the next level's array (right = child + 1)
};

struct BNode {  // node in creation (breadth-first) order
  float lo[3], hi[3];
  int left, right;  // BFS ids, -1 for leaves
  int first, count;
  int axis;
  int size;  // nodes in the subtree
  int pre;   // preorder index
};

struct Counters { int n_next; int n_big_next; int n_nodes; int bad; };

__device__ __forceinline__ unsigned okey(float v) {  // order-preserving key, -0 folded onto +0
  const unsigned u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ u64 key_min(float v, int pos) { return ((u64)okey(v) << 32) | (u64)(0xFFFFFFFFu - (unsigned)pos); }
__device__ __forceinline__ u64 key_max(float v, int pos) { return ((u64)okey(v) << 32) | (u64)(unsigned)pos; }
__device__ __forceinline__ int pos_of_min(u64 k) { return (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)); }
__device__ __forceinline__ int pos_of_max(u64 k) { return (int)(unsigned)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ float min_native(float a, float b) { return a < b ? a : b; }  // minps
__device__ __forceinline__ float max_native(float a, float b) { return a > b ? a : b; }  // maxps

struct Box { float lo[3], hi[3]; };
__device__ __forceinline__ Box load_box(const float* __restrict__ boxes, int item) {
  Box b;
  const float* p = boxes + (size_t)item * 6;
  b.lo[0] = p[0]; b.lo[1] = p[1]; b.lo[2] = p[2]; b.hi[0] = p[3]; b.hi[1] = p[4]; b.hi[2] = p[5];
  return b;
}
__device__ __forceinline__ Box empty_box() {
  Box b;
  for (int a = 0; a < 3; ++a) { b.lo[a] = CUDART_INF_F; b.hi[a] = -CUDART_INF_F; }
  return b;
}
__device__ __forceinline__ void unite(Box& acc, const Box& x) {
  for (int a = 0; a < 3; ++a) { acc.lo[a] = min_native(acc.lo[a], x.lo[a]); acc.hi[a] = max_native(acc.hi[a], x.hi[a]); }
}
// AxisAlignedBoundingBox.Centroid: 0.5 * (pMin + pMax)
__device__ __forceinline__ float centroid(const Box& b, int a) { return 0.5f * (b.lo[a] + b.hi[a]); }
// SurfaceArea: 2 * (dx*dy + dy*dz + dz*dx), evaluated left to right
__device__ __forceinline__ float surface_area(const Box& b) {
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return 2.f * ((dx * dy + dy * dz) + dz * dx);
}

// accumulator layout per big segment: [0..5] bounds (min xyz, max xyz) | [6..11] centroid bounds
constexpr int kAccPerSeg = 12;
// bins per big segment: kBins x (6 keys) then kBins counts (as u64 for alignment simplicity)
constexpr int kBinKeys = kBins * 6;

__global__ void k_init_acc(u64* acc, u64* bins, int* bin_count, int n_big) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_big * kAccPerSeg) acc[i] = ((i % 6) < 3) ? ~0ull : 0ull;
  if (i < n_big * kBinKeys) bins[i] = ((i % 6) < 3) ? ~0ull : 0ull;
  if (i < n_big * kBins) bin_count[i] = 0;
}

__global__ void k_check_finite(const float* __restrict__ boxes, int n, Counters* c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Box b = load_box(boxes, i);
  bool ok = true;
  for (int a = 0; a < 3; ++a) ok = ok && isfinite(b.lo[a]) && isfinite(b.hi[a]);
  if (!ok) c->bad = 1;
}

// A: per item of a big segment, fold its box and its centroid into the segment's accumulators
__global__ void k_accumulate(const float* __restrict__ boxes, const int* __restrict__ idx, const int* __restrict__ segof, const Seg* __restrict__ segs,
                             int n, u64* acc) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int s = segof[pos];
  if (s < 0) return;
  const int big = segs[s].big;
  if (big < 0) return;
  const Box b = load_box(boxes, idx[pos]);
  u64* a = acc + (size_t)big * kAccPerSeg;
  for (int k = 0; k < 3; ++k) {
    atomicMin(a + k, key_min(b.lo[k], pos));
    atomicMax(a + 3 + k, key_max(b.hi[k], pos));
    const float c = centroid(b, k);
    atomicMin(a + 6 + k, key_min(c, pos));
    atomicMax(a + 9 + k, key_max(c, pos));
  }
}

// B: bounds of every active segment, leaf test, split axis
__global__ void k_segments_bounds(const float* __restrict__ boxes, const int* __restrict__ idx, Seg* segs, int n_seg, const u64* __restrict__ acc, BNode* nodes) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  Seg sg = segs[s];
  const int count = sg.last - sg.first;
  Box bounds = empty_box();
  Box cb = empty_box();
  if (sg.big < 0) {
    for (int i = sg.first; i < sg.last; ++i) unite(bounds, load_box(boxes, idx[i]));
  } else {
    const u64* a = acc + (size_t)sg.big * kAccPerSeg;
    for (int k = 0; k < 3; ++k) {
      bounds.lo[k] = load_box(boxes, idx[pos_of_min(a[k])]).lo[k];
      bounds.hi[k] = load_box(boxes, idx[pos_of_max(a[3 + k])]).hi[k];
      cb.lo[k] = centroid(load_box(boxes, idx[pos_of_min(a[6 + k])]), k);
      cb.hi[k] = centroid(load_box(boxes, idx[pos_of_max(a[9 + k])]), k);
    }
  }
  BNode& nd = nodes[sg.node];
  for (int k = 0; k < 3; ++k) { nd.lo[k] = bounds.lo[k]; nd.hi[k] = bounds.hi[k]; }
  nd.first = sg.first; nd.count = count;
  nd.left = nd.right = -1; nd.axis = 0; nd.size = 1; nd.pre = 0;
  if (count <= kMaxLeaf || sg.depth >= kMaxDepth) {  // :138
    sg.mode = 0;
  } else {
    // SplitAxis (Util/BVH.fs:24-27) and Diagonal[axis]
    const float dx = cb.hi[0] - cb.lo[0], dy = cb.hi[1] - cb.lo[1], dz = cb.hi[2] - cb.lo[2];
    const int axis = (dx >= dy && dx >= dz) ? 0 : (dy >= dz ? 1 : 2);
    const float extent = axis == 0 ? dx : (axis == 1 ? dy : dz);
    sg.axis = axis;
    sg.cmin = cb.lo[axis];
    sg.extent = extent;
    sg.mode = extent == 0.f ? 1 : 2;
    nd.axis = axis;
    // base term of the SAH cost: count * centroidBounds.SurfaceArea (:199); kept in `best` bits until D
    sg.best = __float_as_int((float)count * surface_area(cb));
  }
  segs[s] = sg;
}

__device__ __forceinline__ int bin_index(const Box& b, const Seg& sg) {  // :163-169
  const float f = ((float)kBins * (centroid(b, sg.axis) - sg.cmin)) / sg.extent;
  const int bi = (int)f;  // truncation, like F#'s `int`
  return bi < kBins - 1 ? bi : kBins - 1;
}

// C: bin every item of a SAH-split segment
__global__ void k_bin(const float* __restrict__ boxes, const int* __restrict__ idx, const int* __restrict__ segof, const Seg* __restrict__ segs, int n,
                      unsigned char* __restrict__ bin_of, u64* bins, int* bin_count) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int s = segof[pos];
  if (s < 0) return;
  const Seg sg = segs[s];
  if (sg.mode != 2) return;
  const Box b = load_box(boxes, idx[pos]);
  const int bi = bin_index(b, sg);
  bin_of[pos] = (unsigned char)bi;
  u64* k = bins + ((size_t)sg.big * kBins + bi) * 6;
  for (int a = 0; a < 3; ++a) {
    atomicMin(k + a, key_min(b.lo[a], pos));
    atomicMax(k + 3 + a, key_max(b.hi[a], pos));
  }
  atomicAdd(bin_count + (size_t)sg.big * kBins + bi, 1);
}

// D: SAH sweep, children
__global__ void k_segments_split(const float* __restrict__ boxes, const int* __restrict__ idx, Seg* segs, int n_seg, const u64* __restrict__ bins,
                                 const int* __restrict__ bin_count, Seg* next, BNode* nodes, Counters* ctr) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  Seg sg = segs[s];
  if (sg.mode == 0) return;
  const int count = sg.last - sg.first;
  int nl;
  if (sg.mode == 1) {
    nl = count / 2;  // :151
  } else {
    Box bb[kBins];
    int bc[kBins];
    for (int i = 0; i < kBins; ++i) {
      bc[i] = bin_count[(size_t)sg.big * kBins + i];
      bb[i] = empty_box();
      if (bc[i] > 0) {
        const u64* k = bins + ((size_t)sg.big * kBins + i) * 6;
        for (int a = 0; a < 3; ++a) {
          bb[i].lo[a] = load_box(boxes, idx[pos_of_min(k[a])]).lo[a];
          bb[i].hi[a] = load_box(boxes, idx[pos_of_max(k[3 + a])]).hi[a];
        }
      }
    }
    constexpr int nc = kBins - 1;
    float costs[nc];
    for (int i = 0; i < nc; ++i) costs[i] = 0.f;
    Box lb = empty_box(), rb = empty_box();
    int lc = 0, rc = 0;
    for (int i = 0; i < nc; ++i) {  // :187-193
      unite(lb, bb[i]);
      lc += bc[i];
      costs[i] = costs[i] + (float)lc * surface_area(lb);
      unite(rb, bb[nc - i]);
      rc += bc[nc - i];
      costs[nc - 1 - i] = costs[nc - 1 - i] + (float)rc * surface_area(rb);
    }
    float min_cost = CUDART_INF_F;
    int best = 0;
    const float base = __int_as_float(sg.best);
    for (int i = 0; i < nc; ++i) {  // :198-203
      const float cost = costs[i] + base;
      if (cost < min_cost) { min_cost = cost; best = i; }
    }
    sg.best = best;
    nl = 0;
    for (int i = 0; i <= best; ++i) nl += bc[i];
  }
  sg.nl = nl;
  const int child = atomicAdd(&ctr->n_next, 2);
  const int cnode = atomicAdd(&ctr->n_nodes, 2);
  sg.child = child;
  Seg l{}, r{};
  l.first = sg.first; l.last = sg.first + nl; l.depth = sg.depth + 1; l.node = cnode;
  r.first = sg.first + nl; r.last = sg.last; r.depth = sg.depth + 1; r.node = cnode + 1;
  l.big = (l.last - l.first > kMaxLeaf && l.depth < kMaxDepth) ? atomicAdd(&ctr->n_big_next, 1) : -1;
  r.big = (r.last - r.first > kMaxLeaf && r.depth < kMaxDepth) ? atomicAdd(&ctr->n_big_next, 1) : -1;
  next[child] = l;
  next[child + 1] = r;
  nodes[sg.node].left = cnode;
  nodes[sg.node].right = cnode + 1;
  segs[s] = sg;
}

// E1: "goes left" flag per position (finished / leaf / median segments keep their place)
__global__ void k_flags(const int* __restrict__ segof, const Seg* __restrict__ segs, const unsigned char* __restrict__ bin_of, int n, int* __restrict__ flag) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int s = segof[pos];
  int f = 0;
  if (s >= 0 && segs[s].mode == 2) f = (int)bin_of[pos] <= segs[s].best ? 1 : 0;
  flag[pos] = f;
}

// exclusive prefix sum, three passes (block scan, scan of the block sums, add)
constexpr int kScanItems = 4;
__global__ void __launch_bounds__(kThreads) k_scan_block(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ block_sums, int n) {
  __shared__ int sh[kThreads];
  const int base = (blockIdx.x * kThreads + threadIdx.x) * kScanItems;
  int v[kScanItems], sum = 0;
  for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < n ? in[base + k] : 0; sum += v[k]; }
  sh[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < kThreads; off <<= 1) {
    const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  int run = sh[threadIdx.x] - sum;
  for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
  if (threadIdx.x == kThreads - 1) block_sums[blockIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(kThreads) k_scan_sums(int* __restrict__ block_sums, int n_blocks) {
  __shared__ int sh[kThreads];
  int carry = 0;
  for (int b0 = 0; b0 < n_blocks; b0 += kThreads) {
    const int i = b0 + threadIdx.x;
    const int v = i < n_blocks ? block_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kThreads; off <<= 1) {
      const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_blocks) block_sums[i] = carry + sh[threadIdx.x] - v;
    carry += sh[kThreads - 1];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kThreads) k_scan_add(int* __restrict__ out, const int* __restrict__ block_sums, int n) {
  const int base = (blockIdx.x * kThreads + threadIdx.x) * kScanItems;
  const int add = block_sums[blockIdx.x];
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) out[base + k] += add;
}

// E2: scatter (:205-217)
__global__ void k_scatter(const int* __restrict__ idx, const int* __restrict__ segof, const Seg* __restrict__ segs, const int* __restrict__ flag,
                          const int* __restrict__ scan, int n, int* __restrict__ idx_out, int* __restrict__ segof_out) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int s = segof[pos];
  int np = pos, ns = -1;
  if (s >= 0) {
    const Seg sg = segs[s];
    if (sg.mode == 2) {
      const int lefts_before = scan[pos] - scan[sg.first];
      if (flag[pos]) { np = sg.first + lefts_before; ns = sg.child; }
      else { np = sg.last - 1 - ((pos - sg.first) - lefts_before); ns = sg.child + 1; }
    } else if (sg.mode == 1) {
      ns = (pos - sg.first) < sg.nl ? sg.child : sg.child + 1;
    }
  }
  idx_out[np] = idx[pos];
  segof_out[np] = ns;
}

// Flatten (:224-237)
__global__ void k_sizes(BNode* nodes, int begin, int end) {
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  BNode& nd = nodes[i];
  nd.size = nd.left < 0 ? 1 : 1 + nodes[nd.left].size + nodes[nd.right].size;
}
__global__ void k_preorder(BNode* nodes, int begin, int end) {
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  const BNode nd = nodes[i];
  if (nd.left < 0) return;
  nodes[nd.left].pre = nd.pre + 1;
  nodes[nd.right].pre = nd.pre + 1 + nodes[nd.left].size;
}
__global__ void k_emit(const BNode* __restrict__ nodes, int n_nodes, BnBVHNode* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const BNode nd = nodes[i];
  BnBVHNode o;
  for (int k = 0; k < 3; ++k) { o.bounds_min[k] = nd.lo[k]; o.bounds_max[k] = nd.hi[k]; }
  o.visibility_mask = 0;
  if (nd.left < 0) {  // BVHNode.CreateLeaf
    o.right_or_offset = nd.first; o.is_leaf = 1; o.split_axis = 0; o.count = (int8_t)nd.count;
  } else {            // BVHNode.CreateInterior
    o.right_or_offset = nodes[nd.right].pre; o.is_leaf = 0; o.split_axis = (int8_t)nd.axis; o.count = 0;
  }
  out[nd.pre] = o;
}
__global__ void k_iota(int* idx, int* segof, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { idx[i] = i; segof[i] = 0; }
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

inline int blocks(int n) { return (n + kThreads - 1) / kThreads; }

// d_boxes: n x 6 floats; d_nodes: >= max_nodes BnBVHNode; d_perm: n uint32 (may be null).
int build_device(const float* d_boxes, int n, BnBVHNode* d_nodes, uint32_t max_nodes, uint32_t* d_perm, cudaStream_t st, float* ms) {
  const int max_total_nodes = 2 * n - 1;
  const int max_big = n / (kMaxLeaf + 1) + 2;
  DevBuf b_idx[2], b_seg[2], b_segs[2], b_nodes, b_acc, b_bins, b_bc, b_binof, b_flag, b_scan, b_sums, b_ctr;
  const int scan_blocks = (n + kThreads * kScanItems - 1) / (kThreads * kScanItems);
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  for (int k = 0; k < 2; ++k) { ok(b_idx[k].alloc(sizeof(int) * (size_t)n)); ok(b_seg[k].alloc(sizeof(int) * (size_t)n)); ok(b_segs[k].alloc(sizeof(Seg) * ((size_t)n + 2))); }
  ok(b_nodes.alloc(sizeof(BNode) * (size_t)(max_total_nodes + 2)));
  ok(b_acc.alloc(sizeof(u64) * kAccPerSeg * (size_t)max_big));
  ok(b_bins.alloc(sizeof(u64) * kBinKeys * (size_t)max_big));
  ok(b_bc.alloc(sizeof(int) * kBins * (size_t)max_big));
  ok(b_binof.alloc((size_t)n));
  ok(b_flag.alloc(sizeof(int) * (size_t)n));
  ok(b_scan.alloc(sizeof(int) * (size_t)n));
  ok(b_sums.alloc(sizeof(int) * (size_t)(scan_blocks + 1)));
  ok(b_ctr.alloc(sizeof(Counters)));
  if (e != cudaSuccess) { cudaGetLastError(); bnhost::set_error(std::string("bn_bvh_build: cudaMalloc failed: ") + cudaGetErrorString(e)); return BN_ERR_CUDA; }

  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0); cudaEventCreate(&ev1);
  cudaEventRecord(ev0, st);
  Counters h{};
  cudaMemsetAsync(b_ctr.p, 0, sizeof(Counters), st);
  k_check_finite<<<blocks(n), kThreads, 0, st>>>(d_boxes, n, b_ctr.as<Counters>());
  k_iota<<<blocks(n), kThreads, 0, st>>>(b_idx[0].as<int>(), b_seg[0].as<int>(), n);
  Seg root{};
  root.first = 0; root.last = n; root.depth = 0; root.node = 0; root.big = n > kMaxLeaf ? 0 : -1;
  cudaMemcpyAsync(b_segs[0].p, &root, sizeof root, cudaMemcpyHostToDevice, st);
  int n_seg = 1, n_big = n > kMaxLeaf ? 1 : 0, n_nodes = 1, cur = 0;
  std::vector<int> level_start{0};
  int rc = BN_OK;
  while (n_seg > 0) {
    level_start.push_back(n_nodes);
    const int c0[3] = {0, 0, n_nodes};  // n_next, n_big_next, n_nodes (the `bad` flag is left alone)
    cudaMemcpyAsync(b_ctr.p, c0, sizeof c0, cudaMemcpyHostToDevice, st);
    if (n_big > 0) {
      k_init_acc<<<blocks(n_big * kBinKeys), kThreads, 0, st>>>(b_acc.as<u64>(), b_bins.as<u64>(), b_bc.as<int>(), n_big);
      k_accumulate<<<blocks(n), kThreads, 0, st>>>(d_boxes, b_idx[cur].as<int>(), b_seg[cur].as<int>(), b_segs[cur].as<Seg>(), n, b_acc.as<u64>());
    }
    k_segments_bounds<<<blocks(n_seg), kThreads, 0, st>>>(d_boxes, b_idx[cur].as<int>(), b_segs[cur].as<Seg>(), n_seg, b_acc.as<u64>(), b_nodes.as<BNode>());
    if (n_big > 0) {
      k_bin<<<blocks(n), kThreads, 0, st>>>(d_boxes, b_idx[cur].as<int>(), b_seg[cur].as<int>(), b_segs[cur].as<Seg>(), n, b_binof.as<unsigned char>(),
                                            b_bins.as<u64>(), b_bc.as<int>());
      k_segments_split<<<blocks(n_seg), kThreads, 0, st>>>(d_boxes, b_idx[cur].as<int>(), b_segs[cur].as<Seg>(), n_seg, b_bins.as<u64>(), b_bc.as<int>(),
                                                           b_segs[cur ^ 1].as<Seg>(), b_nodes.as<BNode>(), b_ctr.as<Counters>());
      k_flags<<<blocks(n), kThreads, 0, st>>>(b_seg[cur].as<int>(), b_segs[cur].as<Seg>(), b_binof.as<unsigned char>(), n, b_flag.as<int>());
      k_scan_block<<<scan_blocks, kThreads, 0, st>>>(b_flag.as<int>(), b_scan.as<int>(), b_sums.as<int>(), n);
      k_scan_sums<<<1, kThreads, 0, st>>>(b_sums.as<int>(), scan_blocks);
      k_scan_add<<<scan_blocks, kThreads, 0, st>>>(b_scan.as<int>(), b_sums.as<int>(), n);
      k_scatter<<<blocks(n), kThreads, 0, st>>>(b_idx[cur].as<int>(), b_seg[cur].as<int>(), b_segs[cur].as<Seg>(), b_flag.as<int>(), b_scan.as<int>(), n,
                                                b_idx[cur ^ 1].as<int>(), b_seg[cur ^ 1].as<int>());
      cur ^= 1;
    }
    cudaMemcpyAsync(&h, b_ctr.p, sizeof h, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    if (h.bad) { bnhost::set_error("bn_bvh_build: boxes must be finite"); rc = BN_ERR_INVALID; break; }
    n_seg = n_big > 0 ? h.n_next : 0;
    n_big = h.n_big_next;
    n_nodes = h.n_nodes;
  }
  if (e == cudaSuccess && rc == BN_OK) {
    if ((uint32_t)n_nodes > max_nodes) { bnhost::set_error("bn_bvh_build: node buffer too small"); rc = BN_ERR_INVALID; }
    else {
      level_start.push_back(n_nodes);
      const int L = (int)level_start.size() - 1;  // level l holds BFS ids [level_start[l], level_start[l+1])
      for (int l = L - 1; l >= 0; --l)
        if (level_start[l + 1] > level_start[l])
          k_sizes<<<blocks(level_start[l + 1] - level_start[l]), kThreads, 0, st>>>(b_nodes.as<BNode>(), level_start[l], level_start[l + 1]);
      for (int l = 0; l < L; ++l)
        if (level_start[l + 1] > level_start[l])
          k_preorder<<<blocks(level_start[l + 1] - level_start[l]), kThreads, 0, st>>>(b_nodes.as<BNode>(), level_start[l], level_start[l + 1]);
      k_emit<<<blocks(n_nodes), kThreads, 0, st>>>(b_nodes.as<BNode>(), n_nodes, d_nodes);
      if (d_perm) cudaMemcpyAsync(d_perm, b_idx[cur].p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st);
      cudaEventRecord(ev1, st);
      e = cudaStreamSynchronize(st);
      if (e == cudaSuccess && ms) cudaEventElapsedTime(ms, ev0, ev1);
    }
  }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  if (e != cudaSuccess) { cudaGetLastError(); bnhost::set_error(std::string("bn_bvh_build: ") + cudaGetErrorString(e)); return BN_ERR_CUDA; }
  return rc == BN_OK ? n_nodes : rc;
}

}  // namespace bnbuild

extern "C" {

int bn_bvh_build_device(int device, const void* d_boxes, uint32_t n, void* d_nodes, uint32_t max_nodes, void* d_perm, void* cuda_stream, float* ms) {
  if (!d_boxes || !d_nodes || n == 0 || n > (1u << 30)) { bnhost::set_error("bn_bvh_build_device: bad arguments"); return BN_ERR_INVALID; }
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || device < 0 || device >= nd) { cudaGetLastError(); bnhost::set_error("no CUDA device"); return BN_ERR_NO_DEVICE; }
  cudaSetDevice(device);
  return bnbuild::build_device(static_cast<const float*>(d_boxes), (int)n, static_cast<BnBVHNode*>(d_nodes), max_nodes, static_cast<uint32_t*>(d_perm),
                               static_cast<cudaStream_t>(cuda_stream), ms);
}

int bn_bvh_build(int device, const float* boxes, uint32_t n, BnBVHNode* nodes, uint32_t max_nodes, uint32_t* perm, float* ms) {
  if (!boxes || !nodes || n == 0 || n > (1u << 30)) { bnhost::set_error("bn_bvh_build: bad arguments"); return BN_ERR_INVALID; }
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || device < 0 || device >= nd) { cudaGetLastError(); bnhost::set_error("no CUDA device"); return BN_ERR_NO_DEVICE; }
  cudaSetDevice(device);
  bnbuild::DevBuf db, dn, dp;
  if (db.alloc(sizeof(float) * 6 * (size_t)n) != cudaSuccess || dn.alloc(sizeof(BnBVHNode) * (size_t)(2 * n - 1)) != cudaSuccess || dp.alloc(sizeof(uint32_t) * (size_t)n) != cudaSuccess) {
    cudaGetLastError(); bnhost::set_error("bn_bvh_build: cudaMalloc failed"); return BN_ERR_CUDA;
  }
  cudaMemcpy(db.p, boxes, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice);
  const int rc = bnbuild::build_device(db.as<float>(), (int)n, dn.as<BnBVHNode>(), 2 * n - 1, dp.as<uint32_t>(), nullptr, ms);
  if (rc < 0) return rc;
  if ((uint32_t)rc > max_nodes) { bnhost::set_error("bn_bvh_build: node buffer too small"); return BN_ERR_INVALID; }
  cudaMemcpy(nodes, dn.p, sizeof(BnBVHNode) * (size_t)rc, cudaMemcpyDeviceToHost);
  if (perm) cudaMemcpy(perm, dp.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost);
  return rc;
}

}  // extern "C"
