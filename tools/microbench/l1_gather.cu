// L1 data-pipe microbenchmark for the traversal's node fetch: every lane gathers one 16-B-aligned record per iteration
// from a random position of a working set (L1-resident / L2-resident) with
//   mode 0: 8 x LDG.128 (128 B)   mode 1: 4 x LDG.256 (128 B)   mode 2: 4 x LDG.128 (64 B)   mode 3: 2 x LDG.256 (64 B)
//   mode 4: 3 x LDG.128 (48 B, triangle)   mode 5: LDG.256 + LDG.128 (48 B in a 64-B slot)
// and reports records per second.  `active` lanes of every warp take part (SIMT divergence as in the phase loop).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/l1_gather.cu -o tools/microbench/l1_gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldg256(const void* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ float sum4(float4 a) { return a.x + a.y + a.z + a.w; }
__device__ __forceinline__ float sum8(F8 a) { float s = 0; for (int k = 0; k < 8; ++k) s += a.v[k]; return s; }
template <int MODE>
__global__ void __launch_bounds__(128, 9) k(const char* base, unsigned n_rec, unsigned stride, int iters, int active, float* out) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  if ((threadIdx.x & 31) < active) {
    for (int it = 0; it < iters; ++it) {
      s = s * 1664525u + 1013904223u;
      const char* p = base + (size_t)((s >> 8) % n_rec) * stride;
      const float4* q = reinterpret_cast<const float4*>(p);
      if (MODE == 0) { for (int j = 0; j < 8; ++j) acc += sum4(__ldg(q + j)); }
      if (MODE == 1) { for (int j = 0; j < 4; ++j) acc += sum8(ldg256(p + 32 * j)); }
      if (MODE == 2) { for (int j = 0; j < 4; ++j) acc += sum4(__ldg(q + j)); }
      if (MODE == 3) { for (int j = 0; j < 2; ++j) acc += sum8(ldg256(p + 32 * j)); }
      if (MODE == 4) { for (int j = 0; j < 3; ++j) acc += sum4(__ldg(q + j)); }
      if (MODE == 5) { acc += sum8(ldg256(p)) + sum4(__ldg(q + 2)); }
      s += __float_as_uint(acc) & 1u;   // dependent chain as in a tree walk
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  const size_t bytes = 64u << 20;
  char* buf; float* out;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 4); cudaMemset(buf, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 9, iters = 2000;
  const char* names[6] = {"8xLDG.128 (128B)", "4xLDG.256 (128B)", "4xLDG.128 (64B)", "2xLDG.256 (64B)", "3xLDG.128 (48B)", "LDG.256+LDG.128 (48B/64B slot)"};
  for (size_t ws : {(size_t)48 << 10, (size_t)1 << 20, (size_t)6 << 20})
    for (int active : {32, 16})
      for (int mode = 0; mode < 6; ++mode) {
        const unsigned stride = (mode == 0 || mode == 1) ? 128 : ((mode == 4) ? 48 : 64);
        const unsigned n_rec = (unsigned)(ws / stride);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          switch (mode) {
            case 0: k<0><<<grid, 128>>>(buf, n_rec, stride, iters, active, out); break;
            case 1: k<1><<<grid, 128>>>(buf, n_rec, stride, iters, active, out); break;
            case 2: k<2><<<grid, 128>>>(buf, n_rec, stride, iters, active, out); break;
            case 3: k<3><<<grid, 128>>>(buf, n_rec, stride, iters, active, out); break;
            case 4: k<4><<<grid, 128>>>(buf, n_rec, stride, iters, active, out); break;
            case 5: k<5><<<grid, 128>>>(buf, n_rec, stride, iters, active, out); break;
          }
          cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double recs = (double)grid * 4 * active * iters;
        printf("ws=%6zu KB active=%2d %-32s %8.2f G records/s  (%.3f ms)\n", ws >> 10, active, names[mode], recs / ms / 1e6, ms);
      }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
