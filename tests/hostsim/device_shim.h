// TEST INFRASTRUCTURE — lets g++ compile the kernels' __device__ functions (vecmath.cuh, traverse.cuh, shade.cuh) for
// the HOST, so that tests/test_hostsim.py can run the GPU path's own arithmetic on the CPU and compare it bit for bit
// with the oracle.  Nothing in the product includes this file; the product path stays CUDA-only.
// Every intrinsic below is the IEEE operation the CUDA one is defined as (round-to-nearest, no contraction: the
// translation unit is compiled with -ffp-contract=off, as the device code is with -fmad=false).
#pragma once
#include <cuda_runtime.h>  // vector types (float3, float4, uint2 ...) and make_*; no device code is generated

#include <cmath>
#include <cstdint>
#include <cstring>

#define BN_HOSTSIM 1
#undef __forceinline__
#define __forceinline__ inline

static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fsqrt_rn(float x) { return std::sqrt(x); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
// warp-level primitives: only named inside traverse_persistent(), which the host never instantiates; they have to parse
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
struct HostsimDim { unsigned x = 0, y = 0, z = 0; };
static HostsimDim threadIdx, blockIdx, blockDim, gridDim;
