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
// ---- warp-level primitives -----------------------------------------------------------------------------------
// Plain builds: trivial one-lane definitions (traverse_persistent is never instantiated, they only have to parse).
// With BN_HOSTSIM_WARP (warp_emulator.h) they are rendezvous points of 32 lanes that run as fibers in lock step, and
// threadIdx.x reads the running lane's id: the warp-synchronous traversal loop then runs on the host as it is written.
#ifdef BN_HOSTSIM_WARP
unsigned hostsim_warp_sync(int kind, unsigned value, int src_lane);  // 0 REDUX.SUM, 1 ballot, 2 shuffle
unsigned hostsim_lane_id(void);
void hostsim_work(unsigned units);  // BN_WORK (traverse.cuh, BN_TRAV_STATS builds): per-lane work for the SIMT cost model
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return hostsim_warp_sync(0, v, 0); }
static inline unsigned __ballot_sync(unsigned, int p) { return hostsim_warp_sync(1, p ? 1u : 0u, 0); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)hostsim_warp_sync(2, (unsigned)v, src); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return hostsim_warp_sync(2, v, src); }
static inline float __shfl_sync(unsigned, float v, int src) { return __uint_as_float(hostsim_warp_sync(2, __float_as_uint(v), src)); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)hostsim_warp_sync(1, 0u, 0); }  // a rendezvous like any other (a ballot nobody reads)
struct HostsimLaneX { operator unsigned() const { return hostsim_lane_id(); } };
struct HostsimThreadIdx { HostsimLaneX x; unsigned y = 0, z = 0; };
static HostsimThreadIdx threadIdx;
#else
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
struct HostsimThreadIdx { unsigned x = 0, y = 0, z = 0; };
static HostsimThreadIdx threadIdx;
#endif
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }  // lanes are fibers of one thread: no race
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
struct HostsimDim { unsigned x = 1, y = 1, z = 1; };
static HostsimDim blockDim, gridDim;
struct HostsimBlockIdx { unsigned x = 0, y = 0, z = 0; };
static HostsimBlockIdx blockIdx;
