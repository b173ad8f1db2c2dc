// Internal (not part of the C ABI): the BnScene object shared by kernels.cu and mlt.cu.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <vector>

#include "device_scene.h"

using bn::DScene;

struct BnScene {
  int device = 0;
  int num_sms = 0;
  DScene d{};
  void* arena = nullptr;     // every scene array, one allocation (parked with the wave buffers when the scene dies)
  size_t arena_bytes = 0;
  // wave buffers
  size_t cap = 0;
  float4* state[3] = {nullptr, nullptr, nullptr};  // [0], [1]: 3 planes each; [2]: planes 0, 1 of a bounce's paths in its ordering (ray_sort.cuh)
  uint32_t* sort_perm = nullptr;          // ray_sort.cuh: sorted slot -> queue index (cap entries)
  uint16_t* sort_key = nullptr;           // per queued path (cap entries)
  uint32_t* sort_bins = nullptr;          // histogram + cursors (2 x kSortBins)
  float4* hits = nullptr;
  float4* shq = nullptr;                  // 4 planes
  float4* rad = nullptr;
  int* defer_list = nullptr;              // rays deferred to the exact fix-up kernel (cap entries)
  uint32_t* cand = nullptr;               // small-TLAS candidate word per queued ray (k_candidates; cap entries)
  int* counters = nullptr;
  size_t counters_len = 0;
  unsigned long long* shadow_ref = nullptr;
  float* film = nullptr;
  size_t film_len = 0;
  // PSSMLT scratch (mlt.cu), grown on demand and kept for the scene's lifetime: primary-sample
  // arrays, bootstrap weights (device + pinned host), counters
  float* mlt_f = nullptr;
  int* mlt_i = nullptr;
  size_t mlt_len = 0;            // elements of each of the 2+2 primary-sample arrays
  float* mlt_w = nullptr;
  float* mlt_w_host = nullptr;   // pinned
  size_t mlt_w_len = 0;
  unsigned long long* mlt_cnt = nullptr;  // 4 counters
  unsigned int* mlt_acc = nullptr;
  size_t mlt_acc_len = 0;
  int* mlt_wave_i = nullptr;     // wavefront PSSMLT: per-lane sampler + chain state (12 ints / 8 floats per lane)
  float* mlt_wave_f = nullptr;
  size_t mlt_wave_len = 0;
  int* mlt_counters = nullptr;   // ... and the round's queue counters
  size_t mlt_counters_len = 0;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;  // frame / batch timing (created once, destroyed with the scene)
  int* trace_ctr = nullptr;      // bn_trace_device scratch: per-chunk cursor + deferred count
  size_t trace_ctr_len = 0;
  int* trace_dlist = nullptr;    // ... and its deferred list
  size_t trace_dlist_len = 0;
  bool poisoned = false;
  std::vector<cudaEvent_t> events;  // BN_RENDER_PROFILE: start/stop pairs, one per kernel launch
};


// Internal entry points shared by kernels.cu and multi.cu (hidden visibility: not part of the C ABI).
#include "../../../include/barnacle_b200.h"
#include "scene_convert.h"
namespace bnint {
struct Staged;                                                                                   // the flattened scene as one host image (kernels.cu)
int stage_scene(const BnSceneDesc* desc, std::shared_ptr<const Staged>& out);                   // host: flatten + validate (or the cached image)
int scene_from_staged(const Staged& st, const BnCamera& camera, int device, BnScene** out);     // device: one allocation, one copy
int scene_film(BnScene* s, size_t len, float** out);                                             // the scene's device film, grown on demand
// the wavefront's traversal stages for other integrators (mlt.cu): closest hit for the rays (s0, s1) -> hits; any hit for the
// shadow queue (q0..q3) + connect into rad; both with their exact fix-up launch.  Counters: *n_ptr rays, a work cursor and a
// deferred count (zeroed by the caller).  ensure_wave grows the scene's queues to `cap` paths.
int ensure_wave(BnScene* s, size_t cap);
void launch_extend(BnScene* s, cudaStream_t stream, const float4* s0, const float4* s1, float4* hits, const int* n_ptr, int* cursor, int* n_defer);
void launch_shadow(BnScene* s, cudaStream_t stream, const float4* q0, const float4* q1, const float4* q2, const float4* q3, float4* rad, const int* n_ptr,
                   int* cursor, int* n_defer);
int render_on_stream(BnScene* s, const BnRenderParams* p, float* d_film, cudaStream_t stream, BnStats* stats);  // film stays on the device
}  // namespace bnint
