// Multi-device render behind ONE C-ABI call: the reference keeps the whole machine's parallelism inside
// Integrator.Render (Base/Integrator.fs:46-55 — 16x16 tiles over the TPL pool, one writer per pixel), so a drop-in for it
// fans out over the GPUs of the box from inside the call as well: bn_multi_scene_create flattens the scene ONCE and
// uploads it to every device (one host thread each), bn_render_multi gives every device its share of the (pixel,
// sampleId) space — every path is independent and fully determined by its seed, Integrator.fs:35-36 — and combines the
// per-device films on the first device with one kernel that reads the peers' films over NVLink (peer access; staged
// peer copies where the topology offers none), in a fixed device order, so the result is deterministic.
// No NCCL, no torch: a managed host P/Invokes this like bn_render.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/barnacle_b200.h"
#include "scene_internal.h"

namespace bnhost {
void set_error(const std::string& msg);
}

struct BnMultiScene {
  std::vector<int> devices;
  std::vector<BnScene*> scenes;
  std::vector<cudaStream_t> streams;
  std::vector<float*> films;     // every scene's own W*H*3 device film (parked with the scene's buffers between scenes)
  std::vector<char> direct;      // device k's film can be read from devices[0] (same device or peer access enabled)
  float* staging = nullptr;      // on devices[0]: landing buffer for peers without direct access
  size_t staging_len = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

constexpr int kMaxPeers = 15;
struct PeerFilms {
  const float* p[kMaxPeers];
  int n;
};

// dst[i] += sum over the peers, in device order (a fixed order of fp32 additions: the same film every run).  Sample split:
// the partial sums of a pixel's fma chain (Integrator.fs:41-44) are added here; tile split: every pixel is non-zero in
// exactly one film, so the sum is a gather and changes no bit.
__global__ void __launch_bounds__(256) k_film_reduce(float* __restrict__ dst, PeerFilms peers, size_t n) {
  const size_t n4 = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    for (int k = 0; k < peers.n; ++k) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(peers.p[k]) + i);  // streamed once over NVLink
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float a = dst[i];
    for (int k = 0; k < peers.n; ++k) a += peers.p[k][i];
    dst[i] = a;
  }
}

bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  bnhost::set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}

// The share of device `rank` of `n` (same rule as barnacle_b200/multi_gpu.py: sample split when there are at least as
// many samples as devices, else round-robin 16-pixel tile rows — the reference's tile size, Integrator.fs:16).
// Returns false for an empty share.
bool shard_params(const BnRenderParams& base, int mode, int n, int rank, BnRenderParams& out) {
  out = base;
  const int ns = base.sample_end - base.sample_begin;
  if (mode == BN_PARTITION_AUTO) mode = ns >= n ? BN_PARTITION_SAMPLE : BN_PARTITION_TILE;
  if (mode == BN_PARTITION_SAMPLE) {
    out.sample_begin = base.sample_begin + (int)(((long long)rank * ns) / n);
    out.sample_end = base.sample_begin + (int)(((long long)(rank + 1) * ns) / n);
    return out.sample_begin < out.sample_end;
  }
  const int rows = (base.y1 - base.y0 + 15) / 16;
  out.interleave_count = n;
  out.interleave_index = rank;
  return rank < rows && ns > 0;
}

}  // namespace

extern "C" {

int bn_multi_partition(const BnRenderParams* params, int32_t partition, int32_t n_devices, int32_t rank, BnRenderParams* out, int32_t* empty) {
  if (!params || !out || n_devices < 1 || rank < 0 || rank >= n_devices || partition < BN_PARTITION_AUTO || partition > BN_PARTITION_TILE ||
      (partition != BN_PARTITION_SAMPLE && params->interleave_count > 1 && n_devices > 1 && params->sample_end - params->sample_begin < n_devices) ||
      (partition == BN_PARTITION_TILE && params->interleave_count > 1 && n_devices > 1)) {
    bnhost::set_error("bn_multi_partition: bad argument (a window that is already tile-interleaved cannot be tile-split again)");
    return BN_ERR_INVALID;
  }
  const bool any = shard_params(*params, partition, n_devices, rank, *out);
  if (empty) *empty = any ? 0 : 1;
  return BN_OK;
}

int bn_multi_scene_create(const BnSceneDesc* desc, const int32_t* devices, int32_t n_devices, BnMultiScene** out) {
  if (!desc || !devices || !out || n_devices < 1 || n_devices > kMaxPeers + 1) {
    bnhost::set_error("bn_multi_scene_create: bad argument (1..16 devices)");
    return BN_ERR_INVALID;
  }
  *out = nullptr;
  const int ndev = bn_device_count();
  if (ndev <= 0) { bnhost::set_error("no CUDA device available (the hot path has no CPU fallback)"); return BN_ERR_NO_DEVICE; }
  for (int k = 0; k < n_devices; ++k)
    if (devices[k] < 0 || devices[k] >= ndev) { bnhost::set_error("bn_multi_scene_create: device ordinal out of range"); return BN_ERR_INVALID; }
  std::shared_ptr<const bnint::Staged> staged;  // flattened ONCE (or taken from the cache), shared by every upload
  int rc = bnint::stage_scene(desc, staged);
  if (rc != BN_OK) return rc;
  const BnCamera camera = desc->camera;
  auto* m = new BnMultiScene();
  m->devices.assign(devices, devices + n_devices);
  m->scenes.assign(n_devices, nullptr);
  m->streams.assign(n_devices, nullptr);
  m->films.assign(n_devices, nullptr);
  m->direct.assign(n_devices, 0);
  std::vector<int> rcs(n_devices, BN_OK);
  std::vector<std::string> errs(n_devices);
  auto upload = [&](int k) {
    rcs[k] = bnint::scene_from_staged(*staged, camera, m->devices[k], &m->scenes[k]);
    if (rcs[k] == BN_OK && cudaStreamCreateWithFlags(&m->streams[k], cudaStreamNonBlocking) != cudaSuccess) rcs[k] = BN_ERR_CUDA;
    if (rcs[k] != BN_OK) errs[k] = bn_last_error();
  };
  {
    std::vector<std::thread> th;
    for (int k = 1; k < n_devices; ++k) th.emplace_back(upload, k);
    upload(0);
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < n_devices; ++k)
    if (rcs[k] != BN_OK) {
      const std::string e = errs[k];
      const int r = rcs[k];
      bn_multi_scene_destroy(m);
      bnhost::set_error("device " + std::to_string(devices[k]) + ": " + e);
      return r;
    }
  // peer access from the first device to the others (NVLink / NVSwitch on the 8 x B200 box)
  cudaSetDevice(m->devices[0]);
  for (int k = 0; k < n_devices; ++k) {
    if (m->devices[k] == m->devices[0]) { m->direct[k] = 1; continue; }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, m->devices[0], m->devices[k]) == cudaSuccess && can) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[k], 0);
      if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->direct[k] = 1;
    }
    cudaGetLastError();
  }
  if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) {
    bn_multi_scene_destroy(m);
    bnhost::set_error("cudaEventCreate failed");
    return BN_ERR_CUDA;
  }
  *out = m;
  return BN_OK;
}

void bn_multi_scene_destroy(BnMultiScene* m) {
  if (!m) return;
  for (size_t k = 0; k < m->scenes.size(); ++k) {
    cudaSetDevice(m->devices[k]);
    if (m->streams[k]) cudaStreamDestroy(m->streams[k]);
    if (m->scenes[k]) bn_scene_destroy(m->scenes[k]);
  }
  if (!m->devices.empty()) {
    cudaSetDevice(m->devices[0]);
    if (m->staging) cudaFree(m->staging);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
  }
  delete m;
}

int bn_multi_scene_device_count(const BnMultiScene* m) { return m ? (int)m->devices.size() : 0; }

int bn_render_multi(BnMultiScene* m, const BnRenderParams* p, int32_t partition, float* film_rgb, BnStats* stats) {
  if (!m || !p || !film_rgb) { bnhost::set_error("bn_render_multi: NULL argument"); return BN_ERR_INVALID; }
  const int n = (int)m->devices.size();
  if (partition < BN_PARTITION_AUTO || partition > BN_PARTITION_TILE || p->width <= 0 || p->height <= 0 ||
      (p->interleave_count > 1 && n > 1 && (partition == BN_PARTITION_TILE || (partition == BN_PARTITION_AUTO && p->sample_end - p->sample_begin < n)))) {
    bnhost::set_error("bn_render_multi: bad partition / a window that is already tile-interleaved cannot be tile-split again");
    return BN_ERR_INVALID;
  }
  const size_t len = (size_t)p->width * p->height * 3;
  std::vector<int> rcs(n, BN_OK);
  std::vector<std::string> errs(n);
  std::vector<BnStats> st(n);
  auto work = [&](int k) {
    BnStats& s = st[k];
    s = BnStats{};
    if ((rcs[k] = bnint::scene_film(m->scenes[k], len, &m->films[k])) != BN_OK) { errs[k] = bn_last_error(); return; }
    BnRenderParams sp;
    if (shard_params(*p, partition, n, k, sp)) {
      rcs[k] = bnint::render_on_stream(m->scenes[k], &sp, m->films[k], m->streams[k], &s);  // synchronises its stream
      if (rcs[k] != BN_OK) errs[k] = bn_last_error();
    } else if (cudaMemsetAsync(m->films[k], 0, len * sizeof(float), m->streams[k]) != cudaSuccess || cudaStreamSynchronize(m->streams[k]) != cudaSuccess) {
      rcs[k] = BN_ERR_CUDA; errs[k] = "clearing an empty share's film failed";
    }
  };
  {
    std::vector<std::thread> th;
    for (int k = 1; k < n; ++k) th.emplace_back(work, k);
    work(0);
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < n; ++k)
    if (rcs[k] != BN_OK) { bnhost::set_error("device " + std::to_string(m->devices[k]) + ": " + errs[k]); return rcs[k]; }
  // ---- combine on the first device
  if (!cuda_ok(cudaSetDevice(m->devices[0]), "cudaSetDevice")) return BN_ERR_CUDA;
  cudaStream_t s0 = m->streams[0];
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->devices[0]);
  if (!cuda_ok(cudaEventRecord(m->ev0, s0), "cudaEventRecord")) return BN_ERR_CUDA;
  PeerFilms peers{};
  uint64_t reduce_launches = 0;
  auto flush = [&]() {
    if (peers.n == 0) return true;
    k_film_reduce<<<sms * 4, 256, 0, s0>>>(m->films[0], peers, len);
    ++reduce_launches;
    peers.n = 0;
    return cuda_ok(cudaGetLastError(), "k_film_reduce");
  };
  for (int k = 1; k < n; ++k) {
    if (m->direct[k]) {
      peers.p[peers.n++] = m->films[k];
      continue;
    }
    // no peer access: keep the device order of the additions — flush what is pending, land this film, add it
    if (!flush()) return BN_ERR_CUDA;
    if (m->staging_len < len) {
      if (m->staging) cudaFree(m->staging);
      m->staging = nullptr; m->staging_len = 0;
      if (!cuda_ok(cudaMalloc((void**)&m->staging, len * sizeof(float)), "cudaMalloc(staging)")) return BN_ERR_CUDA;
      m->staging_len = len;
    }
    if (!cuda_ok(cudaMemcpyPeerAsync(m->staging, m->devices[0], m->films[k], m->devices[k], len * sizeof(float), s0), "cudaMemcpyPeerAsync")) return BN_ERR_CUDA;
    peers.p[peers.n++] = m->staging;
    if (!flush()) return BN_ERR_CUDA;
  }
  if (!flush()) return BN_ERR_CUDA;
  if (!cuda_ok(cudaEventRecord(m->ev1, s0), "cudaEventRecord")) return BN_ERR_CUDA;
  if (!cuda_ok(cudaMemcpyAsync(film_rgb, m->films[0], len * sizeof(float), cudaMemcpyDeviceToHost, s0), "film -> host")) return BN_ERR_CUDA;
  if (!cuda_ok(cudaStreamSynchronize(s0), "bn_render_multi: combine")) return BN_ERR_CUDA;
  if (stats) {
    BnStats t{};
    float reduce_ms = 0.f;
    cudaEventElapsedTime(&reduce_ms, m->ev0, m->ev1);
    for (int k = 0; k < n; ++k) {
      t.paths += st[k].paths; t.extend_rays += st[k].extend_rays; t.shadow_rays += st[k].shadow_rays; t.shadow_rays_ref += st[k].shadow_rays_ref;
      t.kernel_launches += st[k].kernel_launches;
      // device time of the call = the slowest device (they run side by side) + the combine; per-class times likewise
      t.gpu_ms = std::max(t.gpu_ms, st[k].gpu_ms); t.extend_ms = std::max(t.extend_ms, st[k].extend_ms); t.shade_ms = std::max(t.shade_ms, st[k].shade_ms);
      t.shadow_ms = std::max(t.shadow_ms, st[k].shadow_ms); t.other_ms = std::max(t.other_ms, st[k].other_ms);
    }
    t.gpu_ms += reduce_ms;
    t.other_ms += reduce_ms;
    t.kernel_launches += reduce_launches;
    *stats = t;
  }
  return BN_OK;
}

}  // extern "C"
