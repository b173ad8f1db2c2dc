// TEST INFRASTRUCTURE — runs traverse.cuh's warp-synchronous persistent loop (traverse_persistent: phase votes, stay
// loops, lane refill, small-TLAS scan, identity-instance shortcut, deferral to the exact path) ON THE HOST, as written.
// The 32 lanes of a warp are fibers (ucontext) of one thread; every warp intrinsic is a rendezvous: a lane that reaches
// one yields, and when all 32 wait at the same kind of intrinsic the results are computed and the lanes resume.  That is
// the execution model the kernel is written for (full-mask *_sync intrinsics reached by all lanes in the same order), so
// a scheduling bug that would hang or corrupt a real warp shows up here as a mismatched rendezvous.
// tests/test_hostsim.py compares the hits with the oracle bit for bit.  Never linked into the product.
#define BN_HOSTSIM_WARP 1
#include "device_shim.h"

#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../barnacle_b200/csrc/cuda/scene_convert.h"
#include "../../barnacle_b200/csrc/cuda/traverse.cuh"

namespace {
constexpr int kLanes = 32;
constexpr size_t kFiberStack = 256 * 1024;

struct Warp {
  ucontext_t scheduler;
  ucontext_t lane[kLanes];
  std::vector<char> stack[kLanes];
  bool done[kLanes];
  int kind[kLanes];
  unsigned in[kLanes], out[kLanes];
  int src[kLanes];
  int running = -1;
  uint64_t rendezvous = 0;
  unsigned work[kLanes] = {};   // BN_WORK units since the last rendezvous, per lane
  uint64_t simt_cost = 0;       // sum over rendezvous intervals of the maximum over lanes (+ the intrinsic itself)
  std::string error;
};
Warp* g_warp = nullptr;

struct HsWarpScene {
  bnconv::ConvertedScene cs;
  bn::DScene d;
};

// the IO concept of traverse_persistent (traverse.cuh), over host arrays
struct HostIO {
  const BnRay* rays;
  bn::TraceResult* results;
  int n;
  int* cur;
  std::vector<int>* deferred;
  int count() const { return n; }
  int* cursor() const { return cur; }
  void load(int i, float3& o, float3& d, float& t) const {
    o = bn::f3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
    d = bn::f3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
    t = rays[i].tmax;
  }
  void store(int i, const bn::TraceResult& r) const { results[i] = r; }
  void defer(int i) const { deferred->push_back(i); }
  void prefetch(int) const {}
};

struct Job {
  const bn::DScene* sc;
  HostIO* io;
  uint32_t* cold;
  bool any;
};
Job g_job;

void lane_main(int lane) {
  const bool wide = g_job.sc->wide != nullptr;   // as launch_traverse (kernels.cu)
  if (g_job.any) { if (wide) bn::traverse_persistent<true, true>(*g_job.sc, *g_job.io, g_job.cold + lane, kLanes); else bn::traverse_persistent<true, false>(*g_job.sc, *g_job.io, g_job.cold + lane, kLanes); }
  else { if (wide) bn::traverse_persistent<false, true>(*g_job.sc, *g_job.io, g_job.cold + lane, kLanes); else bn::traverse_persistent<false, false>(*g_job.sc, *g_job.io, g_job.cold + lane, kLanes); }
  g_warp->done[lane] = true;
  swapcontext(&g_warp->lane[lane], &g_warp->scheduler);
}

bool run_warp(Warp& w) {
  g_warp = &w;
  for (int l = 0; l < kLanes; ++l) {
    w.stack[l].assign(kFiberStack, 0);
    w.done[l] = false;
    w.kind[l] = -1;
    getcontext(&w.lane[l]);
    w.lane[l].uc_stack.ss_sp = w.stack[l].data();
    w.lane[l].uc_stack.ss_size = kFiberStack;
    w.lane[l].uc_link = &w.scheduler;
    makecontext(&w.lane[l], (void (*)())lane_main, 1, l);
  }
  for (;;) {
    int n_done = 0;
    for (int l = 0; l < kLanes; ++l) {
      if (w.done[l]) { ++n_done; continue; }
      w.running = l;
      swapcontext(&w.scheduler, &w.lane[l]);  // runs lane l up to its next rendezvous (or to its end)
      if (w.done[l]) ++n_done;
    }
    if (n_done == kLanes) return true;
    if (n_done != 0) { w.error = "a lane left the loop while others wait at a full-mask intrinsic (a real warp would hang)"; return false; }
    for (int l = 1; l < kLanes; ++l)
      if (w.kind[l] != w.kind[0]) { w.error = "lanes wait at different warp intrinsics (divergent *_sync)"; return false; }
    ++w.rendezvous;
    {
      unsigned mx = 0;
      for (int l = 0; l < kLanes; ++l) { if (w.work[l] > mx) mx = w.work[l]; w.work[l] = 0; }
      w.simt_cost += mx + (w.kind[0] == 0 ? 30u : 6u);  // a vote with its unpacking / a ballot + popc + branch
    }
    if (w.kind[0] == 0) {
      unsigned s = 0;
      for (int l = 0; l < kLanes; ++l) s += w.in[l];
      for (int l = 0; l < kLanes; ++l) w.out[l] = s;
    } else if (w.kind[0] == 1) {
      unsigned m = 0;
      for (int l = 0; l < kLanes; ++l) m |= (w.in[l] ? 1u : 0u) << l;
      for (int l = 0; l < kLanes; ++l) w.out[l] = m;
    } else {
      for (int l = 0; l < kLanes; ++l) w.out[l] = w.in[w.src[l] & 31];
    }
  }
}
}  // namespace

unsigned hostsim_lane_id(void) { return (unsigned)g_warp->running; }
void hostsim_work(unsigned units) { g_warp->work[g_warp->running] += units; }
unsigned hostsim_warp_sync(int kind, unsigned value, int src_lane) {
  Warp& w = *g_warp;
  const int l = w.running;
  w.kind[l] = kind; w.in[l] = value; w.src[l] = src_lane;
  swapcontext(&w.lane[l], &w.scheduler);
  w.running = l;
  return w.out[l];
}

extern "C" {

static std::string g_err;
const char* hsw_last_error(void) { return g_err.c_str(); }

// use_flat_tlas: bit 0 = the small-TLAS ordered scan (as bn_scene_create without BN_NO_FLAT_TLAS); bit 1 = binary nodes (BN_BINARY_NODES)
void* hsw_scene_create(const BnSceneDesc* desc, int use_flat_tlas) {
  auto* s = new HsWarpScene();
  if (!bnconv::convert_scene(*desc, s->cs, g_err)) { delete s; return nullptr; }
  if (use_flat_tlas & 2) bnconv::use_binary_nodes(s->cs);
  use_flat_tlas &= 1;
  bn::DScene& d = s->d;
  const bnconv::ConvertedScene& cs = s->cs;
  d.nodes = cs.nodes.data(); d.inst_trav = cs.inst_trav.data(); d.inst_head = cs.inst_head.data(); d.inst_w2o = cs.inst_w2o.data();
  d.inst_o2w = cs.inst_o2w.data(); d.meshes = cs.meshes.data(); d.tris = cs.tris.data(); d.alias = cs.alias.data();
  d.sphere_radii = cs.sphere_radii.data(); d.materials = cs.materials.data(); d.lights = cs.lights.data(); d.light_inst = cs.light_inst.data();
  d.flat_tlas = (use_flat_tlas && !cs.flat_tlas.empty()) ? cs.flat_tlas.data() : nullptr;  // as bn_scene_create (BN_NO_FLAT_TLAS switches it off)
  d.wide = cs.wide.empty() ? nullptr : cs.wide.data();
  d.tlas_wroot = cs.tlas_wroot;
  d.tlas = cs.tlas;
  d.n_inst = (uint32_t)cs.inst_head.size();
  d.n_light_inst = (uint32_t)cs.light_inst.size();
  d.all_finite = cs.all_finite ? 1u : 0u;
  d.cam = cs.cam;
  return s;
}
int hsw_has_wide(void* h) { return static_cast<HsWarpScene*>(h)->d.wide != nullptr; }
void hsw_scene_destroy(void* h) { delete static_cast<HsWarpScene*>(h); }
int hsw_has_flat_tlas(void* h) { return static_cast<HsWarpScene*>(h)->d.flat_tlas != nullptr; }

// One emulated warp drains the whole batch through traverse_persistent, then the deferred rays go through trace_exact
// (the fix-up kernel).  out_stats (may be NULL, 3 values): rendezvous count, deferred rays, SIMT cost (BN_TRAV_STATS builds).
int hsw_trace(void* h, const BnRay* rays, uint64_t n, int any_hit, BnHit* hits, uint64_t* out_stats) {
  const bn::DScene& sc = static_cast<HsWarpScene*>(h)->d;
  std::vector<bn::TraceResult> results((size_t)n);
  std::vector<int> deferred;
  int cursor = 0;
  HostIO io{rays, results.data(), (int)n, &cursor, &deferred};
  std::vector<uint32_t> cold((size_t)bn::kTravColdWords * kLanes, 0u);
  g_job = Job{&sc, &io, cold.data(), any_hit != 0};
  Warp w;
  if (!run_warp(w)) { g_err = w.error; return -1; }
  for (int i : deferred) {  // k_traverse_fixup
    float3 o, d;
    float t;
    io.load(i, o, d, t);
    if (any_hit) bn::trace_exact<true>(sc, o, d, t, results[i]);
    else bn::trace_exact<false>(sc, o, d, t, results[i]);
  }
  for (uint64_t i = 0; i < n; ++i) {  // TraceIO::store (kernels.cu), minus the sphere uv
    const bn::TraceResult& r = results[i];
    BnHit out;
    if (any_hit) {
      out.t = 0.f; out.u = 0.f; out.v = 0.f; out.instance = r.hit ? 1 : 0; out.primitive = 0;
    } else {
      out.t = r.t; out.u = r.u; out.v = r.v; out.instance = r.inst; out.primitive = r.prim;
      if (r.hit) {
        const bool sphere = sc.inst_trav[r.inst].is_sphere != 0u;
        out.primitive = sphere ? 0 : r.prim - (int)sc.inst_trav[r.inst].tri_base;
        if (sphere) { out.u = 0.f; out.v = 0.f; }
      }
    }
    hits[i] = out;
  }
  if (out_stats) { out_stats[0] = w.rendezvous; out_stats[1] = deferred.size(); out_stats[2] = w.simt_cost; }
  return 0;
}

#ifdef BN_TRAV_STATS
// phase statistics of the emulated warp (tools/warp_stats.py): [any * 10 + k] = executions of phase k (0 N, 1 T, 2 E,
// 3 all phases, 4 S), [any * 10 + 5 + k] = lanes that were ready in them — the counters the --stats GPU build keeps
void hsw_stats(unsigned long long* out, int reset) {
  for (int k = 0; k < 24; ++k) { out[k] = bn::g_trav_stats[k]; if (reset) bn::g_trav_stats[k] = 0; }
}
#endif

}  // extern "C"
