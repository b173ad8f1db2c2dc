// PSSMLT on the GPU ("next" row N1): PSSMLTIntegrator.Render, Extensions/Integrator/PSSMLT.fs:379-414.
//
// Phase 1 (bootstrap, :247-273) is embarrassingly parallel: one thread per bootstrap path.
// Phase 2 (chains, :275-377) is a set of sequential Markov chains: one thread per chain, the
// whole mutate -> trace -> accept/reject loop inside the thread ("megakernel"), film splats as
// atomic adds (the reference's Film.Accumulate is a racy read-modify-write).
// Parallelism is therefore the chain count: the reference's default of 1024 chains leaves a
// B200 mostly idle; scenes meant for the GPU should ask for >= 10^5 chains (`n-chains`).
// Every ray goes through trace_lane (per-lane traversal, the op-for-op exact form; a fast-slab
// variant exists behind BN_MLT_FAST_TRACE, bit-identical but slower here — see traverse.cuh), every
// transcendental through include/bn_portable_math.h, so a chain evolves bit-identically to the oracle's.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/barnacle_b200.h"
#include "scene_internal.h"
#include "shade.cuh"
#include "traverse.cuh"
#include "vecmath.cuh"

namespace bnhost {
void set_error(const std::string& msg);
}

namespace bn {

struct MltParams {
  int width, height, max_depth, rr_depth, frame_id, n_bootstrap, n_chains, strategy;
  float p0, p1, large_step_prob;
  int chain_begin, chain_end, mutation_per_chain;
  float inv_eff, inv_b;
  int n_threads;   // stride of the primary-sample arrays
  int xs_len;      // 4 + 7 * max_depth
};

// PrimarySample arrays, one column per thread (coalesced): value, backup, lastMod, modBackup
struct MltState {
  float* value;
  float* backup;
  int* last_mod;
  int* mod_backup;
};

struct MltSampler {  // PSSMLT.fs:35-150
  uint32_t inner;    // Sampler.state
  bool large_step;
  int last_large_step_iteration, current_iteration, sample_index, initialized;
};

BN_DEV float erf_inv(float x) {  // PSSMLT.fs:68-96
  x = net_min(net_max(x, -0.99999f), 0.99999f);
  float w = -bn_logf(__fmaf_rn(x, -x, 1.f));
  if (w < 5.f) {
    w = w - 2.5f;
    float p = 2.81022636e-08f;
    p = __fmaf_rn(p, w, 3.43273939e-07f);
    p = __fmaf_rn(p, w, -3.5233877e-06f);
    p = __fmaf_rn(p, w, -4.39150654e-06f);
    p = __fmaf_rn(p, w, 0.00021858087f);
    p = __fmaf_rn(p, w, -0.00125372503f);
    p = __fmaf_rn(p, w, -0.00417768164f);
    p = __fmaf_rn(p, w, 0.246640727f);
    return __fmaf_rn(p, w, 1.50140941f) * x;
  }
  w = __fsqrt_rn(w) - 3.f;
  float p = -0.000200214257f;
  p = __fmaf_rn(p, w, 0.000100950558f);
  p = __fmaf_rn(p, w, 0.00134934322f);
  p = __fmaf_rn(p, w, -0.00367342844f);
  p = __fmaf_rn(p, w, 0.00573950773f);
  p = __fmaf_rn(p, w, -0.0076224613f);
  p = __fmaf_rn(p, w, 0.00943887047f);
  p = __fmaf_rn(p, w, 1.00167406f);
  return __fmaf_rn(p, w, 2.83297682f) * x;
}

struct MltCtx {
  const MltParams& p;
  const MltState& st;
  int tid;
  MltSampler m;

  BN_DEV void init(uint32_t seed_state) {
    m.inner = seed_state; m.large_step = false;
    m.last_large_step_iteration = 0; m.current_iteration = 0; m.sample_index = 0; m.initialized = 0;
  }
  BN_DEV void start_iteration() {  // :57-60
    m.large_step = m.current_iteration == 0 || lcg(m.inner) < p.large_step_prob;
    m.current_iteration++;
    m.sample_index = 0;
  }
  BN_DEV float next1d() {  // EnsureReady(GetNextIndex()), :62-138
    const int index = m.sample_index++;
    const size_t at = (size_t)index * p.n_threads + tid;
    float value;
    int last_mod;
    if (m.initialized <= index) { value = 0.f; last_mod = 0; m.initialized = index + 1; }
    else { value = st.value[at]; last_mod = st.last_mod[at]; }
    if (last_mod < m.last_large_step_iteration) { value = lcg(m.inner); last_mod = m.last_large_step_iteration; }
    st.backup[at] = value; st.mod_backup[at] = last_mod;  // BackUp
    float v;
    if (m.large_step) {
      v = lcg(m.inner);
    } else if (p.strategy == BN_MLT_GAUSSIAN) {
      const float normal_sample = __fsqrt_rn(2.f) * erf_inv(__fmaf_rn(2.f, lcg(m.inner), -1.f));
      const float effective_sigma = p.p0 * __fsqrt_rn((float)(m.current_iteration - last_mod));
      v = __fmaf_rn(normal_sample, effective_sigma, value);
    } else {  // Kelemen(epsMin = p0, epsMax = p1)
      v = value;
      const float a = bn_logf(p.p1 / p.p0);
      for (int k = last_mod; k <= m.current_iteration - 1; ++k) {
        const float u1 = lcg(m.inner) - 0.5f;
        const float u2 = u1 < 0.f ? 1.f + 2.f * u1 : 2.f * u1;
        v = v + copysignf(p.p1 * bn_expf(-a * u2), u1);
      }
    }
    v = v - floorf(v);
    st.value[at] = v; st.last_mod[at] = m.current_iteration;
    return v;
  }
  BN_DEV void reject() {  // :142-146
    for (int i = 0; i < m.initialized; ++i) {
      const size_t at = (size_t)i * p.n_threads + tid;
      st.value[at] = st.backup[at]; st.last_mod[at] = st.mod_backup[at];
    }
    m.current_iteration--;
  }
  BN_DEV void accept() { if (m.large_step) m.last_large_step_iteration = m.current_iteration; }
};

// Interaction of a closest hit, rebuilt as the intersection routines produce it
// (same code path as k_shade; Primitive.fs:57-58, Mesh.fs:76-78, Sphere.fs:50-75).
struct Surface {
  float3 P;
  Onb onb;
  int material, light;
  bool is_sphere;
  float aux;  // mesh: MeshInstance.EvalPDF for tag 0 | sphere: radius
  Mat43 W2O;
};
BN_DEV Surface rebuild_surface(const DScene& sc, float3 o, float3 d, const TraceResult& h) {
  Surface s;
  const float4* hp = reinterpret_cast<const float4*>(sc.inst_head + h.inst);
  const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
  const uint32_t kind_prim = __float_as_uint(h0.w);
  s.material = __float_as_int(h1.w); s.light = __float_as_int(h2.x); s.aux = h2.y;
  s.W2O = load_mat43(reinterpret_cast<const float4*>(sc.inst_w2o + h.inst));
  const Mat43 O2W = load_mat43(reinterpret_cast<const float4*>(sc.inst_o2w + h.inst));
  const float3 oo = transform_point(o, s.W2O), od = transform_dir(d, s.W2O);
  const float3 pobj = point_at(oo, od, h.t);
  float3 nobj;
  s.is_sphere = (kind_prim & 0x80000000u) != 0u;
  if (s.is_sphere) {
    nobj = normalize(pobj);
    if (h.prim == 0 && dot(nobj, od) > 0.f) nobj = -nobj;
  } else {
    float3 p0, p1, p2;
    load_tri(sc.tris + h.prim, p0, p1, p2);
    nobj = normalize(cross(p1 - p0, p2 - p0));
  }
  s.P = transform_point(pobj, O2W);
  s.onb = transform_onb(onb_from_n(nobj), O2W);
  return s;
}

// PSSMLTIntegrator.Li — PSSMLT.fs:172-245 (fixed 7 dimensions per bounce)
BN_DEV float3 mlt_li(const DScene& sc, float3 o, float3 d, MltCtx& ctx, unsigned long long& rays) {
  const MltParams& p = ctx.p;
  float3 L = splat(0.f), beta = splat(1.f);
  float prev_pdf = 0.f;
  int depth = 0;
  while (depth < p.max_depth) {
    TraceResult h;
    trace_lane<false>(sc, o, d, CUDART_INF_F, h);
    ++rays;
    if (!h.hit) break;
    const Surface sf = rebuild_surface(sc, o, d, h);
    if (sf.light >= 0) {  // :187-198 + UniformLightSampler.Eval
      const float3 wo = normalize(o - sf.P);
      const float cos_wo = dot(sf.onb.n, wo);
      float pdf_surface = sf.aux;
      if (sf.is_sphere) {
        const float j = length(cross(transform_dir(sf.onb.t, sf.W2O), transform_dir(sf.onb.b, sf.W2O)));
        pdf_surface = j / (4.f * kPi * sf.aux * sf.aux);
      }
      const float dist2 = length_sq(o - sf.P);
      const float3 Le = light_eval(load_light(sc, sf.light), dot(wo, sf.onb.n));
      const float lpdf = dist2 * pdf_surface / (net_max(fabsf(cos_wo), 1e-6f) * (float)sc.n_light_inst);
      const float w = depth == 0 ? 1.f : prev_pdf * (1.f / (lpdf + prev_pdf));
      L = vfma(beta, Le * w, L);
    }
    const float u_light = ctx.next1d();
    const float u_emit_x = ctx.next1d(), u_emit_y = ctx.next1d();
    const float u_lobe = ctx.next1d();
    const float u_bsdf_x = ctx.next1d(), u_bsdf_y = ctx.next1d();
    const float u_rr = ctx.next1d();
    if (sf.material < 0) break;
    const GMaterial mat = load_material(sc, sf.material);
    const LightSampleRec ls = light_sampler_sample(sc, sf.P, u_light, u_emit_x, u_emit_y);
    const float dist = length(ls.p - sf.P);
    const float3 wo_l = world_to_local(sf.onb, -d);
    if (ls.pdf != 0.f) {
      TraceResult sh;
      trace_lane<true>(sc, sf.P, ls.wi, dist - 1e-3f, sh);
      ++rays;
      if (!sh.hit) {
        const BsdfEval fe = material_eval(mat, wo_l, world_to_local(sf.onb, ls.wi));
        L = vfma(beta * fe.bsdf, ls.L * (1.f / (fe.pdf + ls.pdf)), L);
      }
    }
    const BsdfSample bs = material_sample(mat, wo_l, u_lobe, u_bsdf_x, u_bsdf_y);
    prev_pdf = bs.eval.pdf;
    if (bs.eval.pdf == 0.f) break;
    o = sf.P;
    d = local_to_world(sf.onb, bs.wi);
    beta = beta * bs.eval.bsdf * (1.f / bs.eval.pdf);
    if (depth >= p.rr_depth) {
      const float q = net_min(1.f, net_max(beta.x, net_max(beta.y, beta.z)));
      if (u_rr < q) beta = beta * (1.f / q);
      else break;
    }
    depth++;
  }
  return L;
}

BN_DEV float luminance(float3 L) { return dot(L, f3(0.2126f, 0.7152f, 0.0722f)); }

// pixel from two primary samples, camera ray from two more, Li (PSSMLT.fs:254-269 / 284-300 / 332-349)
BN_DEV float3 mlt_sample_path(const DScene& sc, MltCtx& ctx, int& px, int& py, unsigned long long& rays) {
  const MltParams& p = ctx.p;
  const float ux = ctx.next1d(), uy = ctx.next1d();
  const float upx = ux * (float)p.width, upy = uy * (float)p.height;
  px = min(p.width - 1, (int)upx);
  py = min(p.height - 1, (int)upy);
  const float ulx = ctx.next1d(), uly = ctx.next1d();
  float3 o, d;
  primary_ray(sc.cam, p.width, p.height, px, py, upx - (float)px, upy - (float)py, ulx, uly, o, d);
  return mlt_li(sc, o, d, ctx, rays) * (1.f / 1.f);
}

BN_DEV uint32_t xxhash32_two(uint32_t x, uint32_t y) {  // Hash.fs:6-15
  const uint32_t p2 = 2246822519u, p3 = 3266489917u, p4 = 668265263u, p5 = 374761393u;
  uint32_t h = y + p5 + x * p3;
  h = p4 * rotl17(h);
  h = p2 * (h ^ (h >> 15));
  h = p3 * (h ^ (h >> 13));
  return h ^ (h >> 16);
}

__global__ void __launch_bounds__(128) k_mlt_bootstrap(DScene sc, MltParams p, MltState st, float* __restrict__ weights, unsigned long long* ray_count) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long rays = 0;
  for (int id = tid; id < p.n_bootstrap; id += p.n_threads) {
    MltCtx ctx{p, st, tid, {}};
    ctx.init(xxhash32_two((uint32_t)p.frame_id, (uint32_t)id));
    ctx.start_iteration();
    int px, py;
    const float3 L = mlt_sample_path(sc, ctx, px, py, rays);
    weights[id] = luminance(L);
  }
  if (rays) atomicAdd(ray_count, rays);
}

BN_DEV void film_splat(float* film, int W, int H, int px, int py, float3 c) {  // Film.Accumulate, Film.fs:48-53 (atomic here)
  float* d = film + ((size_t)(H - py - 1) * W + px) * 3;
  atomicAdd(d, c.x); atomicAdd(d + 1, c.y); atomicAdd(d + 2, c.z);
}

__global__ void __launch_bounds__(128) k_mlt_chains(DScene sc, MltParams p, MltState st, float* __restrict__ film, unsigned int* __restrict__ per_chain_accepted,
                                                    unsigned long long* counters /* [0] rays [1] accepted [2] proposed */) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = p.chain_begin + tid;
  if (chain >= p.chain_end) return;
  unsigned long long rays = 0;
  uint32_t sampler = xxhash32_two((uint32_t)p.frame_id, (uint32_t)chain);  // Sampler(FrameId, chainId), :285
  // AliasTable(BootstrapWeights).Sample never takes an alias: a uniform pick (:286)
  const float u = lcg(sampler) * (float)p.n_bootstrap;
  const int bootstrap_id = min((int)u, p.n_bootstrap - 1);
  MltCtx ctx{p, st, tid, {}};
  ctx.init(xxhash32_two((uint32_t)p.frame_id, (uint32_t)bootstrap_id));
  ctx.start_iteration();
  int px, py;
  float3 L = mlt_sample_path(sc, ctx, px, py, rays);
  float y = luminance(L);
  ctx.accept();
  ctx.m.inner = xxhash32_three((uint32_t)chain, (uint32_t)bootstrap_id, (uint32_t)p.frame_id);  // :307
  float3 radiance = splat(0.f);
  unsigned int accepted = 0;
  for (int k = 0; k < p.mutation_per_chain; ++k) {
    ctx.start_iteration();
    int nx, ny;
    const float3 Ln = mlt_sample_path(sc, ctx, nx, ny, rays);
    const float yn = luminance(Ln);
    const float a = net_min(1.f, yn / y);
    const float w_old = (1.f - a) / __fmaf_rn(y, p.inv_b, p.large_step_prob);
    radiance = radiance + w_old * L;
    const float w_new = (a + (ctx.m.large_step ? 1.f : 0.f)) / __fmaf_rn(yn, p.inv_b, p.large_step_prob);
    if (lcg(sampler) < a) {
      ++accepted;
      film_splat(film, p.width, p.height, px, py, radiance * p.inv_eff);
      radiance = w_new * Ln;
      px = nx; py = ny; L = Ln; y = yn;
      ctx.accept();
    } else {
      if (a > 0.f) film_splat(film, p.width, p.height, nx, ny, (w_new * p.inv_eff) * Ln);
      ctx.reject();
    }
  }
  film_splat(film, p.width, p.height, px, py, radiance * p.inv_eff);
  if (per_chain_accepted) per_chain_accepted[tid] = accepted;
  atomicAdd(counters, rays);
  atomicAdd(counters + 1, (unsigned long long)accepted);
  atomicAdd(counters + 2, (unsigned long long)p.mutation_per_chain);
}


// =====================================================================================================================
// PSSMLT as a WAVEFRONT over chains (round 2).  The megakernels above walk one path per thread with the exact per-lane
// traversal; here a mutation round is one wave through the path tracer's own stages: every chain proposes
// (k_mlt_begin), the proposals' rays go through the warp-synchronous traversal kernels of traverse.cuh (4-wide nodes,
// fast slabs, exact fix-up) bounce by bounce, k_mlt_shade is PSSMLTIntegrator.Li's loop body (PSSMLT.fs:172-245: seven
// primary-sample dimensions per bounce whatever the surface) between the queues, and k_mlt_end is the accept / reject
// step with its splats (:332-377).  The arithmetic per chain is the megakernel's, operation for operation — same
// bootstrap weights, same B, same accepted count per chain — only the order of the atomic film splats differs.
// The bootstrap (:247-273) is the same wave with 4 Mi independent paths.
// =====================================================================================================================
constexpr int kMltCS = 64;  // ints between two counters (256 B: same-line atomics serialise in L2, see kernels.cu)

struct MltWave {
  // path-tracer queues (the scene's wave buffers)
  float4 *a0, *a1, *a2, *b0, *b1, *b2, *hits, *q0, *q1, *q2, *q3, *rad;
  // per chain / bootstrap lane, SoA with stride n
  int* samp;       // [6][n]: MltSampler {inner, large_step, last_large_step_iteration, current_iteration, sample_index, initialized}
  float* chain_f;  // [8][n]: L.xyz, y, radiance.xyz, -
  int* chain_i;    // [6][n]: accept rng, px, py, accepted, proposal px, proposal py (bootstrap id during the first path)
  int* counters;   // per round, re-zeroed by k_mlt_begin: n_active[D+1] | n_shadow[D] | cursors[3D] | n_defer[2D], kMltCS apart
  unsigned long long* shadow_ref;  // reference-equivalent shadow rays of the round
  unsigned long long* totals;      // [0] rays [1] accepted [2] proposed
  int n;           // lanes of this wave
};

BN_DEV void mlt_load_sampler(const MltWave& w, int pid, MltSampler& m) {
  m.inner = (uint32_t)w.samp[pid]; m.large_step = w.samp[w.n + pid] != 0; m.last_large_step_iteration = w.samp[2 * w.n + pid];
  m.current_iteration = w.samp[3 * w.n + pid]; m.sample_index = w.samp[4 * w.n + pid]; m.initialized = w.samp[5 * w.n + pid];
}
BN_DEV void mlt_store_sampler(const MltWave& w, int pid, const MltSampler& m) {
  w.samp[pid] = (int)m.inner; w.samp[w.n + pid] = m.large_step ? 1 : 0; w.samp[2 * w.n + pid] = m.last_large_step_iteration;
  w.samp[3 * w.n + pid] = m.current_iteration; w.samp[4 * w.n + pid] = m.sample_index; w.samp[5 * w.n + pid] = m.initialized;
}

// MODE 0: bootstrap path `id_base + lane` (PSSMLT.fs:250-258) | 1: a chain's first path (:284-300) | 2: a mutation (:330-349)
template <int MODE>
__global__ void __launch_bounds__(128) k_mlt_begin(DScene sc, MltParams p, MltState st, MltWave w, int id_base, int n_counter_ints) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  // this round's counters: everything to zero, n_active[0] = n (all lanes start a path)
  for (int k = gtid; k < n_counter_ints; k += gridDim.x * blockDim.x) w.counters[k] = k == 0 ? w.n : 0;
  if (gtid == 0) *w.shadow_ref = 0ull;
  for (int tid = gtid; tid < w.n; tid += gridDim.x * blockDim.x) {
    MltCtx ctx{p, st, tid, {}};
    if (MODE == 0) {
      ctx.init(xxhash32_two((uint32_t)p.frame_id, (uint32_t)(id_base + tid)));
    } else if (MODE == 1) {
      const int chain = p.chain_begin + tid;
      uint32_t sampler = xxhash32_two((uint32_t)p.frame_id, (uint32_t)chain);  // Sampler(FrameId, chainId), :285
      const float u = lcg(sampler) * (float)p.n_bootstrap;                     // AliasTable.Sample: a uniform pick
      const int bootstrap_id = min((int)u, p.n_bootstrap - 1);
      ctx.init(xxhash32_two((uint32_t)p.frame_id, (uint32_t)bootstrap_id));
      w.chain_i[tid] = (int)sampler;
      w.chain_i[5 * w.n + tid] = bootstrap_id;
    } else {
      mlt_load_sampler(w, tid, ctx.m);
    }
    ctx.start_iteration();
    // pixel from two primary samples, camera ray from two more (PSSMLT.fs:254-262)
    const float ux = ctx.next1d(), uy = ctx.next1d();
    const float upx = ux * (float)p.width, upy = uy * (float)p.height;
    const int px = min(p.width - 1, (int)upx), py = min(p.height - 1, (int)upy);
    const float ulx = ctx.next1d(), uly = ctx.next1d();
    float3 o, d;
    primary_ray(sc.cam, p.width, p.height, px, py, upx - (float)px, upy - (float)py, ulx, uly, o, d);
    w.a0[tid] = make_float4(o.x, o.y, o.z, d.x);
    w.a1[tid] = make_float4(d.y, d.z, 1.f, 1.f);
    w.a2[tid] = make_float4(1.f, 0.f, 0.f, __int_as_float(tid));
    w.rad[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 1) { w.chain_i[1 * w.n + tid] = px; w.chain_i[2 * w.n + tid] = py; }
    else if (MODE == 2) { w.chain_i[4 * w.n + tid] = px; w.chain_i[5 * w.n + tid] = py; }
    mlt_store_sampler(w, tid, ctx.m);
  }
}

// One iteration of PSSMLTIntegrator.Li's loop (PSSMLT.fs:176-243) for every live proposal: in = state A + hit records,
// out = state B (compacted) + shadow queue; L lives in rad[lane].
__global__ void __launch_bounds__(128) k_mlt_shade(DScene sc, MltParams p, MltState st, MltWave w, int bounce, const float4* __restrict__ s0,
                                                   const float4* __restrict__ s1, const float4* __restrict__ s2, float4* __restrict__ o0,
                                                   float4* __restrict__ o1, float4* __restrict__ o2, const int* __restrict__ n_ptr, int* n_out,
                                                   int* n_shadow, int* cursor) {
  const int n = *n_ptr;
  const int lane = threadIdx.x & 31;
  unsigned ref_total = 0;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(cursor, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n) break;
    const int i = base + lane;
    bool alive = false, has_shadow = false, ref_shadow = false;
    float3 P = splat(0.f), nd = splat(0.f), beta = splat(0.f), sh_wi = splat(0.f), sh_a = splat(0.f), sh_b = splat(0.f);
    float bs_pdf = 0.f, sh_tmax = 0.f;
    int pid = 0;
    if (i < n) {
      const float4 a = s0[i], b = s1[i], c = s2[i], h = w.hits[i];
      const float3 o = f3(a.x, a.y, a.z), d = f3(a.w, b.x, b.y);
      beta = f3(b.z, b.w, c.x);
      const float prev_pdf = c.y;
      pid = __float_as_int(c.w);
      TraceResult tr;
      tr.t = h.x; tr.inst = __float_as_int(h.y); tr.prim = __float_as_int(h.z); tr.hit = tr.inst >= 0; tr.u = 0.f; tr.v = 0.f;
      if (tr.hit) {
        const Surface sf = rebuild_surface(sc, o, d, tr);
        if (sf.light >= 0) {  // :187-198 + UniformLightSampler.Eval
          const float3 wo = normalize(o - sf.P);
          const float cos_wo = dot(sf.onb.n, wo);
          float pdf_surface = sf.aux;
          if (sf.is_sphere) {
            const float j = length(cross(transform_dir(sf.onb.t, sf.W2O), transform_dir(sf.onb.b, sf.W2O)));
            pdf_surface = j / (4.f * kPi * sf.aux * sf.aux);
          }
          const float dist2 = length_sq(o - sf.P);
          const float3 Le = light_eval(load_light(sc, sf.light), dot(wo, sf.onb.n));
          const float lpdf = dist2 * pdf_surface / (net_max(fabsf(cos_wo), 1e-6f) * (float)sc.n_light_inst);
          const float wgt = bounce == 0 ? 1.f : prev_pdf * (1.f / (lpdf + prev_pdf));
          const float4 L4 = w.rad[pid];
          const float3 L = vfma(beta, Le * wgt, f3(L4.x, L4.y, L4.z));
          w.rad[pid] = make_float4(L.x, L.y, L.z, 0.f);
        }
        MltCtx ctx{p, st, pid, {}};
        mlt_load_sampler(w, pid, ctx.m);
        const float u_light = ctx.next1d();  // the seven dimensions of a bounce, drawn whatever the surface (:201-205)
        const float u_emit_x = ctx.next1d(), u_emit_y = ctx.next1d();
        const float u_lobe = ctx.next1d();
        const float u_bsdf_x = ctx.next1d(), u_bsdf_y = ctx.next1d();
        const float u_rr = ctx.next1d();
        mlt_store_sampler(w, pid, ctx.m);
        if (sf.material >= 0) {
          const GMaterial mat = load_material(sc, sf.material);
          const LightSampleRec ls = light_sampler_sample(sc, sf.P, u_light, u_emit_x, u_emit_y);
          const float dist = length(ls.p - sf.P);
          const float3 wo_l = world_to_local(sf.onb, -d);
          P = sf.P;
          if (ls.pdf != 0.f) {
            ref_shadow = true;
            const BsdfEval fe = material_eval(mat, wo_l, world_to_local(sf.onb, ls.wi));
            sh_a = beta * fe.bsdf;
            sh_b = ls.L * (1.f / (fe.pdf + ls.pdf));
            // fma(0, finite, L) == L bit for bit: such a connection cannot change L and is not traced
            has_shadow = !(sh_a.x == 0.f && sh_a.y == 0.f && sh_a.z == 0.f && isfinite(sh_b.x) && isfinite(sh_b.y) && isfinite(sh_b.z));
            sh_wi = ls.wi;
            sh_tmax = dist - 1e-3f;
          }
          const BsdfSample bs = material_sample(mat, wo_l, u_lobe, u_bsdf_x, u_bsdf_y);
          bs_pdf = bs.eval.pdf;
          if (bs.eval.pdf != 0.f) {
            nd = local_to_world(sf.onb, bs.wi);
            beta = beta * bs.eval.bsdf * (1.f / bs.eval.pdf);
            bool cont = true;
            if (bounce >= p.rr_depth) {
              const float q = net_min(1.f, net_max(beta.x, net_max(beta.y, beta.z)));
              if (u_rr < q) beta = beta * (1.f / q);
              else cont = false;
            }
            alive = cont && (bounce + 1 < p.max_depth);
          }
        }
      }
    }
    const unsigned m_alive = __ballot_sync(0xffffffffu, alive);
    int pos = 0;
    if (lane == 0 && m_alive) pos = atomicAdd(n_out, __popc(m_alive));
    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m_alive & ((1u << lane) - 1u));
    if (alive) {
      o0[pos] = make_float4(P.x, P.y, P.z, nd.x);
      o1[pos] = make_float4(nd.y, nd.z, beta.x, beta.y);
      o2[pos] = make_float4(beta.z, bs_pdf, 0.f, __int_as_float(pid));
    }
    const unsigned m_sh = __ballot_sync(0xffffffffu, has_shadow);
    int spos = 0;
    if (lane == 0 && m_sh) spos = atomicAdd(n_shadow, __popc(m_sh));
    spos = __shfl_sync(0xffffffffu, spos, 0) + __popc(m_sh & ((1u << lane) - 1u));
    if (has_shadow) {
      w.q0[spos] = make_float4(P.x, P.y, P.z, sh_wi.x);
      w.q1[spos] = make_float4(sh_wi.y, sh_wi.z, sh_tmax, __int_as_float(pid));
      w.q2[spos] = make_float4(sh_a.x, sh_a.y, sh_a.z, sh_b.x);
      w.q3[spos] = make_float4(sh_b.y, sh_b.z, 0.f, 0.f);
    }
    ref_total += (unsigned)__popc(__ballot_sync(0xffffffffu, ref_shadow));
  }
  if (lane == 0 && ref_total) atomicAdd(w.shadow_ref, (unsigned long long)ref_total);
}

// MODE 0: BootstrapWeights[id] = luminance(L) (:263-269) | 1: the chain's start state (:301-309) | 2: accept / reject (:350-376)
template <int MODE>
__global__ void __launch_bounds__(128) k_mlt_end(MltParams p, MltState st, MltWave w, int id_base, float* __restrict__ weights, float* __restrict__ film, int D) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid == 0) {  // rays of the round: every extend ray + every shadow ray the reference traces
    unsigned long long r = *w.shadow_ref;
    for (int b = 0; b < D; ++b) r += (unsigned long long)w.counters[b * kMltCS];
    w.totals[0] += r;
  }
  unsigned long long acc_total = 0;
  for (int tid = gtid; tid < w.n; tid += gridDim.x * blockDim.x) {
    const float4 L4 = w.rad[tid];
    const float3 Ln = f3(L4.x, L4.y, L4.z) * (1.f / 1.f);
    if (MODE == 0) {
      weights[id_base + tid] = luminance(Ln);
      continue;
    }
    MltCtx ctx{p, st, tid, {}};
    mlt_load_sampler(w, tid, ctx.m);
    float* cf = w.chain_f;
    int* ci = w.chain_i;
    if (MODE == 1) {
      const int chain = p.chain_begin + tid;
      const int bootstrap_id = ci[5 * w.n + tid];
      cf[tid] = Ln.x; cf[w.n + tid] = Ln.y; cf[2 * w.n + tid] = Ln.z; cf[3 * w.n + tid] = luminance(Ln);
      cf[4 * w.n + tid] = 0.f; cf[5 * w.n + tid] = 0.f; cf[6 * w.n + tid] = 0.f;
      ci[3 * w.n + tid] = 0;
      ctx.accept();
      ctx.m.inner = xxhash32_three((uint32_t)chain, (uint32_t)bootstrap_id, (uint32_t)p.frame_id);  // :307
      mlt_store_sampler(w, tid, ctx.m);
      continue;
    }
    uint32_t sampler = (uint32_t)ci[tid];
    const int px = ci[w.n + tid], py = ci[2 * w.n + tid], nx = ci[4 * w.n + tid], ny = ci[5 * w.n + tid];
    const float3 L = f3(cf[tid], cf[w.n + tid], cf[2 * w.n + tid]);
    const float y = cf[3 * w.n + tid];
    float3 radiance = f3(cf[4 * w.n + tid], cf[5 * w.n + tid], cf[6 * w.n + tid]);
    const float yn = luminance(Ln);
    const float a = net_min(1.f, yn / y);
    const float w_old = (1.f - a) / __fmaf_rn(y, p.inv_b, p.large_step_prob);
    radiance = radiance + w_old * L;
    const float w_new = (a + (ctx.m.large_step ? 1.f : 0.f)) / __fmaf_rn(yn, p.inv_b, p.large_step_prob);
    if (lcg(sampler) < a) {
      ci[3 * w.n + tid] += 1;
      ++acc_total;
      film_splat(film, p.width, p.height, px, py, radiance * p.inv_eff);
      radiance = w_new * Ln;
      ci[w.n + tid] = nx; ci[2 * w.n + tid] = ny;
      cf[tid] = Ln.x; cf[w.n + tid] = Ln.y; cf[2 * w.n + tid] = Ln.z; cf[3 * w.n + tid] = yn;
      ctx.accept();
    } else {
      if (a > 0.f) film_splat(film, p.width, p.height, nx, ny, (w_new * p.inv_eff) * Ln);
      ctx.reject();
    }
    cf[4 * w.n + tid] = radiance.x; cf[5 * w.n + tid] = radiance.y; cf[6 * w.n + tid] = radiance.z;
    ci[tid] = (int)sampler;
    mlt_store_sampler(w, tid, ctx.m);
  }
  if (MODE == 2) {
    for (int off = 16; off > 0; off >>= 1) acc_total += __shfl_down_sync(0xffffffffu, acc_total, off);
    if ((threadIdx.x & 31) == 0 && acc_total) atomicAdd(w.totals + 1, acc_total);
  }
}

// after the last mutation: the pending radiance of every chain (:377), per-chain accepted counts, proposed total
__global__ void __launch_bounds__(128) k_mlt_finish(MltParams p, MltWave w, float* __restrict__ film, unsigned int* __restrict__ per_chain_accepted) {
  for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < w.n; tid += gridDim.x * blockDim.x) {
    const float3 radiance = f3(w.chain_f[4 * w.n + tid], w.chain_f[5 * w.n + tid], w.chain_f[6 * w.n + tid]);
    film_splat(film, p.width, p.height, w.chain_i[w.n + tid], w.chain_i[2 * w.n + tid], radiance * p.inv_eff);
    if (per_chain_accepted) per_chain_accepted[tid] = (unsigned int)w.chain_i[3 * w.n + tid];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) w.totals[2] += (unsigned long long)p.mutation_per_chain * (unsigned long long)w.n;
}

}  // namespace bn

using namespace bn;

namespace {

bool ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  bnhost::set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}
#define MLT_CUDA(call)                            \
  do {                                            \
    if (!ok((call), #call)) return BN_ERR_CUDA;   \
  } while (0)

int validate(const BnScene* s, const BnMltParams* p) {
  if (!s || !p || p->width <= 0 || p->height <= 0 || p->mutations_per_pixel <= 0 || p->max_depth < 0 || p->n_bootstrap <= 0 || p->n_chains <= 0 ||
      p->chain_begin < 0 || p->chain_end > p->n_chains || p->chain_begin > p->chain_end || (p->strategy != BN_MLT_GAUSSIAN && p->strategy != BN_MLT_KELEMEN)) {
    bnhost::set_error("bn_render_pssmlt: invalid BnMltParams");
    return BN_ERR_INVALID;
  }
  if (s->poisoned) { bnhost::set_error("scene is unusable after an earlier CUDA error"); return BN_ERR_CUDA; }
  if (s->d.n_light_inst == 0) { bnhost::set_error("LightSamplerBase: No light primitives found."); return BN_ERR_NO_LIGHT; }
  return BN_OK;
}

// Scratch owned by the scene and reused by every render (no allocation in the steady state).
int ensure_mlt_state(BnScene* s, size_t n_threads, int xs_len, MltState& st) {
  const size_t n = n_threads * (size_t)xs_len;
  if (s->mlt_len < n) {
    if (s->mlt_f) cudaFree(s->mlt_f);
    if (s->mlt_i) cudaFree(s->mlt_i);
    s->mlt_f = nullptr; s->mlt_i = nullptr; s->mlt_len = 0;
    MLT_CUDA(cudaMalloc((void**)&s->mlt_f, 2 * n * sizeof(float)));
    MLT_CUDA(cudaMalloc((void**)&s->mlt_i, 2 * n * sizeof(int)));
    s->mlt_len = n;
  }
  st.value = s->mlt_f; st.backup = s->mlt_f + n; st.last_mod = s->mlt_i; st.mod_backup = s->mlt_i + n;
  return BN_OK;
}
int ensure_mlt_misc(BnScene* s, size_t n_weights, size_t n_acc) {
  if (!s->mlt_cnt) MLT_CUDA(cudaMalloc((void**)&s->mlt_cnt, 4 * sizeof(unsigned long long)));
  if (s->mlt_w_len < n_weights) {
    if (s->mlt_w) cudaFree(s->mlt_w);
    if (s->mlt_w_host) cudaFreeHost(s->mlt_w_host);
    s->mlt_w = nullptr; s->mlt_w_host = nullptr; s->mlt_w_len = 0;
    MLT_CUDA(cudaMalloc((void**)&s->mlt_w, n_weights * sizeof(float)));
    MLT_CUDA(cudaMallocHost((void**)&s->mlt_w_host, n_weights * sizeof(float)));
    s->mlt_w_len = n_weights;
  }
  if (s->mlt_acc_len < n_acc) {
    if (s->mlt_acc) cudaFree(s->mlt_acc);
    s->mlt_acc = nullptr; s->mlt_acc_len = 0;
    MLT_CUDA(cudaMalloc((void**)&s->mlt_acc, n_acc * sizeof(unsigned int)));
    s->mlt_acc_len = n_acc;
  }
  return BN_OK;
}

MltParams device_params(const BnMltParams* p) {
  MltParams d{};
  d.width = p->width; d.height = p->height; d.max_depth = p->max_depth; d.rr_depth = p->rr_depth; d.frame_id = p->frame_id;
  d.n_bootstrap = p->n_bootstrap; d.n_chains = p->n_chains; d.strategy = p->strategy; d.p0 = p->p0; d.p1 = p->p1;
  d.large_step_prob = p->large_step_prob; d.chain_begin = p->chain_begin; d.chain_end = p->chain_end;
  d.xs_len = 4 + 7 * p->max_depth;
  return d;
}

// ---- wavefront driver -------------------------------------------------------------------------------------------------
struct WaveHost {
  MltWave w{};
  MltState st{};
  MltParams dp{};
  int D = 0;
  int n_counter_ints = 0;
  int grid = 0;
};

// Queues (the scene's wave buffers), per-lane sampler / chain arrays and the round's counters for `n` lanes.
int setup_wave(BnScene* s, const BnMltParams* p, int n, WaveHost& h) {
  h.dp = device_params(p);
  h.dp.n_threads = n;
  h.D = p->max_depth;
  int rc = bnint::ensure_wave(s, (size_t)n);
  if (rc != BN_OK) return rc;
  if ((rc = ensure_mlt_state(s, (size_t)n, h.dp.xs_len, h.st)) != BN_OK) return rc;
  const size_t need_i = (size_t)12 * n, need_f = (size_t)8 * n;
  if (s->mlt_wave_len < (size_t)n) {
    if (s->mlt_wave_i) cudaFree(s->mlt_wave_i);
    if (s->mlt_wave_f) cudaFree(s->mlt_wave_f);
    s->mlt_wave_i = nullptr; s->mlt_wave_f = nullptr; s->mlt_wave_len = 0;
    MLT_CUDA(cudaMalloc((void**)&s->mlt_wave_i, need_i * sizeof(int)));
    MLT_CUDA(cudaMalloc((void**)&s->mlt_wave_f, need_f * sizeof(float)));
    s->mlt_wave_len = (size_t)n;
  }
  h.n_counter_ints = ((h.D + 1) + h.D + 3 * h.D + 2 * h.D) * kMltCS;
  if (s->mlt_counters_len < (size_t)h.n_counter_ints + 16) {
    if (s->mlt_counters) cudaFree(s->mlt_counters);
    s->mlt_counters = nullptr; s->mlt_counters_len = 0;
    MLT_CUDA(cudaMalloc((void**)&s->mlt_counters, ((size_t)h.n_counter_ints + 16) * sizeof(int)));
    s->mlt_counters_len = (size_t)h.n_counter_ints + 16;
  }
  if (!s->mlt_cnt) MLT_CUDA(cudaMalloc((void**)&s->mlt_cnt, 4 * sizeof(unsigned long long)));
  const size_t cp = s->cap;
  MltWave& w = h.w;
  w.a0 = s->state[0]; w.a1 = s->state[0] + cp; w.a2 = s->state[0] + 2 * cp;
  w.b0 = s->state[1]; w.b1 = s->state[1] + cp; w.b2 = s->state[1] + 2 * cp;
  w.hits = s->hits; w.q0 = s->shq; w.q1 = s->shq + cp; w.q2 = s->shq + 2 * cp; w.q3 = s->shq + 3 * cp; w.rad = s->rad;
  w.samp = s->mlt_wave_i; w.chain_i = s->mlt_wave_i + (size_t)6 * n; w.chain_f = s->mlt_wave_f;
  w.counters = s->mlt_counters;
  w.shadow_ref = reinterpret_cast<unsigned long long*>(s->mlt_counters + h.n_counter_ints + (h.n_counter_ints & 1));  // 8-B aligned tail
  w.totals = s->mlt_cnt;
  w.n = n;
  h.grid = std::max(1, std::min(s->num_sms * 8, (n + 127) / 128));
  return BN_OK;
}

// The bounces of one round: (extend + fix-up, Li's loop body, shadow + connect + fix-up) x maxDepth.  Reads state A first.
void enqueue_bounces(BnScene* s, const WaveHost& h, cudaStream_t stream) {
  const MltWave& w = h.w;
  const int D = h.D, CS = kMltCS;
  int* n_active = w.counters;
  int* n_shadow = n_active + (D + 1) * CS;
  int* cursors = n_shadow + D * CS;
  int* n_defer = cursors + 3 * D * CS;
  float4 *A0 = w.a0, *A1 = w.a1, *A2 = w.a2, *B0 = w.b0, *B1 = w.b1, *B2 = w.b2;
  for (int b = 0; b < D; ++b) {
    bnint::launch_extend(s, stream, A0, A1, w.hits, n_active + b * CS, cursors + (3 * b) * CS, n_defer + (2 * b) * CS);
    k_mlt_shade<<<h.grid, 128, 0, stream>>>(s->d, h.dp, h.st, w, b, A0, A1, A2, B0, B1, B2, n_active + b * CS, n_active + (b + 1) * CS, n_shadow + b * CS,
                                            cursors + (3 * b + 1) * CS);
    bnint::launch_shadow(s, stream, w.q0, w.q1, w.q2, w.q3, w.rad, n_shadow + b * CS, cursors + (3 * b + 2) * CS, n_defer + (2 * b + 1) * CS);
    std::swap(A0, B0); std::swap(A1, B1); std::swap(A2, B2);
  }
}

// Which form runs a phase with `n` independent lanes.  A wave pays ~40 kernel boundaries per round of maxDepth bounces
// (measured on the B200: 1.5 ms per round at 65 536 lanes, 2.4 ms at 262 144), the per-thread megakernel pays divergence and the
// exact per-lane traversal (0.84 ms per mutation whatever the chain count beyond ~65 536): the wave wins from ~10^5 lanes up
// (bootstrap, 4 Mi paths: 36 -> 18 ms; chains at 262 144: 148 -> 96 ms; chains at 65 536: 134 -> 243 ms, so those stay per
// thread).  BN_MLT_MEGAKERNEL=1 / BN_MLT_WAVEFRONT=1 force one form (A/B, tests).
bool use_wavefront(int n) {
  if (std::getenv("BN_MLT_MEGAKERNEL")) return false;
  if (std::getenv("BN_MLT_WAVEFRONT")) return true;
  return n >= 131072;
}

// phase 1 as waves of up to 1 Mi independent paths
int run_bootstrap_wavefront(BnScene* s, const BnMltParams* p, cudaStream_t stream, const float** weights_host, uint64_t& rays, double& ms) {
  int rc = ensure_mlt_misc(s, (size_t)p->n_bootstrap, 0);
  if (rc != BN_OK) return rc;
  const int chunk = std::min(p->n_bootstrap, 1 << 20);
  WaveHost h;
  if ((rc = setup_wave(s, p, chunk, h)) != BN_OK) return rc;
  MLT_CUDA(cudaMemsetAsync(s->mlt_cnt, 0, 4 * sizeof(unsigned long long), stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  for (int base = 0; base < p->n_bootstrap; base += chunk) {
    h.w.n = std::min(chunk, p->n_bootstrap - base);   // (stride of the per-lane arrays stays `chunk` via dp.n_threads / w.samp layout)
    WaveHost hh = h;
    hh.w.n = h.w.n;
    // the SoA stride of samp / chain arrays is w.n: re-point them for a shorter last chunk
    hh.w.chain_i = hh.w.samp + (size_t)6 * hh.w.n;
    hh.dp.n_threads = hh.w.n;
    hh.grid = std::max(1, std::min(s->num_sms * 8, (hh.w.n + 127) / 128));
    k_mlt_begin<0><<<hh.grid, 128, 0, stream>>>(s->d, hh.dp, hh.st, hh.w, base, hh.n_counter_ints);
    enqueue_bounces(s, hh, stream);
    k_mlt_end<0><<<hh.grid, 128, 0, stream>>>(hh.dp, hh.st, hh.w, base, s->mlt_w, nullptr, hh.D);
  }
  cudaEventRecord(e1, stream);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->mlt_w_host, s->mlt_w, sizeof(float) * (size_t)p->n_bootstrap, cudaMemcpyDeviceToHost, stream);
  unsigned long long r = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&r, s->mlt_cnt, sizeof r, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  float t = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e != cudaSuccess) { s->poisoned = true; ok(e, "pssmlt bootstrap (wavefront)"); return BN_ERR_CUDA; }
  *weights_host = s->mlt_w_host;
  rays = r; ms = t;
  return BN_OK;
}

// phase 2: every mutation round is one wave over the chains of this shard; the round's launch sequence never changes, so it is
// captured once as a CUDA graph and replayed (42 launches per round would otherwise be launch-bound at 65 536 chains)
int run_chains_wavefront(BnScene* s, const BnMltParams* p, MltParams dp_in, float* d_film, cudaStream_t stream, unsigned int* per_chain_host, BnMltStats& out) {
  const int n_run = p->chain_end - p->chain_begin;
  WaveHost h;
  int rc = setup_wave(s, p, n_run, h);
  if (rc != BN_OK) return rc;
  h.dp.mutation_per_chain = dp_in.mutation_per_chain; h.dp.inv_eff = dp_in.inv_eff; h.dp.inv_b = dp_in.inv_b;
  if ((rc = ensure_mlt_misc(s, (size_t)p->n_bootstrap, per_chain_host ? (size_t)n_run : 0)) != BN_OK) return rc;
  unsigned int* d_acc = per_chain_host ? s->mlt_acc : nullptr;
  cudaStream_t cs = stream;
  cudaStream_t own = nullptr;
  if (cs == nullptr) {  // stream capture needs a non-default stream
    MLT_CUDA(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    cs = own;
    cudaDeviceSynchronize();
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaError_t e = cudaMemsetAsync(s->mlt_cnt, 0, 4 * sizeof(unsigned long long), cs);
  cudaEventRecord(e0, cs);
  // the first path of every chain
  k_mlt_begin<1><<<h.grid, 128, 0, cs>>>(s->d, h.dp, h.st, h.w, 0, h.n_counter_ints);
  enqueue_bounces(s, h, cs);
  k_mlt_end<1><<<h.grid, 128, 0, cs>>>(h.dp, h.st, h.w, 0, nullptr, d_film, h.D);
  // one mutation round, captured
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  if (e == cudaSuccess) e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e == cudaSuccess) {
    k_mlt_begin<2><<<h.grid, 128, 0, cs>>>(s->d, h.dp, h.st, h.w, 0, h.n_counter_ints);
    enqueue_bounces(s, h, cs);
    k_mlt_end<2><<<h.grid, 128, 0, cs>>>(h.dp, h.st, h.w, 0, nullptr, d_film, h.D);
    e = cudaStreamEndCapture(cs, &graph);
  }
  if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
  for (int k = 0; e == cudaSuccess && k < h.dp.mutation_per_chain; ++k) e = cudaGraphLaunch(exec, cs);
  if (e == cudaSuccess) {
    k_mlt_finish<<<h.grid, 128, 0, cs>>>(h.dp, h.w, d_film, d_acc);
    e = cudaGetLastError();
  }
  cudaEventRecord(e1, cs);
  unsigned long long cnt[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, s->mlt_cnt, sizeof cnt, cudaMemcpyDeviceToHost, cs);
  if (e == cudaSuccess && per_chain_host) e = cudaMemcpyAsync(per_chain_host, d_acc, sizeof(unsigned int) * (size_t)n_run, cudaMemcpyDeviceToHost, cs);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
  float t = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  if (own) cudaStreamDestroy(own);
  if (e != cudaSuccess) { s->poisoned = true; ok(e, "pssmlt chains (wavefront)"); return BN_ERR_CUDA; }
  out.rays += cnt[0]; out.accepted = cnt[1]; out.proposed = cnt[2]; out.chains_ms = t;
  return BN_OK;
}

// phase 1 on the device; *weights_host points at BootstrapWeights (pinned, owned by the scene)
int run_bootstrap(BnScene* s, const BnMltParams* p, cudaStream_t stream, const float** weights_host, uint64_t& rays, double& ms) {
  if (use_wavefront(p->n_bootstrap)) return run_bootstrap_wavefront(s, p, stream, weights_host, rays, ms);
  MltParams dp = device_params(p);
  const int threads = std::min<long long>((long long)s->num_sms * 16 * 128, ((long long)p->n_bootstrap + 127) / 128 * 128);
  dp.n_threads = threads;
  MltState st{};
  int rc = ensure_mlt_state(s, (size_t)threads, dp.xs_len, st);
  if (rc != BN_OK) return rc;
  if ((rc = ensure_mlt_misc(s, (size_t)p->n_bootstrap, 0)) != BN_OK) return rc;
  unsigned long long* d_rays = s->mlt_cnt + 3;
  MLT_CUDA(cudaMemsetAsync(d_rays, 0, sizeof(unsigned long long), stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  k_mlt_bootstrap<<<threads / 128, 128, 0, stream>>>(s->d, dp, st, s->mlt_w, d_rays);
  cudaEventRecord(e1, stream);
  cudaError_t e = cudaMemcpyAsync(s->mlt_w_host, s->mlt_w, sizeof(float) * (size_t)p->n_bootstrap, cudaMemcpyDeviceToHost, stream);
  unsigned long long r = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&r, d_rays, sizeof r, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  float t = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e != cudaSuccess) { s->poisoned = true; ok(e, "pssmlt bootstrap"); return BN_ERR_CUDA; }
  *weights_host = s->mlt_w_host;
  rays = r; ms = t;
  return BN_OK;
}

int render_pssmlt(BnScene* s, const BnMltParams* p, float* d_film, cudaStream_t stream, BnMltStats* stats, unsigned int* per_chain_host) {
  int rc = validate(s, p);
  if (rc != BN_OK) return rc;
  MLT_CUDA(cudaSetDevice(s->device));
  const float* w = nullptr;
  uint64_t rays_boot = 0;
  double ms_boot = 0;
  rc = run_bootstrap(s, p, stream, &w, rays_boot, ms_boot);
  if (rc != BN_OK) return rc;
  float sum = 0.f;
  for (int k = 0; k < p->n_bootstrap; ++k) sum = sum + w[k];  // Array.average (PSSMLT.fs:394): sequential fp32 sum / n, on the host
  const float B = sum / (float)p->n_bootstrap;
  BnMltStats out{};
  out.b = B; out.rays = rays_boot; out.bootstrap_ms = ms_boot;
  const int n_run = p->chain_end - p->chain_begin;
  if (B != 0.f && n_run > 0) {
    MltParams dp = device_params(p);
    dp.mutation_per_chain = (int)(((uint64_t)p->mutations_per_pixel * (uint64_t)p->width * (uint64_t)p->height + (uint64_t)p->n_chains - 1ull) / (uint64_t)p->n_chains);
    dp.inv_eff = 1.0f / ((float)dp.mutation_per_chain * (float)p->n_chains / (float)(p->width * p->height));
    dp.inv_b = 1.0f / B;
    if (use_wavefront(n_run)) {
      rc = run_chains_wavefront(s, p, dp, d_film, stream, per_chain_host, out);
      if (rc != BN_OK) return rc;
      if (stats) *stats = out;
      return BN_OK;
    }
    const int threads = (n_run + 127) / 128 * 128;
    dp.n_threads = threads;
    MltState st{};
    rc = ensure_mlt_state(s, (size_t)threads, dp.xs_len, st);
    if (rc != BN_OK) return rc;
    if ((rc = ensure_mlt_misc(s, (size_t)p->n_bootstrap, per_chain_host ? (size_t)threads : 0)) != BN_OK) return rc;
    unsigned long long* d_cnt = s->mlt_cnt;
    unsigned int* d_acc = per_chain_host ? s->mlt_acc : nullptr;
    MLT_CUDA(cudaMemsetAsync(d_cnt, 0, 3 * sizeof(unsigned long long), stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
    k_mlt_chains<<<threads / 128, 128, 0, stream>>>(s->d, dp, st, d_film, d_acc, d_cnt);
    cudaEventRecord(e1, stream);
    unsigned long long cnt[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && per_chain_host) e = cudaMemcpyAsync(per_chain_host, d_acc, sizeof(unsigned int) * (size_t)n_run, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    float t = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) { s->poisoned = true; ok(e, "pssmlt chains"); return BN_ERR_CUDA; }
    out.rays += cnt[0]; out.accepted = cnt[1]; out.proposed = cnt[2]; out.chains_ms = t;
  }
  if (stats) *stats = out;
  return BN_OK;
}

}  // namespace

extern "C" {

int bn_pssmlt_bootstrap(BnScene* s, const BnMltParams* p, float* weights) {
  if (!weights) { bnhost::set_error("bn_pssmlt_bootstrap: NULL argument"); return BN_ERR_INVALID; }
  int rc = validate(s, p);
  if (rc != BN_OK) return rc;
  MLT_CUDA(cudaSetDevice(s->device));
  const float* w = nullptr;
  uint64_t rays = 0;
  double ms = 0;
  rc = run_bootstrap(s, p, nullptr, &w, rays, ms);
  if (rc == BN_OK) std::copy(w, w + p->n_bootstrap, weights);
  return rc;
}

int bn_render_pssmlt_device(BnScene* s, const BnMltParams* p, void* d_film, void* stream, BnMltStats* stats) {
  if (!d_film) { bnhost::set_error("bn_render_pssmlt_device: NULL argument"); return BN_ERR_INVALID; }
  return render_pssmlt(s, p, static_cast<float*>(d_film), static_cast<cudaStream_t>(stream), stats, nullptr);
}

// `stats->reserved` != 0 on entry is not used; per-chain accepted counts are exposed through the debug entry below.
int bn_render_pssmlt(BnScene* s, const BnMltParams* p, float* film, BnMltStats* stats) {
  if (!film) { bnhost::set_error("bn_render_pssmlt: NULL argument"); return BN_ERR_INVALID; }
  int rc = validate(s, p);
  if (rc != BN_OK) return rc;
  MLT_CUDA(cudaSetDevice(s->device));
  const size_t len = (size_t)p->width * p->height * 3;
  float* d = nullptr;
  MLT_CUDA(cudaMalloc((void**)&d, len * sizeof(float)));
  if (!ok(cudaMemcpy(d, film, len * sizeof(float), cudaMemcpyHostToDevice), "copy film")) { cudaFree(d); return BN_ERR_CUDA; }
  rc = render_pssmlt(s, p, d, nullptr, stats, nullptr);
  if (rc == BN_OK && !ok(cudaMemcpy(film, d, len * sizeof(float), cudaMemcpyDeviceToHost), "copy film")) rc = BN_ERR_CUDA;
  cudaFree(d);
  return rc;
}

// parity-test entry (declared in the header): per-chain accepted mutation counts next to the film
int bn_debug_render_pssmlt_chains(BnScene* s, const BnMltParams* p, float* film, BnMltStats* stats, unsigned int* per_chain_accepted) {
  if (!film || !per_chain_accepted) return BN_ERR_INVALID;
  int rc = validate(s, p);
  if (rc != BN_OK) return rc;
  MLT_CUDA(cudaSetDevice(s->device));
  const size_t len = (size_t)p->width * p->height * 3;
  float* d = nullptr;
  MLT_CUDA(cudaMalloc((void**)&d, len * sizeof(float)));
  cudaMemcpy(d, film, len * sizeof(float), cudaMemcpyHostToDevice);
  rc = render_pssmlt(s, p, d, nullptr, stats, per_chain_accepted);
  if (rc == BN_OK) cudaMemcpy(film, d, len * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}

}  // extern "C"
