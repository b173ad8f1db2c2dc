// TEST INFRASTRUCTURE — the kernels' own __device__ functions (traverse.cuh: trace_lane_impl in its fast and exact
// forms; shade.cuh: camera, materials, light sampler) compiled for the HOST through device_shim.h, over the device
// layout that bn_scene_create would upload (scene_convert.cpp).  tests/test_hostsim.py compares the results bit for
// bit with the oracle: a CPU-side regression net for the GPU path's arithmetic and data layout.  What it cannot
// cover is the warp-synchronous scheduling of traverse_persistent and the wavefront in kernels.cu — the -m gpu
// parity tests do that on the B200.  Built by the test with g++ -ffp-contract=off; never linked into the product.
#include "device_shim.h"

#include <string>
#include <vector>

#include "../../barnacle_b200/csrc/cuda/scene_convert.h"
#include "../../barnacle_b200/csrc/cuda/traverse.cuh"
#include "../../barnacle_b200/csrc/cuda/shade.cuh"

namespace {
struct HsScene {
  bnconv::ConvertedScene cs;
  bn::DScene d;
};
std::string g_error;
}  // namespace

extern "C" {

const char* hs_last_error(void) { return g_error.c_str(); }

void* hs_scene_create(const BnSceneDesc* desc) {
  auto* s = new HsScene();
  if (!bnconv::convert_scene(*desc, s->cs, g_error)) { delete s; return nullptr; }
  bn::DScene& d = s->d;
  const bnconv::ConvertedScene& cs = s->cs;
  d.nodes = cs.nodes.data(); d.inst_trav = cs.inst_trav.data(); d.inst_head = cs.inst_head.data(); d.inst_w2o = cs.inst_w2o.data();
  d.inst_o2w = cs.inst_o2w.data(); d.meshes = cs.meshes.data(); d.tris = cs.tris.data(); d.alias = cs.alias.data();
  d.sphere_radii = cs.sphere_radii.data(); d.materials = cs.materials.data(); d.lights = cs.lights.data(); d.light_inst = cs.light_inst.data();
  d.flat_tlas = nullptr;  // the ordered scan belongs to traverse_persistent; the per-lane walk uses the tree
  d.wide = nullptr;       // ... and so do the 4-wide nodes
  d.tlas_wroot = 0;
  d.tlas = cs.tlas;
  d.n_inst = (uint32_t)cs.inst_head.size();
  d.n_light_inst = (uint32_t)cs.light_inst.size();
  d.all_finite = cs.all_finite ? 1u : 0u;
  d.cam = cs.cam;
  return s;
}
void hs_scene_destroy(void* h) { delete static_cast<HsScene*>(h); }
int hs_scene_max_stack(void* h) { return static_cast<HsScene*>(h)->cs.max_stack; }

// mode 0: exact form only (what the fix-up kernel runs); mode 1: fast form with the exact form as fallback when the
// ray does not qualify (what the phases + the deferral amount to).  fast_used (may be NULL) counts rays the fast form finished.
int hs_trace(void* h, const BnRay* rays, uint64_t n, int any_hit, int mode, BnHit* hits, uint64_t* fast_used) {
  const bn::DScene& sc = static_cast<HsScene*>(h)->d;
  uint64_t fast = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const float3 o = bn::f3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
    const float3 d = bn::f3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
    bn::TraceResult r;
    bool done = false;
    if (mode == 1 && sc.all_finite != 0u && bn::slab_fast_ok(o, bn::rcp3(d))) {
      done = any_hit ? bn::trace_lane_impl<true, true>(sc, o, d, rays[i].tmax, r) : bn::trace_lane_impl<false, true>(sc, o, d, rays[i].tmax, r);
      fast += done ? 1 : 0;
    }
    if (!done) {
      if (any_hit) bn::trace_lane_impl<true, false>(sc, o, d, rays[i].tmax, r);
      else bn::trace_lane_impl<false, false>(sc, o, d, rays[i].tmax, r);
    }
    BnHit out;
    if (any_hit) {
      out.t = 0.f; out.u = 0.f; out.v = 0.f; out.instance = r.hit ? 1 : 0; out.primitive = 0;
    } else {  // TraceIO::store (kernels.cu), minus the sphere uv (libm atan2 / acos: not part of the parity contract)
      out.t = r.t; out.u = r.u; out.v = r.v; out.instance = r.inst; out.primitive = r.prim;
      if (r.hit) {
        const bool sphere = sc.inst_trav[r.inst].is_sphere != 0u;
        out.primitive = sphere ? 0 : r.prim - (int)sc.inst_trav[r.inst].tri_base;
        if (sphere) { out.u = 0.f; out.v = 0.f; }
      }
    }
    hits[i] = out;
  }
  if (fast_used) *fast_used = fast;
  return 0;
}

// CameraBase.GeneratePrimaryRay from explicit samples: u = uPixel.xy, uLens.xy
void hs_camera_ray(void* h, int width, int height, int x, int y, const float* u, BnRay* out) {
  const bn::DScene& sc = static_cast<HsScene*>(h)->d;
  float3 o, d;
  bn::primary_ray(sc.cam, width, height, x, y, u[0], u[1], u[2], u[3], o, d);
  out->origin[0] = o.x; out->origin[1] = o.y; out->origin[2] = o.z;
  out->direction[0] = d.x; out->direction[1] = d.y; out->direction[2] = d.z;
  out->tmax = INFINITY;
}

static bn::GMaterial to_gmat(const BnMaterial* m) { return bn::GMaterial{m->type, m->base_color[0], m->base_color[1], m->base_color[2], m->p0, m->p1, 0.f, 0.f}; }

void hs_material_eval(const BnMaterial* m, const float* wo, const float* wi, float* out) {
  const bn::BsdfEval e = bn::material_eval(to_gmat(m), bn::f3(wo[0], wo[1], wo[2]), bn::f3(wi[0], wi[1], wi[2]));
  out[0] = e.bsdf.x; out[1] = e.bsdf.y; out[2] = e.bsdf.z; out[3] = e.pdf;
}
void hs_material_sample(const BnMaterial* m, const float* wo, float ulobe, const float* u, float* out) {
  const bn::BsdfSample b = bn::material_sample(to_gmat(m), bn::f3(wo[0], wo[1], wo[2]), ulobe, u[0], u[1]);
  out[0] = b.eval.bsdf.x; out[1] = b.eval.bsdf.y; out[2] = b.eval.bsdf.z; out[3] = b.eval.pdf;
  out[4] = b.wi.x; out[5] = b.wi.y; out[6] = b.wi.z;
}
void hs_light_sample(void* h, const float* p, float usel, const float* ul, float* out) {
  const bn::DScene& sc = static_cast<HsScene*>(h)->d;
  const bn::LightSampleRec r = bn::light_sampler_sample(sc, bn::f3(p[0], p[1], p[2]), usel, ul[0], ul[1]);
  out[0] = r.p.x; out[1] = r.p.y; out[2] = r.p.z; out[3] = r.L.x; out[4] = r.L.y; out[5] = r.L.z; out[6] = r.pdf;
  out[7] = r.wi.x; out[8] = r.wi.y; out[9] = r.wi.z;
}
// One camera path through the kernels' device functions, in the order the wavefront runs them for that path:
// raygen (k_raygen's body, restated: four draws, primary_ray) -> per bounce: extend (trace_lane_impl, fast form with the
// exact form as fallback = phases + fix-up kernel) -> shade_lane (the shade kernel's lane body, shade.cuh) -> shadow ray
// (any hit) + connect (ShadowIO::store: one fma).  Returns the path's radiance; rays[0] / rays[1] count extend / shadow rays.
static float3 hs_path(const bn::DScene& sc, const BnRenderParams& p, int x, int y, int sample, uint64_t* rays) {
  uint32_t rng = bn::xxhash32_three((uint32_t)x, (uint32_t)y, (uint32_t)(p.frame_id * p.spp + sample));
  const float upx = bn::lcg(rng), upy = bn::lcg(rng);
  const float ulx = bn::lcg(rng), uly = bn::lcg(rng);
  float3 o, d;
  bn::primary_ray(sc.cam, p.width, p.height, x, y, upx, upy, ulx, uly, o, d);
  float4 rad = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a = make_float4(o.x, o.y, o.z, d.x), b = make_float4(d.y, d.z, 1.f, 1.f), c = make_float4(1.f, 0.f, __uint_as_float(rng), __int_as_float(0));
  const int n_bounces = p.integrator == BN_INTEGRATOR_PATH_TRACING ? p.max_depth : (p.integrator == BN_INTEGRATOR_DIRECT ? 2 : 1);
  auto trace = [&](bool any, float3 ro, float3 rd, float tmax) {
    bn::TraceResult r;
    bool done = false;
    if (sc.all_finite != 0u && bn::slab_fast_ok(ro, bn::rcp3(rd)))
      done = any ? bn::trace_lane_impl<true, true>(sc, ro, rd, tmax, r) : bn::trace_lane_impl<false, true>(sc, ro, rd, tmax, r);
    if (!done) {
      if (any) bn::trace_lane_impl<true, false>(sc, ro, rd, tmax, r);
      else bn::trace_lane_impl<false, false>(sc, ro, rd, tmax, r);
    }
    return r;
  };
  for (int bounce = 0; bounce < n_bounces; ++bounce) {
    const bn::TraceResult hit = trace(false, bn::f3(a.x, a.y, a.z), bn::f3(a.w, b.x, b.y), INFINITY);
    rays[0]++;
    const float4 h = make_float4(hit.t, __int_as_float(hit.inst), __int_as_float(hit.prim), 0.f);  // ExtendIO::store
    bool alive = false, has_shadow = false, ref_shadow = false;
    float3 P = bn::splat(0.f), nd = bn::splat(0.f), beta = bn::splat(0.f), sh_wi = bn::splat(0.f), sh_a = bn::splat(0.f), sh_b = bn::splat(0.f);
    float bs_pdf = 0.f, sh_tmax = 0.f;
    uint32_t rng2 = 0;
    int pid = 0;
    bn::shade_lane(sc, p.integrator, p.rr_depth, p.max_depth, p.flags, bounce, a, b, c, h, &rad, alive, has_shadow, ref_shadow, P, nd, beta, bs_pdf, rng2, pid,
                   sh_wi, sh_a, sh_b, sh_tmax);
    if (has_shadow) {
      rays[1]++;
      if (!trace(true, P, sh_wi, sh_tmax).hit) {  // ShadowIO::store
        const float3 L = bn::vfma(sh_a, sh_b, bn::f3(rad.x, rad.y, rad.z));
        rad = make_float4(L.x, L.y, L.z, 0.f);
      }
    }
    if (!alive) break;
    a = make_float4(P.x, P.y, P.z, nd.x); b = make_float4(nd.y, nd.z, beta.x, beta.y); c = make_float4(beta.z, bs_pdf, __uint_as_float(rng2), __int_as_float(0));
  }
  return bn::f3(rad.x, rad.y, rad.z);
}

// radiance: [sample][y][x][3] over the full image (bn_render_radiance's layout); film (may be NULL): Film.Pixels layout,
// accumulated like k_accumulate (fma(1/spp, L, acc) over sampleId ascending).  stats (may be NULL): extend, shadow rays.
int hs_render(void* h, const BnRenderParams* p, float* radiance, float* film, uint64_t* stats) {
  const bn::DScene& sc = static_cast<HsScene*>(h)->d;
  uint64_t rays[2] = {0, 0};
  const float inv_spp = 1.f / (float)p->spp;
  for (int y = 0; y < p->height; ++y)
    for (int x = 0; x < p->width; ++x) {
      float3 acc = bn::splat(0.f);
      for (int s = 0; s < p->spp; ++s) {
        const float3 L = hs_path(sc, *p, x, y, s, rays);
        if (radiance) {
          float* o = radiance + (((size_t)s * p->height + y) * p->width + x) * 3;
          o[0] = L.x; o[1] = L.y; o[2] = L.z;
        }
        acc = bn::vfma(bn::splat(inv_spp), bn::operator*(L, 1.f / 1.f), acc);  // * rcp(camera pdf), pdf == 1 (k_accumulate)
      }
      if (film) {
        float* px = film + ((size_t)(p->height - y - 1) * p->width + x) * 3;
        px[0] = acc.x; px[1] = acc.y; px[2] = acc.z;
      }
    }
  if (stats) { stats[0] = rays[0]; stats[1] = rays[1]; }
  return 0;
}

uint32_t hs_xxhash32_three(uint32_t x, uint32_t y, uint32_t z) { return bn::xxhash32_three(x, y, z); }
float hs_lcg(uint32_t* state) { return bn::lcg(*state); }

}  // extern "C"
