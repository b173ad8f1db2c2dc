// Device-side restatement of the shading half of the hot path:
//   Sampler / Hash                       Base/Sampler.fs, Util/Hash.fs
//   cameras                              Base/Camera.fs:12-20, Pinhole.fs:12-27, ThinLens.fs:12-35
//   materials                            Extensions/Material/{Lambertian,Mirror,Dielectric,PBR}.fs
//   DiffuseLight.Eval                    Base/Light.fs:49-53
//   UniformLightSampler.Sample/Eval      Extensions/LightSampler/Uniform.fs:13-29,40-49
//   MeshInstance/SphereInstance.Sample/EvalPDF  Mesh.fs:289-304, Sphere.fs:90-126
//   AliasTable.Sample                    Util/AliasTable.fs:52-62
// Transcendentals are the bit-reproducible definitions of
// include/bn_portable_math.h (the oracle's "portable" mode evaluates the same).
#pragma once
#include "../../../include/barnacle_b200.h"
#include "../../../include/bn_portable_math.h"
#include "device_scene.h"
#include "vecmath.cuh"

namespace bn {

// ---- Util/Hash.fs ------------------------------------------------------------
BN_DEV uint32_t rotl17(uint32_t h) { return (h << 17) | (h >> 15); }
BN_DEV uint32_t xxhash32_three(uint32_t x, uint32_t y, uint32_t z) {  // Hash.fs:17-28
  const uint32_t p2 = 2246822519u, p3 = 3266489917u, p4 = 668265263u, p5 = 374761393u;
  uint32_t h = z + p5 + x * p3;
  h = p4 * rotl17(h);
  h = h + y * p3;
  h = p4 * rotl17(h);
  h = p2 * (h ^ (h >> 15));
  h = p3 * (h ^ (h >> 13));
  return h ^ (h >> 16);
}
BN_DEV float lcg(uint32_t& seed) {  // Hash.fs:30-32
  seed = 0x00269ec3u + seed * 0x000343fdu;
  return __uint_as_float((seed >> 9) | 0x3f800000u) - 1.f;
}

// ---- cameras -------------------------------------------------------------------
BN_DEV float2 sample_disk_concentric(float ux, float uy) {  // ThinLens.fs:12-23
  float x = ux * 2.f - 1.f, y = uy * 2.f - 1.f;
  if (x == 0.f || y == 0.f) return make_float2(0.f, 0.f);
  float r, theta;
  if (fabsf(x) > fabsf(y)) { r = x; theta = kPi / 4.f * (y / x); }
  else { r = y; theta = kPi / 2.f - kPi / 4.f * (x / y); }
  float s, c;
  bn_sincosf(theta, &s, &c);
  return make_float2(r * c, r * s);
}
// CameraBase.GeneratePrimaryRay (Camera.fs:12-20) over Pinhole/ThinLens.GenerateRay
BN_DEV void primary_ray(const GCamera& cam, int w, int h, int x, int y, float upx, float upy, float ulx, float uly, float3& ro, float3& rd) {
  const float vh = cam.viewport_h;
  const float vw = vh * cam.aspect;
  // pixelLocation = upperLeft + (x+u)*deltaU + (y+v)*deltaV, lane by lane (Pinhole.fs:19-26)
  float3 loc = f3(-(0.5f * vw) + ((float)x + upx) * (vw / (float)w), -(0.5f * vh) + ((float)y + upy) * (vh / (float)h), -1.f);
  float3 o = f3(0.f, 0.f, 0.f);
  float3 d = normalize(loc);
  if (cam.type == 1u && cam.aperture > 0.f) {  // ThinLens.fs:25-35
    float2 dk = sample_disk_concentric(ulx, uly);
    float3 origin = f3(cam.aperture * dk.x, cam.aperture * dk.y, 0.f);
    float3 focus = f3(cam.focus * d.x, cam.focus * d.y, cam.focus * d.z);  // PointAt from the zero origin: fma(t, d, 0)
    d = normalize(focus - origin);
    o = origin;
  }
  Mat43 M;
#pragma unroll
  for (int i = 0; i < 12; ++i) M.m[i] = cam.c2w[i];
  o = transform_point(o, M);
  d = normalize(transform_dir(d, M));
  ro = point_at(o, d, cam.push_forward);
  rd = d;
}

// ---- materials -------------------------------------------------------------------
struct BsdfEval { float3 bsdf; float pdf; };
struct BsdfSample { BsdfEval eval; float3 wi; };

BN_DEV float3 base_color(const GMaterial& m) { return f3(m.r, m.g, m.b); }

BN_DEV BsdfEval lambert_eval(const GMaterial& m, float3 wo, float3 wi) {  // Lambertian.fs:11-16
  BsdfEval e;
  if (wi.z * wo.z < 0.f || net_min(fabsf(wi.z), fabsf(wo.z)) < 1e-6f) { e.bsdf = splat(0.f); e.pdf = 0.f; return e; }
  e.pdf = fabsf(wi.z) / kPi;
  e.bsdf = base_color(m) * e.pdf;
  return e;
}
BN_DEV float3 cosine_hemisphere(float ux, float uy) {  // Lambertian.fs:19-22, PBR.fs:60-63
  float ct = __fsqrt_rn(ux), st = __fsqrt_rn(1.f - ux);
  float sp, cp;
  bn_sincosf(2.f * kPi * uy, &sp, &cp);
  return f3(st * cp, st * sp, ct);
}
BN_DEV float pbr_lambda(float alpha, float3 w) {  // PBR.fs:13-19
  float sin2 = __fmaf_rn(w.x, w.x, w.y * w.y);
  if (sin2 == 0.f) return 0.f;
  float tan2 = sin2 / (w.z * w.z);
  float a2t2 = alpha * alpha * tan2;
  return (-1.f + __fsqrt_rn(1.f + a2t2)) / 2.f;
}
BN_DEV float pbr_d(float alpha, float3 wh) {  // PBR.fs:20-23
  float ch = fabsf(wh.z);
  float x = 1.f + __fmaf_rn(alpha, alpha, -1.f) * ch * ch;
  return alpha * alpha / (kPi * (x * x));
}
BN_DEV BsdfEval pbr_eval(const GMaterial& m, float3 wo, float3 wi) {  // PBR.fs:37-49 (no validity checks)
  const float3 base = base_color(m);
  const float metallic = m.p0, alpha = m.p1;
  float3 wh = normalize(wo + wi);
  float dd = pbr_d(alpha, wh);
  float g = 1.f / (1.f + pbr_lambda(alpha, wo) + pbr_lambda(alpha, wi));
  float spec = dd * g / (4.f * fabsf(wo.z));
  float c = 1.f - dot(wo, wh);
  float c2 = c * c;
  float3 fc = base + (splat(1.f) - base) * (c2 * c2 * c);
  float3 metal = spec * fc;
  const float f0 = 0.04f;
  float f = f0 + (1.f - f0) * c2 * c2 * c;
  float3 diffuse = base * fabsf(wi.z) / kPi;
  float3 inner = vfma(diffuse, splat(1.f - f), splat(spec) * f);
  BsdfEval e;
  e.bsdf = vfma(inner, splat(1.f - metallic), metal * metallic);
  float tmix = 0.5f * (1.f - metallic);
  e.pdf = __fmaf_rn(dd * fabsf(wh.z) / (4.f * dot(wo, wh)), 1.f - tmix, (fabsf(wi.z) / kPi) * tmix);
  return e;
}
BN_DEV BsdfEval material_eval(const GMaterial& m, float3 wo, float3 wi) {
  if (m.type == 0u) return lambert_eval(m, wo, wi);
  if (m.type == 3u) return pbr_eval(m, wo, wi);
  BsdfEval e;  // Mirror.fs:11, Dielectric.fs:13-14
  e.bsdf = splat(0.f); e.pdf = 0.f;
  return e;
}
BN_DEV BsdfSample material_sample(const GMaterial& m, float3 wo, float ulobe, float ux, float uy) {
  BsdfSample s;
  const float3 base = base_color(m);
  if (m.type == 0u) {  // Lambertian.fs:18-26
    float3 wi = cosine_hemisphere(ux, uy);
    float pdf = wi.z / kPi;
    s.eval.bsdf = base * pdf; s.eval.pdf = pdf;
    s.wi = wo.z > 0.f ? wi : -wi;
  } else if (m.type == 1u) {  // Mirror.fs:12-14
    s.eval.bsdf = base; s.eval.pdf = 1.f;
    s.wi = f3(-wo.x, -wo.y, wo.z);
  } else if (m.type == 2u) {  // Dielectric.fs:15-31
    float ior = m.p0;
    float iorp = wo.z > 0.f ? 1.f / ior : ior;
    float cos2 = 1.f - iorp * iorp * __fmaf_rn(-wo.z, wo.z, 1.f);
    if (cos2 <= 0.f) {
      s.eval.bsdf = base; s.eval.pdf = 1.f; s.wi = f3(-wo.x, -wo.y, wo.z);
    } else {
      float a = iorp - 1.f, b = iorp + 1.f;
      float r0 = a * a / (b * b);
      float c = 1.f - fabsf(wo.z);
      float c2 = c * c;
      float r = r0 + (1.f - r0) * c2 * c2 * c;
      if (ulobe < r) {
        s.eval.bsdf = r * base; s.eval.pdf = r; s.wi = f3(-wo.x, -wo.y, wo.z);
      } else {
        float ct = __fsqrt_rn(cos2);
        s.eval.bsdf = (1.f - r) * base; s.eval.pdf = 1.f - r;
        s.wi = f3(-wo.x * iorp, -wo.y * iorp, -copysignf(ct, wo.z));
      }
    }
  } else {  // PBR.fs:50-64
    float metallic = m.p0, alpha = m.p1;
    float3 wi;
    if (ulobe < 1.f - 0.5f * (1.f - metallic)) {
      float th = bn_atanf(alpha * __fsqrt_rn(ux / (1.f - ux)));
      float ph = 2.f * kPi * uy;
      float st, ct, sp, cp;
      bn_sincosf(th, &st, &ct);
      bn_sincosf(ph, &sp, &cp);
      float3 wh = f3(st * cp, st * sp, ct);
      wi = 2.f * dot(wo, wh) * wh - wo;
    } else {
      wi = cosine_hemisphere(ux, uy);
    }
    s.eval = pbr_eval(m, wo, wi);
    s.wi = wi;
  }
  return s;
}

// ---- lights ------------------------------------------------------------------------
BN_DEV float3 light_eval(const GLight& l, float woz) {  // DiffuseLight.Eval, Light.fs:49-53
  if (fabsf(woz) > 1e-6f && (woz > 0.f || l.two_sided)) return f3(l.r, l.g, l.b);
  return splat(0.f);
}
BN_DEV GLight load_light(const DScene& sc, int id) {
  float4 v = __ldg(reinterpret_cast<const float4*>(sc.lights + id));
  GLight l; l.r = v.x; l.g = v.y; l.b = v.z; l.two_sided = __float_as_uint(v.w);
  return l;
}
BN_DEV GMaterial load_material(const DScene& sc, int id) {
  const float4* p = reinterpret_cast<const float4*>(sc.materials + id);
  float4 a = __ldg(p), b = __ldg(p + 1);
  GMaterial m; m.type = __float_as_uint(a.x); m.r = a.y; m.g = a.z; m.b = a.w; m.p0 = b.x; m.p1 = b.y; m.pad0 = 0.f; m.pad1 = 0.f;
  return m;
}
BN_DEV void load_tri(const GTri* t, float3& p0, float3& p1, float3& p2) {
  const float4* p = reinterpret_cast<const float4*>(t);
  float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  p0 = f3(a.x, a.y, a.z); p1 = f3(b.x, b.y, b.z); p2 = f3(c.x, c.y, c.z);
}

struct LightSampleRec { float3 p, L, wi; float pdf; };

// UniformLightSampler.Sample (Uniform.fs:13-29)
BN_DEV LightSampleRec light_sampler_sample(const DScene& sc, float3 p, float usel, float ulx, float uly) {
  const int n = (int)sc.n_light_inst;
  usel = usel * (float)n;
  int id = min((int)usel, n - 1);
  usel = usel - (float)id;
  const uint32_t slot = __ldg(sc.light_inst + id);
  const float4* hp = reinterpret_cast<const float4*>(sc.inst_head + slot);
  const float4 h0 = __ldg(hp), h2 = __ldg(hp + 2);
  const uint32_t kind_prim = __float_as_uint(h0.w);
  const Mat43 M = load_mat43(reinterpret_cast<const float4*>(sc.inst_o2w + slot));
  float3 ip, inorm;
  float pdf_surface;
  if (kind_prim & 0x80000000u) {  // SphereInstance.Sample, Sphere.fs:90-113
    float radius = __ldg(sc.sphere_radii + (kind_prim & 0x7FFFFFFFu));
    float st, ct;
    bn_sincosf(2.f * kPi * ulx, &st, &ct);
    float cphi = __fmaf_rn(-2.f, uly, 1.f);
    float sphi = __fsqrt_rn(__fmaf_rn(-cphi, cphi, 1.f));
    float3 nl = f3(ct * sphi, st * sphi, cphi);
    Onb f = onb_from_n(nl);
    ip = transform_point(nl * radius, M);
    float3 np = cross(transform_dir(f.t, M), transform_dir(f.b, M));
    float inv_j = 1.f / length(np);
    inorm = inv_j * np;
    pdf_surface = inv_j / (4.f * kPi * radius * radius);
  } else {  // MeshInstance.Sample, Mesh.fs:289-298
    const GMesh* mesh = sc.meshes + kind_prim;
    const float4 m2 = __ldg(reinterpret_cast<const float4*>(mesh) + 2);
    const uint32_t tri_base = __float_as_uint(m2.x), tri_count = __float_as_uint(m2.y), alias_base = __float_as_uint(m2.z);
    // AliasTable.Sample, AliasTable.fs:52-62
    float u = usel * (float)tri_count;
    int idx = (int)u;
    const GAlias* table = sc.alias + alias_base;
    GAlias e = table[idx];
    u = u - (float)idx;
    float pdf_tri;
    if (u < e.prob) pdf_tri = e.pdf;
    else { idx = e.alias; pdf_tri = table[idx].pdf; }
    float3 q0, q1, q2;
    load_tri(sc.tris + tri_base + idx, q0, q1, q2);
    q0 = transform_point(q0, M); q1 = transform_point(q1, M); q2 = transform_point(q2, M);  // Triangle.Transform
    // Triangle.Sample, Mesh.fs:89-99
    float uvx, uvy;
    if (ulx < uly) { uvx = 0.5f * ulx; uvy = __fmaf_rn(-0.5f, ulx, uly); }
    else { uvx = __fmaf_rn(-0.5f, uly, ulx); uvy = 0.5f * uly; }
    ip = (uvx * q1 + uvy * q2) + (1.f - uvx - uvy) * q0;
    float3 nn = cross(q1 - q0, q2 - q0);
    float pdf = 2.f / length(nn);
    inorm = (0.5f * pdf) * nn;
    pdf_surface = pdf_tri * pdf;
  }
  float3 wo = normalize(p - ip);
  float cos_wo = dot(inorm, wo);
  float dist2 = length_sq(p - ip);
  LightSampleRec r;
  r.p = ip;
  r.L = light_eval(load_light(sc, __float_as_int(h2.x)), dot(wo, inorm));
  r.pdf = dist2 * pdf_surface / (net_max(fabsf(cos_wo), 1e-6f) * (float)n);
  r.wi = -wo;
  return r;
}

// ---- one lane of the shade kernel: one iteration of Li's loop body -------------------------------------------
// (PathTracing.fs:30-79; Direct.fs:10-40 and Normal.fs:10-17 as modes).  Inputs: the path's state planes a, b, c and
// its hit record h; `rad` is the per-path radiance array (read-modify-written for emitter hits).  Everything the
// kernel then writes to the queues comes back through the reference parameters (initialised by the caller).  A function of its own so that tests/hostsim can run the very
// same code on the host against the oracle; k_shade (kernels.cu) only does the queue I/O around it.
BN_DEV void shade_lane(const DScene& sc, const int integrator, const int rr_depth, const int max_depth, const uint32_t flags, const int bounce,
                       const float4 a, const float4 b, const float4 c, const float4 h, float4* __restrict__ rad,
                       bool& alive, bool& has_shadow, bool& ref_shadow, float3& P, float3& nd, float3& beta, float& bs_pdf, uint32_t& rng, int& pid,
                       float3& sh_wi, float3& sh_a, float3& sh_b, float& sh_tmax) {
  const float3 o = f3(a.x, a.y, a.z), d = f3(a.w, b.x, b.y);
  beta = f3(b.z, b.w, c.x);
  const float prev_pdf = c.y;
  rng = __float_as_uint(c.z);
  pid = __float_as_int(c.w);
  const float t = h.x;
  const int inst = __float_as_int(h.y), prim = __float_as_int(h.z);
  if (inst >= 0) {
    const float4* hp = reinterpret_cast<const float4*>(sc.inst_head + inst);
    const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
    const uint32_t kind_prim = __float_as_uint(h0.w);
    const int material = __float_as_int(h1.w), light = __float_as_int(h2.x);
    const Mat43 W2O = load_mat43(reinterpret_cast<const float4*>(sc.inst_w2o + inst));
    const Mat43 O2W = load_mat43(reinterpret_cast<const float4*>(sc.inst_o2w + inst));
    // rebuild the interaction exactly as the intersection routines produced it
    const float3 oo = transform_point(o, W2O), od = transform_dir(d, W2O);
    const float3 pobj = point_at(oo, od, t);
    float3 nobj;
    const bool is_sphere = (kind_prim & 0x80000000u) != 0u;
    if (is_sphere) {  // Sphere.fs:50-62 / 64-75 (normal flipped on the near root only)
      nobj = normalize(pobj);
      if (prim == 0 && dot(nobj, od) > 0.f) nobj = -nobj;
    } else {          // Mesh.fs:76-78
      float3 p0, p1, p2;
      load_tri(sc.tris + prim, p0, p1, p2);  // prim is the scene-wide triangle index
      nobj = normalize(cross(p1 - p0, p2 - p0));
    }
    // LocalGeometry.Transform (Primitive.fs:57-58)
    P = transform_point(pobj, O2W);
    const Onb onb = transform_onb(onb_from_n(nobj), O2W);

    if (integrator == BN_INTEGRATOR_NORMAL) {  // NormalIntegrator.Li (Normal.fs:10-17)
      const float3 c = 0.5f * (onb.n + splat(1.f));
      rad[pid] = make_float4(c.x, c.y, c.z, 0.f);
    } else if (integrator == BN_INTEGRATOR_DIRECT && bounce == 0) {
      if (light >= 0) {  // Direct.fs:16-17: L + EvalEmit(-ray.Direction)
        const float3 Le = light_eval(load_light(sc, light), dot(-d, onb.n));
        const float4 L4 = rad[pid];
        rad[pid] = make_float4(L4.x + Le.x, L4.y + Le.y, L4.z + Le.z, 0.f);
      }
    } else if (light >= 0) {  // PathTracing.fs:30-40 + UniformLightSampler.Eval (Uniform.fs:40-49)
      const float3 wo = normalize(o - P);
      const float cos_wo = dot(onb.n, wo);
      float pdf_surface;
      if (is_sphere) {  // SphereInstance.EvalPDF, Sphere.fs:115-126
        const float radius = h2.y;
        const float j = length(cross(transform_dir(onb.t, W2O), transform_dir(onb.b, W2O)));
        pdf_surface = j / (4.f * kPi * radius * radius);
      } else {          // MeshInstance.EvalPDF with tag = 0 (Mesh.fs:300-304), host-precomputed
        pdf_surface = h2.y;
      }
      const float dist2 = length_sq(o - P);
      const float3 Le = light_eval(load_light(sc, light), dot(wo, onb.n));
      const float lpdf = dist2 * pdf_surface / (net_max(fabsf(cos_wo), 1e-6f) * (float)sc.n_light_inst);
      // PathTracing.fs:33-38 MIS weight | Direct.fs:36-38: bsdf * L * (1 / lightPdf), no MIS
      const float w = integrator == BN_INTEGRATOR_DIRECT ? (1.f / lpdf) : (bounce == 0 ? 1.f : prev_pdf * (1.f / (lpdf + prev_pdf)));
      const float4 L4 = rad[pid];
      const float3 L = vfma(beta, Le * w, f3(L4.x, L4.y, L4.z));
      rad[pid] = make_float4(L.x, L.y, L.z, 0.f);
    }
    const bool scatter = integrator == BN_INTEGRATOR_PATH_TRACING || (integrator == BN_INTEGRATOR_DIRECT && bounce == 0);
    if (material >= 0 && scatter) {
      const bool direct = integrator == BN_INTEGRATOR_DIRECT;
      const GMaterial mat = load_material(sc, material);
      const float usel = lcg(rng);
      const float ulx = lcg(rng), uly = lcg(rng);
      const LightSampleRec ls = light_sampler_sample(sc, P, usel, ulx, uly);  // PathTracing.fs:43
      const float dist = length(ls.p - P);
      const float3 wo_l = world_to_local(onb, -d);
      if (direct || ls.pdf != 0.f) {  // PathTracing.fs:47-59 | Direct.fs:25-29 traces whatever the pdf
        ref_shadow = true;
        const BsdfEval fe = material_eval(mat, wo_l, world_to_local(onb, ls.wi));
        sh_a = direct ? fe.bsdf : beta * fe.bsdf;
        sh_b = ls.L * (1.f / (direct ? ls.pdf : fe.pdf + ls.pdf));
        // fma(0, finite, L) == L bit for bit: the connection cannot change the image
        const bool null_contrib = sh_a.x == 0.f && sh_a.y == 0.f && sh_a.z == 0.f && isfinite(sh_b.x) && isfinite(sh_b.y) && isfinite(sh_b.z);
        has_shadow = !null_contrib || (flags & BN_RENDER_TRACE_NULL_SHADOW);
        sh_wi = ls.wi;
        sh_tmax = dist - 1e-3f;
      }
      const float ulobe = lcg(rng);
      const float ubx = lcg(rng), uby = lcg(rng);
      const BsdfSample bs = material_sample(mat, wo_l, ulobe, ubx, uby);  // :61
      if (direct) {  // Direct.fs:31-38: the sampled direction is followed whatever its pdf; its weight is bsdf alone
        nd = local_to_world(onb, bs.wi);
        beta = bs.eval.bsdf;
        bs_pdf = bs.eval.pdf;
        alive = true;
      } else if (bs.eval.pdf != 0.f) {
        nd = local_to_world(onb, bs.wi);
        beta = beta * bs.eval.bsdf * (1.f / bs.eval.pdf);
        bs_pdf = bs.eval.pdf;
        bool cont = true;
        if (bounce >= rr_depth) {  // :69-75
          const float q = net_min(1.f, net_max(beta.x, net_max(beta.y, beta.z)));
          if (lcg(rng) < q) beta = beta * (1.f / q);
          else cont = false;
        }
        alive = cont && (bounce + 1 < max_depth);
      }
    }
  }
}

}  // namespace bn
