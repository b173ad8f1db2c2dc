"""Function-level pin of the oracle's light sampling: UniformLightSampler.Sample with
MeshInstance.Sample / Triangle.Sample / AliasTable.Sample and SphereInstance.Sample, against a second restatement of
Uniform.fs:13-29, Mesh.fs:84-99,289-298, AliasTable.fs:14-62 and Sphere.fs:90-113 written independently in float64
numpy — plus what the formulas must satisfy geometrically (the point lies on the emitter, wi points at it, the area
pdf of an untransformed shape is 1 / area).  CPU only, libm mode."""
import json
import math

import numpy as np
import pytest

from barnacle_b200.scene import Scene
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene


@pytest.fixture(autouse=True)
def _libm_mode(oracle_lib):
    oracle_ffi.set_portable_math(False)
    yield


def _scene(primitives, instances, nodes, transforms, lights):
    return json.dumps({"nodes": nodes, "instances": instances, "transforms": transforms, "primitives": primitives, "materials": [],
                       "lights": lights, "integrator": {"type": "path-tracing", "spp": 1}, "camera": {"type": "pinhole", "fov": 90.0},
                       "film": {"width": 8, "height": 8, "tone-mapping": "identity"}})


def _mat(inst, which):
    return np.array(getattr(inst, which)[:], dtype=np.float64).reshape(4, 4)


def _diffuse_light_eval(light, woz):  # Light.fs:49-53
    if abs(woz) > 1e-6 and (woz > 0 or light.two_sided):
        return np.array(light.emission[:], dtype=np.float64)
    return np.zeros(3)


def ref_light_sample(desc, p, usel, ul):
    """Uniform.fs:13-29 over the flattened scene (float64; every value read from BnSceneDesc)."""
    n = desc.light_instance_count
    usel = np.float32(np.float32(usel) * np.float32(n))
    lid = min(int(usel), n - 1)
    usel = float(np.float32(usel - np.float32(lid)))
    inst = desc.instances[desc.light_instances[lid]]
    o2w = _mat(inst, "object_to_world")
    if inst.prim_kind == 0:  # MeshInstance.Sample, Mesh.fs:289-298
        m = desc.meshes[inst.prim_id]
        u = np.float32(np.float32(usel) * np.float32(m.tri_count))  # AliasTable.Sample, AliasTable.fs:52-62
        idx = int(u)
        e = desc.alias[m.alias_offset + idx]
        if float(u - np.float32(idx)) >= e.prob:
            idx = e.alias
        pdf_tri = desc.alias[m.alias_offset + idx].pdf
        def vertex(k):
            v = 3 * (m.vertex_offset + desc.triangles[3 * (m.tri_offset + idx) + k])
            return np.array([desc.vertices[v], desc.vertices[v + 1], desc.vertices[v + 2]], dtype=np.float64) @ o2w[:3, :3] + o2w[3, :3]
        tri = [vertex(0), vertex(1), vertex(2)]
        ux, uy = float(np.float32(ul[0])), float(np.float32(ul[1]))
        a, b = (0.5 * ux, uy - 0.5 * ux) if ux < uy else (ux - 0.5 * uy, 0.5 * uy)  # Triangle.Sample, Mesh.fs:89-99
        pos = a * tri[1] + b * tri[2] + (1 - a - b) * tri[0]
        nn = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        pdf_area = 2 / np.linalg.norm(nn)
        normal = 0.5 * pdf_area * nn
        pdf_surface = pdf_tri * pdf_area
    else:  # SphereInstance.Sample, Sphere.fs:90-113
        radius = float(desc.sphere_radii[inst.prim_id])
        th = 2 * math.pi * float(np.float32(ul[0]))
        cphi = 1 - 2 * float(np.float32(ul[1]))
        sphi = math.sqrt(1 - cphi * cphi)
        nl = np.array([math.cos(th) * sphi, math.sin(th) * sphi, cphi])
        axis = np.array([0.0, 1.0, 0.0]) if abs(nl[0]) > 0.1 else np.array([1.0, 0.0, 0.0])  # OrthonormalBasis, Primitive.fs:15-23
        t = np.cross(nl, axis)
        t /= np.linalg.norm(t)
        b = np.cross(nl, t)
        pos = (nl * radius) @ o2w[:3, :3] + o2w[3, :3]
        npr = np.cross(t @ o2w[:3, :3], b @ o2w[:3, :3])
        inv_j = 1 / np.linalg.norm(npr)
        normal = inv_j * npr
        pdf_surface = inv_j / (4 * math.pi * radius * radius)
    wo = (p - pos) / np.linalg.norm(p - pos)
    cos_wo = float(normal @ wo)
    L = _diffuse_light_eval(desc.lights[inst.light_id], cos_wo)
    pdf = float((p - pos) @ (p - pos)) * pdf_surface / (max(abs(cos_wo), 1e-6) * n)
    return pos, L, pdf, -wo, normal


def _check(scene, n_samples, seed, rtol=2e-4):
    desc = scene.desc.contents
    oracle = OracleScene(scene.desc)
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_samples):
        p = rng.uniform(-3, 3, size=3).astype(np.float32)
        usel, ul = float(rng.random(dtype=np.float32)), rng.random(2, dtype=np.float32)
        got = oracle.light_sample(p, usel, ul)
        pos, L, pdf, wi, normal = ref_light_sample(desc, p.astype(np.float64), usel, ul)
        np.testing.assert_allclose(got[0:3], pos, rtol=rtol, atol=2e-5)
        np.testing.assert_allclose(got[7:10], wi, rtol=rtol, atol=2e-5)
        if abs(float(normal @ wi)) > 1e-3:  # away from the emitter's horizon, where |cos| is all rounding
            np.testing.assert_allclose(got[3:6], L, rtol=1e-6)
            assert got[6] == pytest.approx(pdf, rel=2e-3)
        out.append((p, got, pos, normal))
    return out


def test_cbox_quad_light_matches_restatement(scene_loader):
    scene = scene_loader("cbox_pt")
    samples = _check(scene, 300, seed=11)
    desc = scene.desc.contents
    inst = desc.instances[desc.light_instances[0]]
    lo, hi = np.array(inst.bounds_min[:]), np.array(inst.bounds_max[:])
    pts = np.array([g[0:3] for _, g, _, _ in samples])
    assert ((pts >= lo - 1e-3) & (pts <= hi + 1e-3)).all()              # every point lies on the emitter
    assert np.ptp(pts, axis=0).max() > 0.5 * (hi - lo).max()             # and they spread over it


def test_two_lights_alias_quirk_and_transformed_sphere(lib):
    """A unit quad (two equal triangles) and a non-uniformly scaled sphere, both emitters.  Pins the instance pick
    (min(int(u N), N - 1) and the remapped u, Uniform.fs:14-16), the alias-table pick (every prob is 1, so the
    triangle is int(u n) and its pdf is area / total — 1/2 each here) and the sphere's Jacobian (Sphere.fs:100-113)."""
    tr = [{"keyframes": [{"scale": [1.0, 2.0, 0.5], "translation": [5.0, 1.0, -2.0]}]},
          {"keyframes": [{"translation": [-4.0, 6.0, 0.0]}]}]
    text = _scene([{"type": "quad"}, {"type": "sphere", "radius": 1.5}],
                  [{"primitive": 0, "light": 0}, {"primitive": 1, "light": 1}],
                  [{"children": [1, 2, 3]}, {"instances": [0], "transform": 1}, {"instances": [1], "transform": 0}, {"has-camera": True}],
                  tr, [{"type": "diffuse", "emission": [3.0, 2.0, 1.0]}, {"type": "diffuse", "emission": [1.0, 4.0, 9.0], "two-sided": False}])
    scene = Scene.LoadString(text)
    desc = scene.desc.contents
    assert desc.light_instance_count == 2
    samples = _check(scene, 400, seed=12)
    oracle = OracleScene(scene.desc)
    kinds = [desc.instances[desc.light_instances[k]].prim_kind for k in range(2)]
    quad_slot = kinds.index(0)
    # untransformed-shape area pdf: the quad is 2 x 2 (area 4); picked with probability 1/2, its two triangles with the
    # alias table's pdf 1/2 each, uniform within (2 / |n| = 1 / 2): pdf_A = 1/2 * 1/2 = 1/4 = 1 / area (Q1 does not bias equal triangles)
    p = np.array([-4.0, 9.0, 0.0], dtype=np.float32)                      # 3 above the quad's centre
    got = oracle.light_sample(p, (quad_slot + 0.25) / 2, [0.3, 0.6])
    d2 = float(((p - got[0:3]) ** 2).sum())
    cos = abs(float(got[8]))                                              # quad normal is +-y
    assert got[6] == pytest.approx(d2 * 0.25 / (cos * 2), rel=1e-5)
    assert got[1] == pytest.approx(6.0, abs=1e-6) and abs(got[0] + 4.0) <= 1.0 + 1e-6 and abs(got[2]) <= 1.0 + 1e-6
    # the one-sided sphere emits only outwards: a sampled point is lit iff its (transformed) normal faces p
    sph = [(pt, g, pos, nrm) for pt, g, pos, nrm in samples if abs(g[1] - 6.0) > 1e-3 or abs(g[0] + 4.0) > 1.001]
    assert len(sph) > 100
    facing = [float(nrm @ (pt - pos)) / np.linalg.norm(pt - pos) for pt, g, pos, nrm in sph]
    lit = [bool(g[3:6].any()) for pt, g, pos, nrm in sph]
    assert all(l == (f > 0) for l, f in zip(lit, facing) if abs(f) > 1e-3) and any(lit) and not all(lit)
    inside = oracle.light_sample(np.array([5.0, 1.0, -2.0], dtype=np.float32), (1 - quad_slot + 0.5) / 2, [0.2, 0.7])
    assert not inside[3:6].any()                                          # from its centre every point faces away


def ref_light_eval(desc, inst, origin, p, n, t, b):
    """UniformLightSampler.Eval (Uniform.fs:40-49) with MeshInstance.EvalPDF (Mesh.fs:300-304; the tag is always 0 after
    LocalGeometry.Transform) and SphereInstance.EvalPDF (Sphere.fs:115-126), float64."""
    n_lights = desc.light_instance_count
    if inst.prim_kind == 0:
        m = desc.meshes[inst.prim_id]
        o2w = _mat(inst, "object_to_world")

        def vertex(k):
            v = 3 * (m.vertex_offset + desc.triangles[3 * m.tri_offset + k])
            return np.array([desc.vertices[v], desc.vertices[v + 1], desc.vertices[v + 2]], dtype=np.float64) @ o2w[:3, :3] + o2w[3, :3]
        p0, p1, p2 = vertex(0), vertex(1), vertex(2)
        pdf_surface = desc.alias[m.alias_offset].pdf / (0.5 * np.linalg.norm(np.cross(p1 - p0, p2 - p0)))
    else:
        w2o = _mat(inst, "world_to_object")
        radius = float(desc.sphere_radii[inst.prim_id])
        pdf_surface = np.linalg.norm(np.cross(t @ w2o[:3, :3], b @ w2o[:3, :3])) / (4 * math.pi * radius * radius)
    wo = (origin - p) / np.linalg.norm(origin - p)
    cos_wo = float(n @ wo)
    return _diffuse_light_eval(desc.lights[inst.light_id], cos_wo), float((origin - p) @ (origin - p)) * pdf_surface / (max(abs(cos_wo), 1e-6) * n_lights), cos_wo


def test_light_eval_at_emitter_hits_matches_restatement(lib):
    """The MIS side: a BSDF-sampled ray that lands on an emitter is weighted with UniformLightSampler.Eval's pdf."""
    from barnacle_b200.scene import RAY_DTYPE
    tr = [{"keyframes": [{"scale": [1.0, 2.0, 0.5], "rotation": [0.4, 0.0, -0.3], "translation": [5.0, 1.0, -2.0]}]},
          {"keyframes": [{"scale": [3.0, 1.0, 2.0], "translation": [-4.0, 6.0, 0.0]}]}]
    text = _scene([{"type": "quad"}, {"type": "sphere", "radius": 1.5}],
                  [{"primitive": 0, "light": 0}, {"primitive": 1, "light": 1}],
                  [{"children": [1, 2, 3]}, {"instances": [0], "transform": 1}, {"instances": [1], "transform": 0}, {"has-camera": True}],
                  tr, [{"type": "diffuse", "emission": [3.0, 2.0, 1.0]}, {"type": "diffuse", "emission": [1.0, 4.0, 9.0], "two-sided": False}])
    scene = Scene.LoadString(text)
    desc = scene.desc.contents
    oracle = OracleScene(scene.desc)
    rng = np.random.default_rng(21)
    checked = {0: 0, 1: 0}
    for _ in range(600):
        k = int(rng.integers(0, 2))
        inst = desc.instances[k]
        lo, hi = np.array(inst.bounds_min[:]), np.array(inst.bounds_max[:])
        origin = rng.uniform(-10, 10, size=3)
        target = lo + (hi - lo) * rng.random(3)
        d = (target - origin) / np.linalg.norm(target - origin)
        ray = np.zeros(1, dtype=RAY_DTYPE)
        ray[0] = (origin.astype(np.float32), d.astype(np.float32), np.inf)
        got = oracle.light_eval_hit(ray)
        if got is None:
            continue
        hit = oracle.trace(ray)[0]
        _, g = oracle.closest_geom(ray)
        p, n, t, b = (g[j].astype(np.float64) for j in range(4))
        hit_inst = desc.instances[int(hit["instance"])]
        L, pdf, cos_wo = ref_light_eval(desc, hit_inst, ray["origin"][0].astype(np.float64), p, n, t, b)
        if abs(cos_wo) < 1e-3:
            continue
        np.testing.assert_allclose(got[:3], L, rtol=1e-6)
        assert got[3] == pytest.approx(pdf, rel=5e-4)
        checked[hit_inst.prim_kind] += 1
    assert checked[0] > 50 and checked[1] > 50, checked
