"""Pins the CONTROL FLOW of the oracle's PathTracingIntegrator.Li: a second restatement of
PathTracing.fs:14-81 in Python — the order of the random draws, emitter MIS, next-event estimation with its
shadow ray, BSDF sampling, throughput update, Russian roulette, every termination — driving the oracle's
building blocks (closest / any hit, light sampler, materials: each pinned on its own by the other
tests/test_oracle_*.py) through their test entry points, compared with the radiance the oracle's own Li returns
for the same (pixel, sample) seeds.  fp32 numpy arithmetic in the reference's operation order: measured, all
1 680 paths of the three scenes come out bit-identical; the test asks for agreement to 1e-4 on every path and
bit-identity on 98 % of them (the fused multiply-adds are emulated through float64)."""
import numpy as np
import pytest

from barnacle_b200.scene import RAY_DTYPE, make_params
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene
from test_oracle_camera import PySampler

F = np.float32


def _dot(a, b):  # Vector3.Dot: (x x' + y y') + z z'
    return F(F(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])


def _fma3(a, b, c):  # Vector3.FusedMultiplyAdd
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _ray(o, d, tmax):
    r = np.zeros(1, dtype=RAY_DTYPE)
    r[0] = (o, d, tmax)
    return r


def restated_li(oracle, desc, o, d, sampler, max_depth, rr_depth):
    L, beta = np.zeros(3, F), np.ones(3, F)
    depth, bsdf_pdf = 0, F(0)
    while depth < max_depth:
        ray = _ray(o, d, np.inf)                                        # t <- infinityf (:25)
        hit = oracle.trace(ray)[0]
        if hit["instance"] < 0:                                         # :27-28
            break
        _, g = oracle.closest_geom(ray)
        p, n, t, b = (g[k].astype(F) for k in range(4))
        inst = desc.instances[int(hit["instance"])]
        to_local = lambda v: np.array([_dot(v, t), _dot(v, b), _dot(v, n)], F)       # OrthonormalBasis.WorldToLocal
        to_world = lambda v: (F(v[0]) * t + F(v[1]) * b) + F(v[2]) * n               # LocalToWorld
        if inst.light_id >= 0:                                          # :30-41
            le = oracle.light_eval_hit(ray)
            mis = F(1) if depth == 0 else F(bsdf_pdf * (F(1) / F(le[3] + bsdf_pdf)))
            L = _fma3(beta, le[:3] * mis, L)
        if inst.material_id < 0:                                        # :78-79
            break
        mat = desc.materials[inst.material_id]
        usel, ul = sampler.next1d(), sampler.next2d()                   # :43 (argument order = draw order)
        ls = oracle.light_sample(p, usel, ul)
        ls_p, ls_L, ls_pdf, ls_wi = ls[0:3], ls[3:6], ls[6], ls[7:10]
        diff = ls_p - p
        dist = F(np.sqrt(_dot(diff, diff)))
        wo_local = to_local(-d)
        if ls_pdf != 0 and oracle.trace(_ray(p, ls_wi, F(dist - F(1e-3))), any_hit=True)[0]["instance"] == 0:   # :47-50
            e = oracle_ffi.material_eval(mat, wo_local, to_local(ls_wi))
            L = _fma3(beta * e[:3], ls_L * F(F(1) / F(e[3] + ls_pdf)), L)                                       # :53-59
        bs = oracle_ffi.material_sample(mat, wo_local, sampler.next1d(), sampler.next2d())                      # :61
        bsdf_pdf = bs[3]
        if bsdf_pdf == 0:                                               # :63-64
            break
        o, d = p, to_world(bs[4:7])                                     # :66
        beta = (beta * bs[:3]) * F(F(1) / bsdf_pdf)                     # :67
        if depth >= rr_depth:                                           # :69-75
            thr = min(F(1), max(beta[0], max(beta[1], beta[2])))
            if sampler.next1d() < thr:
                beta = beta * F(F(1) / thr)
            else:
                break
        depth += 1
    return L


@pytest.mark.parametrize("name,max_depth,rr_depth", [("cbox_pt", 8, 5), ("material_sweep", 8, 2), ("cbox_bunny", 4, 5)])
def test_li_control_flow_matches_restatement(scene_loader, oracle_lib, name, max_depth, rr_depth):
    oracle_ffi.set_portable_math(False)
    scene = scene_loader(name)
    desc = scene.desc.contents
    oracle = OracleScene(scene.desc)
    w, h, spp = 20, 14, 2
    p = make_params(w, h, spp, max_depth=max_depth, rr_depth=rr_depth)
    want = oracle.render_radiance(p, threads=1)                         # [spp, h, w, 3]
    rays = oracle.primary_rays(p).reshape(spp, h, w)
    n = bad = lit = exact = 0
    for s in range(spp):
        for y in range(h):
            for x in range(w):
                sampler = PySampler(x, y, s)
                sampler.next2d(), sampler.next2d()                      # the camera's four draws (Integrator.fs:39)
                r = rays[s, y, x]
                got = restated_li(oracle, desc, r["origin"].astype(F), r["direction"].astype(F), sampler, max_depth, rr_depth)
                ref = want[s, y, x]
                ok = np.allclose(got, ref, rtol=1e-4, atol=1e-6, equal_nan=True)
                n += 1
                bad += not ok
                exact += np.array_equal(got.view(np.uint32), ref.view(np.uint32))
                lit += bool(np.nan_to_num(ref).any())
    assert lit > 0.5 * n, "degenerate test: most paths carry no radiance"
    assert bad == 0, f"{bad} of {n} paths differ from the oracle's Li"
    assert exact >= 0.98 * n, f"only {exact} of {n} paths are bit-identical (measured: all of them)"


# ---- the two other ProgressiveIntegrators ("next" row N4) ------------------------------------------------------

def _hit(oracle, desc, o, d):
    ray = _ray(o, d, np.inf)
    hit = oracle.trace(ray)[0]
    if hit["instance"] < 0:
        return None
    _, g = oracle.closest_geom(ray)
    return ray, desc.instances[int(hit["instance"])], tuple(g[k].astype(F) for k in range(4))


def restated_direct_li(oracle, desc, o, d, sampler):  # Direct.fs:10-40
    L = np.zeros(3, F)
    h = _hit(oracle, desc, o, d)
    if h is None:
        return L
    _, inst, (p, n, t, b) = h
    to_local = lambda v: np.array([_dot(v, t), _dot(v, b), _dot(v, n)], F)
    to_world = lambda v: (F(v[0]) * t + F(v[1]) * b) + F(v[2]) * n
    if inst.light_id >= 0:                                              # :16-17, DiffuseLight.Eval (Light.fs:49-53)
        light = desc.lights[inst.light_id]
        woz = to_local(-d)[2]
        if abs(woz) > F(1e-6) and (woz > 0 or light.two_sided):
            L = L + np.array(light.emission[:], F)
    if inst.material_id < 0:
        return L
    mat = desc.materials[inst.material_id]
    ls = oracle.light_sample(p, sampler.next1d(), sampler.next2d())     # :20-21
    diff = ls[0:3] - p
    dist = F(np.sqrt(_dot(diff, diff)))
    wo_local = to_local(-d)
    if oracle.trace(_ray(p, ls[7:10], F(dist - F(1e-3))), any_hit=True)[0]["instance"] == 0:   # traced whatever the pdf (:25)
        e = oracle_ffi.material_eval(mat, wo_local, to_local(ls[7:10]))
        with np.errstate(divide="ignore", invalid="ignore"):
            L = _fma3(e[:3], ls[3:6] * F(F(1) / ls[6]), L)              # no MIS, a true division (:28-29)
    bs = oracle_ffi.material_sample(mat, wo_local, sampler.next1d(), sampler.next2d())          # :31
    h2 = _hit(oracle, desc, p, to_world(bs[4:7]))                       # followed whatever its pdf (:32-36)
    if h2 is not None and h2[1].light_id >= 0:
        le = oracle.light_eval_hit(h2[0])
        with np.errstate(divide="ignore", invalid="ignore"):
            L = _fma3(bs[:3], le[:3] * F(F(1) / le[3]), L)              # weighted by 1 / pdf_light (sic, :37-38)
    return L


@pytest.mark.parametrize("name", ["cbox_pt", "material_sweep"])
def test_direct_and_normal_integrators_match_restatement(scene_loader, oracle_lib, name):
    from barnacle_b200 import _ffi
    oracle_ffi.set_portable_math(False)
    scene = scene_loader(name)
    desc = scene.desc.contents
    oracle = OracleScene(scene.desc)
    w, h, spp = 20, 14, 2
    rays = oracle.primary_rays(make_params(w, h, spp)).reshape(spp, h, w)
    direct = oracle.render_radiance(make_params(w, h, spp, integrator=_ffi.BN_INTEGRATOR_DIRECT), threads=1)
    normal = oracle.render_radiance(make_params(w, h, spp, integrator=_ffi.BN_INTEGRATOR_NORMAL), threads=1)
    n = exact = lit = 0
    for s in range(spp):
        for y in range(h):
            for x in range(w):
                r = rays[s, y, x]
                o, d = r["origin"].astype(F), r["direction"].astype(F)
                hit = _hit(oracle, desc, o, d)
                want_n = np.zeros(3, F) if hit is None else F(0.5) * (hit[2][1] + F(1))      # Normal.fs:14-17
                np.testing.assert_array_equal(normal[s, y, x], want_n)
                sampler = PySampler(x, y, s)
                sampler.next2d(), sampler.next2d()
                got = restated_direct_li(oracle, desc, o, d, sampler)
                ref = direct[s, y, x]
                assert np.allclose(got, ref, rtol=1e-4, atol=1e-6, equal_nan=True), (x, y, s, got, ref)
                n += 1
                exact += np.array_equal(got.view(np.uint32), ref.view(np.uint32))
                lit += bool(np.nan_to_num(ref).any())
    assert lit > 0.5 * n and exact >= 0.98 * n, (lit, exact, n)


def test_film_is_the_fma_chain_of_the_per_sample_radiance(scene_loader, oracle_lib):
    """ProgressiveIntegrator.RenderTile (Integrator.fs:29-44) + Film.SetPixel (Film.fs:41-46): per pixel,
    accum = fma(1/spp, Li * 1/pdf, accum) over sampleId ascending, stored at the Y-flipped index."""
    oracle_ffi.set_portable_math(False)
    scene = scene_loader("cbox_pt")
    oracle = OracleScene(scene.desc)
    w, h, spp = 24, 18, 5
    p = make_params(w, h, spp)
    film, _ = oracle.render(p, threads=1)
    rad = oracle.render_radiance(p, threads=1)                           # [spp, h, w, 3], row y = image row y
    inv = F(F(1) / F(spp))
    acc = np.zeros((h, w, 3), F)
    for s in range(spp):
        acc = (np.float64(inv) * rad[s].astype(np.float64) + acc.astype(np.float64)).astype(F)
    want = acc[::-1].reshape(h * w, 3)                                   # index (H - y - 1) * W + x
    np.testing.assert_array_equal(film.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
