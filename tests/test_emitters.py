"""Emitter branches no modelled scene reaches: sphere emitters under rotation + non-uniform scale
(SphereInstance.Sample / EvalPDF, Sphere.fs:90-126), one-sided lights (DiffuseLight.Eval, Light.fs:49-53), an emitter with a
material, three emitters at once (UniformLightSampler's uSelect remap, Uniform.fs:13-29).  CPU half: the kernels' device
functions on the host (tests/hostsim) against the oracle, bit for bit, plus sanity that the branches are really taken.  GPU
half (-m gpu): bn_render / bn_render_radiance / bn_trace against the oracle, bit for bit."""
import numpy as np
import pytest

import emitter_scenes
from barnacle_b200 import _ffi
from barnacle_b200.scene import make_params
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene
from test_hostsim import HostScene, _bits_equal, hs  # noqa: F401  (fixture)

NAMES = sorted(emitter_scenes.SCENES)
_CACHE = {}


def _scene(name):
    if name not in _CACHE:
        _CACHE[name] = emitter_scenes.load(name)
    return _CACHE[name]


def test_scenes_hold_what_they_claim(lib):
    want = {"sphere_emitter": (1, 1, [1]), "one_sided": (2, 1, [0, 0]), "two_emitters": (3, 1, [1, 0, 1])}
    for name in NAMES:
        d = _scene(name).desc.contents
        lights = [d.instances[int(d.light_instances[k])] for k in range(d.light_instance_count)]
        n_light, n_sphere, sided = want[name]
        assert len(lights) == n_light
        assert sum(1 for i in lights if i.prim_kind == _ffi.BN_PRIM_SPHERE) == n_sphere
        assert sorted(int(d.lights[i.light_id].two_sided) for i in lights) == sorted(sided)
        sph = next(i for i in lights if i.prim_kind == _ffi.BN_PRIM_SPHERE)
        m = np.array(sph.object_to_world[:], dtype=np.float64).reshape(4, 4)[:3, :3]
        sv = np.linalg.svd(m, compute_uv=False)
        assert sv[0] / sv[-1] > 1.25                                       # genuinely non-uniform scale
        assert abs(m - np.diag(np.diag(m))).max() > 0.1                    # ... and rotated


@pytest.mark.parametrize("name", NAMES)
def test_emitter_branches_through_the_device_functions_equal_the_oracle(hs, lib, oracle_lib, name):
    oracle_ffi.set_portable_math(True)
    scene = _scene(name)
    oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
    for integrator in (_ffi.BN_INTEGRATOR_PATH_TRACING, _ffi.BN_INTEGRATOR_DIRECT):
        p = make_params(48, 48, 2, max_depth=6, rr_depth=3, integrator=integrator)
        rad, film, n_ext, n_sh = host.render(p)
        assert _bits_equal(rad, oracle.render_radiance(p, threads=1))
        want_film, st = oracle.render(p, threads=1, counters=True)
        assert _bits_equal(film, want_film)
        assert (np.nan_to_num(film).sum(axis=1) > 0).mean() > 0.5          # the emitters really light the room
        if integrator == _ffi.BN_INTEGRATOR_PATH_TRACING:
            assert n_ext == st["extend_rays"] and n_sh == st["shadow_rays_nonnull"]
    # light sampling with the device code at random shading points: every emitter is picked, sphere samples included
    rng = np.random.default_rng(5)
    d = scene.desc.contents
    import ctypes as C
    seen_zero_L = seen_pos_L = 0
    for _ in range(400):
        pnt = rng.uniform(5, 95, size=3).astype(np.float32)
        ul = rng.random(3, dtype=np.float32)
        got = np.zeros(10, np.float32)
        hs.hs_light_sample(host.h, pnt.ctypes.data, float(ul[0]), ul[1:].ctypes.data, got.ctypes.data)
        want = oracle.light_sample(pnt, float(ul[0]), ul[1:])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        L = want[3:6]
        seen_zero_L += int((L == 0).all())
        seen_pos_L += int((L > 0).any())
    assert seen_pos_L > 50
    if name != "sphere_emitter":
        assert seen_zero_L > 20                                             # one-sided emitters seen from behind give L = 0


def test_one_sided_quad_is_dark_from_above(lib, oracle_lib):
    """DiffuseLight.Eval: wo.z > 0 || twoSided.  The quad's normal is -Y: a ray arriving from above (between quad and ceiling)
    sees no emission from a one-sided quad, one from below does; the two-sided quad emits both ways."""
    oracle_ffi.set_portable_math(True)
    from barnacle_b200.scene import RAY_DTYPE
    for name, lit_from_above in (("one_sided", False), ("two_emitters", True)):
        oracle = OracleScene(_scene(name).desc)
        rays = np.zeros(2, dtype=RAY_DTYPE)
        rays["origin"] = [[50.0, 81.55, 80.0], [50.0, 70.0, 80.0]]         # 5 cm above the quad (below the ceiling) / below it
        rays["direction"] = [[0.0, -1.0, 0.0], [0.0, 1.0, 0.0]]
        rays["tmax"] = np.inf
        above, below = oracle.light_eval_hit(rays[0:1]), oracle.light_eval_hit(rays[1:2])
        assert above is not None and below is not None                     # both rays end on the emitter quad
        assert bool(np.any(above[:3] > 0)) == lit_from_above, (name, above)
        assert np.any(below[:3] > 0)


# ---- GPU half ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_radiance_film_and_ray_counts_bitwise(name):
    oracle_ffi.set_portable_math(True)
    scene = _scene(name)
    oracle, gpu = OracleScene(scene.desc), scene.gpu()
    for integrator in (_ffi.BN_INTEGRATOR_PATH_TRACING, _ffi.BN_INTEGRATOR_DIRECT):
        p = make_params(96, 96, 4, integrator=integrator)
        assert _bits_equal(gpu.render_radiance(p), oracle.render_radiance(p))
        film, st = gpu.render(p)
        want, ost = oracle.render(p, counters=True)
        assert _bits_equal(film, want)
        assert (np.nan_to_num(film).sum(axis=1) > 0).mean() > 0.5
        assert (st.paths, st.extend_rays, st.shadow_rays_ref, st.shadow_rays) == (ost["paths"], ost["extend_rays"], ost["shadow_rays"], ost["shadow_rays_nonnull"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_hits_on_emitter_scenes(name):
    from conftest import random_rays
    from test_gpu_trace_parity import assert_closest_equal
    from test_hostsim import _adversarial
    scene = _scene(name)
    oracle, gpu = OracleScene(scene.desc), scene.gpu()
    for rays in (oracle.primary_rays(make_params(96, 96, 1)), random_rays(scene, 1 << 16, seed=17), _adversarial(scene, 8000, seed=18)):
        want = oracle.trace(rays)
        assert_closest_equal(scene, gpu.trace(rays), want)
        tm = rays.copy()
        tm["tmax"] = np.where(want["instance"] >= 0, want["t"] * np.float32(1.5), np.float32(50.0))
        tm["tmax"][::2] = np.where(want["instance"][::2] >= 0, want["t"][::2] * np.float32(0.5), np.float32(5.0))
        assert np.array_equal(gpu.trace(tm, any_hit=True)["instance"], oracle.trace(tm, any_hit=True)["instance"])
