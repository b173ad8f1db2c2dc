"""GPU parity on randomised scenes (north_star gate 1 on tree shapes the four modelled scenes do not have): triangle soups and
spheres under rotation and non-uniform scale, primitives shared by several instances, TLAS leaves with several instances,
1 .. 300 instances (<= 16: the ordered small-TLAS scan; more: the tree walk).  The oracle is pinned on the very same scenes
against a float64 brute force (tests/test_oracle_bruteforce.py) and the kernels' device code, run on the host, returns the
oracle's hits and films on them (tests/test_hostsim.py); here bn_trace and bn_render do, on the B200.  Named to run last."""
import numpy as np
import pytest

from barnacle_b200.scene import Scene, make_params
from conftest import random_rays
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene
from test_gpu_trace_parity import assert_closest_equal
from test_hostsim import _adversarial
from test_oracle_bruteforce import _random_scene_json

pytestmark = pytest.mark.gpu


def _aimed_rays(scene, rng, n, seed):
    desc = scene.desc.contents
    rays = random_rays(scene, n, seed=seed)
    pick = rng.integers(0, desc.instance_count, size=n)
    lo = np.array([desc.instances[int(k)].bounds_min[:] for k in pick], dtype=np.float64)
    hi = np.array([desc.instances[int(k)].bounds_max[:] for k in pick], dtype=np.float64)
    d = lo + (hi - lo) * rng.random((n, 3)) - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return rays


@pytest.mark.parametrize("seed,n_instances", [(201, 2), (202, 15), (203, 16), (204, 17), (205, 90), (206, 300)])
def test_hits_on_randomised_scenes(seed, n_instances):
    rng = np.random.default_rng(seed)
    scene = Scene.LoadString(_random_scene_json(rng, n_instances))
    oracle, gpu = OracleScene(scene.desc), scene.gpu()
    for rays in (_aimed_rays(scene, rng, 20000, seed), _adversarial(scene, 6000, seed + 1)):
        want = oracle.trace(rays)
        assert_closest_equal(scene, gpu.trace(rays), want)
        tm = rays.copy()
        tm["tmax"] = np.where(want["instance"] >= 0, want["t"] * np.float32(1.5), np.float32(50.0))
        tm["tmax"][::2] = np.where(want["instance"][::2] >= 0, want["t"][::2] * np.float32(0.5), np.float32(5.0))
        assert np.array_equal(gpu.trace(tm, any_hit=True)["instance"], oracle.trace(tm, any_hit=True)["instance"])
    scene.close()


@pytest.mark.parametrize("seed,n_instances,emitters", [(227, 6, "quad"), (217, 40, "quad"), (228, 40, "quad"), (227, 6, "mixed"), (217, 40, "mixed")])   # seeds whose camera sees lit geometry
def test_films_on_randomised_scenes(seed, n_instances, emitters):
    oracle_ffi.set_portable_math(True)
    scene = Scene.LoadString(_random_scene_json(np.random.default_rng(seed), n_instances, emitters))
    oracle, gpu = OracleScene(scene.desc), scene.gpu()
    p = make_params(96, 64, 4, max_depth=6, rr_depth=3)
    film, st = gpu.render(p)
    want, ost = oracle.render(p, counters=True)
    same = (film.view(np.uint32) == want.view(np.uint32)) | (np.isnan(film) & np.isnan(want))
    assert same.all(), f"{int((~same).sum())} film values differ"
    assert np.nan_to_num(want).any()
    assert (st.extend_rays, st.shadow_rays_ref, st.shadow_rays) == (ost["extend_rays"], ost["shadow_rays"], ost["shadow_rays_nonnull"])
    scene.close()
