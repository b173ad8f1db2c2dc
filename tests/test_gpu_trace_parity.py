"""Hit parity on fixed ray batches (north_star gate 1): closest-hit instance and
primitive indices from the CUDA traversal must equal the oracle's bit-exactly; t
and barycentrics are compared BITWISE as well for triangles (stronger than the
1e-5 relative bar) and within 1e-5 for sphere uv (libdevice vs glibc atan2/acos).
All calls go through the C ABI (bn_trace)."""
import numpy as np
import pytest

from conftest import load_scene, random_rays, scene_aabb
from barnacle_b200.scene import RAY_DTYPE, make_params
from oracle.oracle_ffi import OracleScene

pytestmark = pytest.mark.gpu

SCENES = ["cbox_pt", "cbox_bunny", "material_sweep", "bunny_instanced_small"]


def is_sphere_inst(scene):
    d = scene.desc.contents
    return np.array([d.instances[i].prim_kind == 1 for i in range(d.instance_count)])


def assert_closest_equal(scene, g, o):
    assert np.array_equal(g["instance"], o["instance"]), f"instance mismatches: {(g['instance'] != o['instance']).sum()}"
    assert np.array_equal(g["primitive"], o["primitive"]), f"primitive mismatches: {(g['primitive'] != o['primitive']).sum()}"
    assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32)), "t differs bitwise"
    hit = o["instance"] >= 0
    sph = np.zeros(len(o), dtype=bool)
    sph[hit] = is_sphere_inst(scene)[o["instance"][hit]]
    tri = hit & ~sph
    assert np.array_equal(g["u"][tri].view(np.uint32), o["u"][tri].view(np.uint32))
    assert np.array_equal(g["v"][tri].view(np.uint32), o["v"][tri].view(np.uint32))
    np.testing.assert_allclose(g["u"][sph], o["u"][sph], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g["v"][sph], o["v"][sph], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", SCENES)
def test_primary_rays_closest(name):
    """Batch (i): the config's own primary rays, sampleId 0..3."""
    scene = load_scene(name)
    W, H = (96, 96) if name != "bunny_instanced_small" else (128, 72)
    p = make_params(W, H, 4)
    oracle = OracleScene(scene.desc)
    rays = oracle.primary_rays(p)
    o = oracle.trace(rays)
    g = scene.gpu().trace(rays)
    assert (o["instance"] >= 0).mean() > 0.5
    assert_closest_equal(scene, g, o)


@pytest.mark.parametrize("name", SCENES)
def test_random_rays_closest_and_any(name):
    """Batch (ii): 2^18 seeded rays, origins uniform in the scene AABB, uniform directions."""
    scene = load_scene(name)
    rays = random_rays(scene, 1 << 18, seed=20240 + len(name))
    oracle = OracleScene(scene.desc)
    o = oracle.trace(rays)
    g = scene.gpu().trace(rays)
    assert_closest_equal(scene, g, o)
    # any-hit with finite tmax: half the closest distance + noise, so both outcomes occur
    rays2 = rays.copy()
    rng = np.random.Generator(np.random.PCG64(99))
    tm = np.where(np.isfinite(o["t"]), o["t"], 100.0) * rng.uniform(0.2, 1.8, size=len(rays)).astype(np.float32)
    rays2["tmax"] = tm.astype(np.float32)
    oa = oracle.trace(rays2, any_hit=True)
    ga = scene.gpu().trace(rays2, any_hit=True)
    assert 0.1 < oa["instance"].mean() < 0.9
    assert np.array_equal(ga["instance"], oa["instance"])


def test_secondary_rays_bunny():
    """Batch (iii): rays leaving first-hit points (no origin offset — exercises the 1e-3 slab tMin)."""
    scene = load_scene("cbox_bunny")
    oracle = OracleScene(scene.desc)
    rays = oracle.primary_rays(make_params(128, 128, 2))
    o = oracle.trace(rays)
    hit = o["instance"] >= 0
    rng = np.random.Generator(np.random.PCG64(5))
    sec = np.zeros(hit.sum(), dtype=RAY_DTYPE)
    sec["origin"] = (rays["origin"][hit] + o["t"][hit, None] * rays["direction"][hit]).astype(np.float32)
    d = rng.normal(size=(hit.sum(), 3))
    sec["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    sec["tmax"] = np.inf
    assert_closest_equal(scene, scene.gpu().trace(sec), oracle.trace(sec))
    sec["tmax"] = rng.uniform(1, 150, size=len(sec)).astype(np.float32)
    assert np.array_equal(scene.gpu().trace(sec, any_hit=True)["instance"], oracle.trace(sec, any_hit=True)["instance"])


@pytest.mark.parametrize("name", ["cbox_pt", "cbox_bunny"])
def test_adversarial_rays(name):
    """Batch (iv): axis-parallel directions (zero components, +-0), origins on box faces
    (0*inf = NaN slab lanes), rays through shared edges / vertices."""
    scene = load_scene(name)
    lo, hi = scene_aabb(scene)
    rng = np.random.Generator(np.random.PCG64(11))
    n = 20000
    rays = np.zeros(n, dtype=RAY_DTYPE)
    org = lo + (hi - lo) * rng.random((n, 3), dtype=np.float32)
    # snap some origin coordinates onto wall / box planes
    planes = np.array([1.0, 99.0, 0.0, 81.6, 181.0, 50.0, 40.0, 60.0], dtype=np.float32)
    snap = rng.random((n, 3)) < 0.3
    org[snap] = rng.choice(planes, size=snap.sum())
    rays["origin"] = org
    d = np.zeros((n, 3), dtype=np.float32)
    axis = rng.integers(0, 3, size=n)
    sign = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=n)
    d[np.arange(n), axis] = sign
    two = rng.random(n) < 0.3           # diagonal in a plane: still one exact-zero component
    ax2 = (axis + 1) % 3
    d[two, ax2[two]] = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=two.sum())
    negz = rng.random((n, 3)) < 0.5     # sprinkle negative zeros
    d = np.where((d == 0) & negz, np.float32(-0.0), d)
    rays["direction"] = d
    rays["tmax"] = np.inf
    # rays aimed exactly at quad corners / the shared diagonal of the floor quad
    k = 2000
    tgt = np.zeros((k, 3), dtype=np.float32)
    s = rng.random(k).astype(np.float32)
    tgt[:, 0] = 1.0 + 98.0 * s
    tgt[:, 2] = 181.0 * s               # floor diagonal (0,1,2)/(0,2,3) shared edge
    o2 = np.tile(np.array([50.0, 40.0, 90.0], dtype=np.float32), (k, 1))
    rays["origin"][:k] = o2
    rays["direction"][:k] = tgt - o2
    oracle = OracleScene(scene.desc)
    assert_closest_equal(scene, scene.gpu().trace(rays), oracle.trace(rays))
    rays["tmax"] = rng.uniform(0.5, 200, size=n).astype(np.float32)
    assert np.array_equal(scene.gpu().trace(rays, any_hit=True)["instance"], oracle.trace(rays, any_hit=True)["instance"])


def test_empty_batch_and_miss():
    scene = load_scene("cbox_pt")
    g = scene.gpu()
    assert len(g.trace(np.zeros(0, dtype=RAY_DTYPE))) == 0
    rays = np.zeros(3, dtype=RAY_DTYPE)
    rays["origin"] = [[500, 500, 500]] * 3
    rays["direction"] = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    rays["tmax"] = np.inf
    h = g.trace(rays)
    assert (h["instance"] == -1).all() and np.isinf(h["t"]).all()
