"""The GPU path's own device functions, compiled for the HOST (tests/hostsim/), against the oracle — bit for bit.

traverse.cuh's per-lane walk in its fast and exact slab forms over the converted device layout (64-B two-box nodes,
pre-gathered triangles, pseudo nodes for multi-instance TLAS leaves), and shade.cuh's camera, material and light-sampler
functions, are plain C++ once the CUDA intrinsics are spelled as the IEEE operations they are (device_shim.h).  This
gives a CPU-side regression net for the arithmetic and the data layout of the CUDA path: an edit that breaks parity
shows up here, before any GPU time is spent.  It does NOT replace the -m gpu tests (nvcc's code generation, the
warp-synchronous phase loop and the wavefront are only exercised on the B200), and it is not a fallback: nothing in
barnacle_b200/ can reach this code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from barnacle_b200 import _ffi
from barnacle_b200.scene import RAY_DTYPE, Scene, make_params
from conftest import random_rays
from oracle import oracle_ffi
from oracle.oracle_ffi import HIT_DTYPE, OracleScene
from test_oracle_bruteforce import _random_scene_json

HERE = os.path.dirname(os.path.abspath(__file__))


def _build_hostsim(root, tag="", defines=()):
    src = os.path.join(HERE, "hostsim")
    out_dir = os.path.join(src, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib_path = os.path.join(out_dir, f"libhostsim{tag}.so")
    csrc = os.path.join(root, "barnacle_b200", "csrc", "cuda")
    deps = [os.path.join(src, f) for f in ("hostsim.cpp", "device_shim.h")] + \
           [os.path.join(csrc, f) for f in ("traverse.cuh", "shade.cuh", "vecmath.cuh", "device_scene.h", "scene_convert.cpp", "scene_convert.h", "traverse_limits.h")] + \
           [os.path.join(root, "include", "bn_portable_math.h")]
    if not os.path.exists(lib_path) or any(os.path.getmtime(d) > os.path.getmtime(lib_path) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include", *["-D" + d for d in defines],
                        os.path.join(src, "hostsim.cpp"), os.path.join(csrc, "scene_convert.cpp"), "-o", lib_path], check=True)
    lib = C.CDLL(lib_path)
    lib.hs_scene_create.restype = C.c_void_p
    lib.hs_scene_create.argtypes = [C.c_void_p]
    lib.hs_scene_destroy.argtypes = [C.c_void_p]
    lib.hs_last_error.restype = C.c_char_p
    lib.hs_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.hs_camera_ray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.hs_material_eval.argtypes = [C.c_void_p] * 4
    lib.hs_material_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    lib.hs_light_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    lib.hs_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_xxhash32_three.restype = C.c_uint32
    lib.hs_xxhash32_three.argtypes = [C.c_uint32] * 3
    lib.hs_lcg.restype = C.c_float
    lib.hs_lcg.argtypes = [C.POINTER(C.c_uint32)]
    return lib


@pytest.fixture(scope="module")
def hs(root):
    return _build_hostsim(root)


@pytest.fixture(scope="module")
def hs_shared_rcp(root):
    """The same functions built with the default-off shade experiment -DBN_EXP_SHARED_RCP (vecmath.cuh)."""
    return _build_hostsim(root, "_shared_rcp", ["BN_EXP_SHARED_RCP"])


class HostScene:
    def __init__(self, hs, scene):
        self.hs = hs
        self.h = hs.hs_scene_create(C.cast(scene.desc, C.c_void_p))
        assert self.h, hs.hs_last_error()

    def trace(self, rays, any_hit=False, mode=1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        fast = C.c_uint64(0)
        assert self.hs.hs_trace(self.h, rays.ctypes.data, len(rays), 1 if any_hit else 0, mode, hits.ctypes.data, C.byref(fast)) == 0
        return hits, fast.value

    def render(self, params):
        """(radiance [spp, h, w, 3], film [h*w, 3], extend rays, shadow rays) through the kernels' device functions."""
        rad = np.zeros((params.spp, params.height, params.width, 3), dtype=np.float32)
        film = np.zeros((params.height * params.width, 3), dtype=np.float32)
        st = np.zeros(2, dtype=np.uint64)
        assert self.hs.hs_render(self.h, C.byref(params), rad.ctypes.data, film.ctypes.data, st.ctypes.data) == 0
        return rad, film, int(st[0]), int(st[1])

    def __del__(self):
        self.hs.hs_scene_destroy(self.h)


def _same_hits(desc, a, b):
    assert np.array_equal(a["instance"], b["instance"])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    hit = a["instance"] >= 0
    mesh = np.zeros(len(a), dtype=bool)
    mesh[hit] = [desc.instances[int(i)].prim_kind == 0 for i in a["instance"][hit]]
    assert np.array_equal(a["primitive"][mesh], b["primitive"][mesh])
    assert np.array_equal(a["u"][mesh].view(np.uint32), b["u"][mesh].view(np.uint32))
    assert np.array_equal(a["v"][mesh].view(np.uint32), b["v"][mesh].view(np.uint32))


def _adversarial(scene, n, seed):
    """Axis-parallel directions (exact and signed zeros), origins on box faces: the NaN lanes of the slab test."""
    rng = np.random.default_rng(seed)
    rays = random_rays(scene, n, seed=seed)
    d = rays["direction"].copy()
    zero = rng.random((n, 3)) < 0.4
    d[zero] = 0.0
    d[zero & (rng.random((n, 3)) < 0.5)] = -0.0
    d[np.all(d == 0, axis=1)] = [0.0, 0.0, 1.0]
    rays["direction"] = d
    desc = scene.desc.contents
    pick = rng.integers(0, desc.instance_count, size=n)
    face = rng.random(n) < 0.5
    for k in np.flatnonzero(face):                                        # origin coordinate exactly on an instance box plane
        inst = desc.instances[int(pick[k])]
        ax = int(rng.integers(0, 3))
        rays["origin"][k, ax] = inst.bounds_min[ax] if rng.random() < 0.5 else inst.bounds_max[ax]
    return rays


@pytest.mark.parametrize("name", ["cbox_pt", "cbox_bunny", "material_sweep", "bunny_instanced_small"])
def test_device_traversal_on_the_host_equals_the_oracle(hs, scene_loader, name):
    scene = scene_loader(name)
    desc = scene.desc.contents
    oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
    batches = {"primary": oracle.primary_rays(make_params(48, 48, 1)), "random": random_rays(scene, 6000, seed=3),
               "adversarial": _adversarial(scene, 4000, seed=4)}
    for label, rays in batches.items():
        want = oracle.trace(rays)
        for mode in (0, 1):
            got, fast = host.trace(rays, mode=mode)
            _same_hits(desc, got, want)
            if mode == 1 and label != "adversarial":
                assert fast > 0.99 * len(rays)                            # the fast slab form carries almost every ray
        if label == "adversarial":
            _, fast = host.trace(rays, mode=1)
            assert fast < 0.9 * len(rays)                                 # ... and these really take the exact form
        tm = rays.copy()
        tm["tmax"] = np.where(want["instance"] >= 0, want["t"] * np.float32(1.5), np.float32(50.0))
        tm["tmax"][::2] = np.where(want["instance"][::2] >= 0, want["t"][::2] * np.float32(0.5), np.float32(5.0))
        want_any = oracle.trace(tm, any_hit=True)["instance"]
        for mode in (0, 1):
            assert np.array_equal(host.trace(tm, any_hit=True, mode=mode)[0]["instance"], want_any)


@pytest.mark.parametrize("seed,n_instances", [(31, 2), (32, 14), (33, 60), (34, 250)])
def test_device_traversal_on_randomised_scenes(hs, lib, seed, n_instances):
    rng = np.random.default_rng(seed)
    scene = Scene.LoadString(_random_scene_json(rng, n_instances))
    desc = scene.desc.contents
    oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
    rays = random_rays(scene, 3000, seed=seed)
    pick = rng.integers(0, desc.instance_count, size=len(rays))
    lo = np.array([desc.instances[int(k)].bounds_min[:] for k in pick], dtype=np.float64)
    hi = np.array([desc.instances[int(k)].bounds_max[:] for k in pick], dtype=np.float64)
    d = lo + (hi - lo) * rng.random((len(rays), 3)) - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    want = oracle.trace(rays)
    assert (want["instance"] >= 0).mean() > 0.2
    for mode in (0, 1):
        _same_hits(desc, host.trace(rays, mode=mode)[0], want)
    adv = _adversarial(scene, 2000, seed=seed + 100)
    want = oracle.trace(adv)
    for mode in (0, 1):
        _same_hits(desc, host.trace(adv, mode=mode)[0], want)


def test_shared_reciprocal_experiment_keeps_the_shading_functions_bit_identical(hs_shared_rcp, scene_loader, oracle_lib):
    """normalize() through one shared reciprocal: every vector division in the camera, material and
    light-sampler functions goes through it, and every result keeps its bits."""
    test_device_shading_functions_on_the_host_equal_the_oracle(hs_shared_rcp, scene_loader, oracle_lib)


def test_device_shading_functions_on_the_host_equal_the_oracle(hs, scene_loader, oracle_lib):
    oracle_ffi.set_portable_math(True)                                    # the fixed fp32 transcendentals the kernels evaluate
    rng = np.random.default_rng(9)
    # hash + LCG
    for _ in range(200):
        x, y, z = (int(v) for v in rng.integers(0, 2 ** 32, size=3))
        assert hs.hs_xxhash32_three(x, y, z) == oracle_lib.bo_xxhash32_three(x, y, z)
    a, b = C.c_uint32(12345), C.c_uint32(12345)
    for _ in range(100):
        assert hs.hs_lcg(C.byref(a)) == oracle_lib.bo_lcg(C.byref(b)) and a.value == b.value
    # materials
    unit = lambda: (lambda v: (v / np.linalg.norm(v)).astype(np.float32))(rng.normal(size=3))
    for kind, p0, p1 in ((_ffi.BN_MAT_LAMBERTIAN, 0, 0), (_ffi.BN_MAT_MIRROR, 0, 0), (_ffi.BN_MAT_DIELECTRIC, 1.5, 0), (_ffi.BN_MAT_DIELECTRIC, 1.1, 0),
                         (_ffi.BN_MAT_PBR, 0.0, 0.16), (_ffi.BN_MAT_PBR, 1.0, 0.0025), (_ffi.BN_MAT_PBR, 0.5, 1.0)):
        m = _ffi.BnMaterial()
        m.type, m.p0, m.p1 = kind, p0, p1
        m.base_color[:] = [0.8, 0.5, 0.3]
        for _ in range(300):
            wo, wi, u = unit(), unit(), rng.random(3, dtype=np.float32)
            got = np.zeros(4, np.float32)
            hs.hs_material_eval(C.addressof(m), wo.ctypes.data, wi.ctypes.data, got.ctypes.data)
            assert np.array_equal(got.view(np.uint32), oracle_ffi.material_eval(m, wo, wi).view(np.uint32))
            got = np.zeros(7, np.float32)
            hs.hs_material_sample(C.addressof(m), wo.ctypes.data, float(u[0]), u[1:].ctypes.data, got.ctypes.data)
            assert np.array_equal(got.view(np.uint32), oracle_ffi.material_sample(m, wo, float(u[0]), u[1:]).view(np.uint32))
    # camera + light sampler over whole scenes
    for name in ("cbox_pt", "material_sweep"):
        scene = scene_loader(name)
        oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
        for _ in range(300):
            x, y = int(rng.integers(0, 64)), int(rng.integers(0, 48))
            u = rng.random(4, dtype=np.float32)
            got = np.zeros(1, dtype=RAY_DTYPE)
            hs.hs_camera_ray(host.h, 64, 48, x, y, u.ctypes.data, got.ctypes.data)
            want = oracle.camera_ray(64, 48, x, y, u[:2], u[2:])
            assert got.tobytes() == want.tobytes()
            p = rng.uniform(-50, 150, size=3).astype(np.float32)
            ul = rng.random(3, dtype=np.float32)
            got = np.zeros(10, np.float32)
            hs.hs_light_sample(host.h, p.ctypes.data, float(ul[0]), ul[1:].ctypes.data, got.ctypes.data)
            assert np.array_equal(got.view(np.uint32), oracle.light_sample(p, float(ul[0]), ul[1:]).view(np.uint32))


def _bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


@pytest.mark.parametrize("name,w,h,spp,max_depth,rr_depth", [("cbox_pt", 40, 40, 3, 8, 5), ("material_sweep", 48, 27, 3, 8, 2), ("cbox_bunny", 32, 32, 2, 6, 3),
                                                       ("bunny_instanced_small", 32, 18, 2, 5, 5)])
def test_whole_paths_through_the_device_functions_equal_the_oracle(hs, scene_loader, oracle_lib, name, w, h, spp, max_depth, rr_depth):
    """raygen -> (extend -> shade_lane -> shadow + connect) x bounces -> accumulate, with the kernels' own device code on the
    host: per-path radiance, film and ray counts are the oracle's, bit for bit."""
    oracle_ffi.set_portable_math(True)
    scene = scene_loader(name)
    oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
    for integrator in (_ffi.BN_INTEGRATOR_PATH_TRACING, _ffi.BN_INTEGRATOR_DIRECT, _ffi.BN_INTEGRATOR_NORMAL):
        p = make_params(w, h, spp, max_depth=max_depth, rr_depth=rr_depth, integrator=integrator)
        rad, film, n_ext, n_sh = host.render(p)
        assert _bits_equal(rad, oracle.render_radiance(p, threads=1))
        want_film, st = oracle.render(p, threads=1, counters=True)           # instrumented: also counts the non-null NEE rays
        assert _bits_equal(film, want_film)
        assert np.nan_to_num(film).any()
        if integrator == _ffi.BN_INTEGRATOR_PATH_TRACING:
            assert n_ext == st["extend_rays"] and n_sh == st["shadow_rays_nonnull"]   # null-contribution NEE rays are not traced


def test_shared_reciprocal_experiment_keeps_whole_paths_bit_identical(hs_shared_rcp, scene_loader, oracle_lib):
    """DESIGN.md §8 item 7 at the level of the image: every normalize() of the shading code through the shared reciprocal,
    same film."""
    for name, w, h, spp in (("cbox_pt", 40, 40, 3), ("material_sweep", 48, 27, 3)):
        test_whole_paths_through_the_device_functions_equal_the_oracle(hs_shared_rcp, scene_loader, oracle_lib, name, w, h, spp, 8, 3)


# ---- the warp-synchronous persistent loop itself, on an emulated warp (tests/hostsim/warp_emulator.cpp) ----------------
def _build_warp_emulator(root, tag="", defines=()):
    src = os.path.join(HERE, "hostsim")
    out_dir = os.path.join(src, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib_path = os.path.join(out_dir, f"libwarp{tag}.so")
    csrc = os.path.join(root, "barnacle_b200", "csrc", "cuda")
    deps = [os.path.join(src, f) for f in ("warp_emulator.cpp", "device_shim.h")] + \
           [os.path.join(csrc, f) for f in ("traverse.cuh", "vecmath.cuh", "device_scene.h", "scene_convert.cpp", "scene_convert.h", "traverse_limits.h")]
    if not os.path.exists(lib_path) or any(os.path.getmtime(d) > os.path.getmtime(lib_path) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include", *["-D" + d for d in defines],
                        os.path.join(src, "warp_emulator.cpp"), os.path.join(csrc, "scene_convert.cpp"), "-o", lib_path], check=True)
    lib = C.CDLL(lib_path)
    lib.hsw_scene_create.restype = C.c_void_p
    lib.hsw_scene_create.argtypes = [C.c_void_p, C.c_int]
    lib.hsw_scene_destroy.argtypes = [C.c_void_p]
    lib.hsw_has_flat_tlas.argtypes = [C.c_void_p]
    lib.hsw_last_error.restype = C.c_char_p
    lib.hsw_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def warp(root):
    return _build_warp_emulator(root)


@pytest.fixture(scope="module")
def warp_any_unordered(root):
    """The same loop built with -DBN_ANY_UNORDERED=0: shadow rays walk front to back like closest-hit rays (the default since
    round 2 is the unordered any-hit walk, adopted after its A/B on the B200)."""
    return _build_warp_emulator(root, "_any_ordered", ["BN_ANY_UNORDERED=0"])


def _warp_trace(lib, scene, rays, any_hit, flat=True, binary=False):
    h = lib.hsw_scene_create(C.cast(scene.desc, C.c_void_p), (1 if flat else 0) | (2 if binary else 0))
    assert h, lib.hsw_last_error()
    try:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        st = np.zeros(3, dtype=np.uint64)
        assert lib.hsw_trace(h, rays.ctypes.data, len(rays), 1 if any_hit else 0, hits.ctypes.data, st.ctypes.data) == 0, lib.hsw_last_error()
        return hits, int(st[0]), int(st[1]), bool(lib.hsw_has_flat_tlas(h))
    finally:
        lib.hsw_scene_destroy(h)


def _check_persistent_loop(lib, scene, n_random, seed, flat=True, binary=False):
    desc = scene.desc.contents
    oracle = OracleScene(scene.desc)
    batches = [("primary", oracle.primary_rays(make_params(32, 24, 1))), ("random", random_rays(scene, n_random, seed=seed)),
               ("adversarial", _adversarial(scene, n_random // 2, seed=seed + 1)), ("short", random_rays(scene, 37, seed=seed + 2)),
               ("empty", random_rays(scene, 0, seed=0))]
    used_flat = False
    for label, rays in batches:
        want = oracle.trace(rays)
        got, rendezvous, deferred, used_flat = _warp_trace(lib, scene, rays, False, flat, binary)
        _same_hits(desc, got, want)
        if label == "adversarial":
            assert deferred > 0                                            # the NaN-lane rays went through the exact path
        if len(rays):
            assert rendezvous > 0
        tm = rays.copy()
        tm["tmax"] = np.where(want["instance"] >= 0, want["t"] * np.float32(1.5), np.float32(50.0))
        tm["tmax"][::2] = np.where(want["instance"][::2] >= 0, want["t"][::2] * np.float32(0.5), np.float32(5.0))
        got_any = _warp_trace(lib, scene, tm, True, flat, binary)[0]["instance"]
        assert np.array_equal(got_any, oracle.trace(tm, any_hit=True)["instance"])
    return used_flat


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("name,flat", [("cbox_pt", True), ("cbox_pt", False), ("cbox_bunny", True), ("cbox_bunny", False), ("material_sweep", True), ("bunny_instanced_small", True)])
def test_persistent_traversal_loop_on_an_emulated_warp_equals_the_oracle(warp, scene_loader, name, flat, binary):
    """traverse_persistent as written — phase votes, stay loops, refill, small-TLAS scan (the two Cornell-box scenes; also with
    the scan switched off), identity-instance shortcut, deferral — run by 32 fibers in lock step: same hits as the oracle.
    binary = False: the 4-wide nodes (the default fast path); True: the binary two-box nodes (BN_BINARY_NODES)."""
    used_flat = _check_persistent_loop(warp, scene_loader(name), 3000, seed=41, flat=flat, binary=binary)
    assert used_flat == (flat and name in ("cbox_pt", "cbox_bunny"))


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("seed,n_instances", [(51, 3), (52, 16), (53, 17), (54, 120), (55, 300)])
def test_persistent_traversal_loop_on_randomised_scenes(warp, lib, seed, n_instances, binary):
    rng = np.random.default_rng(seed)
    _check_persistent_loop(warp, Scene.LoadString(_random_scene_json(rng, n_instances)), 2000, seed=seed, binary=binary)


def test_wide_nodes_are_really_used(warp, scene_loader, lib):
    """The default conversion yields 4-wide nodes for every modelled scene (so the tests above exercise them), and the
    A/B switch really drops them."""
    warp.hsw_has_wide.argtypes = [C.c_void_p]
    for name in ("cbox_pt", "cbox_bunny", "material_sweep", "bunny_instanced_small"):
        for flag, want in ((1, 1), (3, 0)):
            h = warp.hsw_scene_create(C.cast(scene_loader(name).desc, C.c_void_p), flag)
            assert h and warp.hsw_has_wide(h) == want, name
            warp.hsw_scene_destroy(h)


def test_unordered_any_hit_experiment_gives_the_same_answers(warp_any_unordered, scene_loader, lib):
    """Shadow rays walk left-first whatever the ray direction by default (BN_ANY_UNORDERED=1; covered by every other test of
    this file); with the front-to-back any-hit walk (-DBN_ANY_UNORDERED=0) the occlusion answers are the same oracle's."""
    for name in ("cbox_bunny", "material_sweep", "bunny_instanced_small"):
        _check_persistent_loop(warp_any_unordered, scene_loader(name), 2500, seed=61)
    _check_persistent_loop(warp_any_unordered, Scene.LoadString(_random_scene_json(np.random.default_rng(62), 40)), 2000, seed=62)


@pytest.mark.parametrize("tag,defines", [("_noswz", ["BN_WIDE_SWIZZLE=0"]), ("_swz256", ["BN_WIDE_SWIZZLE=1", "BN_WIDE_LDG256=1"]), ("_ldg256", ["BN_WIDE_SWIZZLE=0", "BN_WIDE_LDG256=1"])])
def test_wide_node_layout_switches_give_the_same_hits(root, scene_loader, lib, tag, defines):
    """The bank-swizzled node layout (BN_WIDE_SWIZZLE: chunk j of node i at j ^ (i & 7), converter and kernel agree; the default
    since the paths are ordered between bounces) against the plain layout (-DBN_WIDE_SWIZZLE=0), and the 256-bit node fetch
    (-DBN_WIDE_LDG256=1, measured on the B200 and left off) on either: same hits, same any-hit answers."""
    variant = _build_warp_emulator(root, tag, defines)
    for name in ("cbox_bunny", "material_sweep", "bunny_instanced_small"):
        _check_persistent_loop(variant, scene_loader(name), 2500, seed=91)
    _check_persistent_loop(variant, Scene.LoadString(_random_scene_json(np.random.default_rng(92), 120)), 2000, seed=92)


def test_stay_refill_experiment_gives_the_same_hits(root, scene_loader, lib):
    """-DBN_EXP_STAY_REFILL=14: a different moment to leave the stay loops, the same hits."""
    variant = _build_warp_emulator(root, "_stay_refill", ["BN_EXP_STAY_REFILL=14"])
    for name in ("material_sweep", "bunny_instanced_small", "cbox_bunny"):
        _check_persistent_loop(variant, scene_loader(name), 2500, seed=71)


def test_scan_leaf_experiment_gives_the_same_hits(root, scene_loader, lib):
    """-DBN_EXP_SCAN_LEAF: single-leaf identity instances (the Cornell-box walls) tested inside the scan phase —
    same hits, same any-hit answers, on the two scenes that use the scan and on small randomised scenes (<= 16 instances)."""
    variant = _build_warp_emulator(root, "_scan_leaf", ["BN_EXP_SCAN_LEAF"])
    for name in ("cbox_pt", "cbox_bunny"):
        assert _check_persistent_loop(variant, scene_loader(name), 3000, seed=81)
    for seed, n in ((82, 5), (83, 15)):
        _check_persistent_loop(variant, Scene.LoadString(_random_scene_json(np.random.default_rng(seed), n)), 2000, seed=seed)


@pytest.mark.parametrize("seed,n_instances,emitters", [(227, 6, "quad"), (217, 40, "quad"), (228, 40, "quad"), (227, 6, "mixed"), (217, 40, "mixed")])   # the scenes tests/test_zgpu_random_scenes.py renders on the B200
def test_whole_paths_on_randomised_scenes(hs, lib, oracle_lib, seed, n_instances, emitters):
    """Paths that start inside overlapping triangle soups, graze spheres under non-uniform scale, leave through gaps: the
    device functions still return the oracle's radiance, film and ray counts ("mixed": with a one-sided sphere emitter and a
    one-sided quad besides the two-sided quad)."""
    oracle_ffi.set_portable_math(True)
    scene = Scene.LoadString(_random_scene_json(np.random.default_rng(seed), n_instances, emitters))
    oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
    p = make_params(96, 64, 2, max_depth=6, rr_depth=3)
    rad, film, n_ext, n_sh = host.render(p)
    assert _bits_equal(rad, oracle.render_radiance(p, threads=1))
    want_film, st = oracle.render(p, threads=1, counters=True)
    assert _bits_equal(film, want_film) and (np.nan_to_num(film).sum(axis=1) > 0).mean() > 0.05
    assert n_ext == st["extend_rays"] and n_sh == st["shadow_rays_nonnull"]



def test_stack_top_in_registers_gives_the_same_hits(root, scene_loader, lib):
    """-DBN_STACK_TOP_REG=1 (measured on the B200): the top entry of the traversal stack in registers, the pop starting the
    load of the entry below without waiting for it — same stack discipline, same hits, same any-hit answers."""
    variant = _build_warp_emulator(root, "_stack_top", ["BN_STACK_TOP_REG=1"])
    for name in ("cbox_bunny", "material_sweep", "bunny_instanced_small"):
        _check_persistent_loop(variant, scene_loader(name), 2500, seed=51)
    _check_persistent_loop(variant, Scene.LoadString(_random_scene_json(np.random.default_rng(52), 120)), 2000, seed=52)


def test_two_entry_pop_gives_the_same_hits(root, scene_loader, lib):
    """-DBN_POP2=1: closest-hit pops look at the two topmost stack entries at once — same stack discipline, same hits."""
    variant = _build_warp_emulator(root, "_pop2", ["BN_POP2=1"])
    for name in ("cbox_bunny", "material_sweep", "bunny_instanced_small"):
        _check_persistent_loop(variant, scene_loader(name), 2500, seed=41)
    _check_persistent_loop(variant, Scene.LoadString(_random_scene_json(np.random.default_rng(42), 120)), 2000, seed=42)


def test_split_phase_refill_gives_the_same_hits(root, scene_loader, lib):
    """-DBN_SPLIT_REFILL=1 (measured on the B200): the cursor's atomic issued at one vote, its result used at the
    next, ONE phase step of the lanes that still hold a ray in between; the claim stays exact.  Every ray is traced exactly once
    and the hits are the oracle's."""
    variant = _build_warp_emulator(root, "_split_refill", ["BN_SPLIT_REFILL=1"])
    for name in ("cbox_pt", "cbox_bunny", "material_sweep", "bunny_instanced_small"):
        _check_persistent_loop(variant, scene_loader(name), 2500, seed=61)
    _check_persistent_loop(variant, Scene.LoadString(_random_scene_json(np.random.default_rng(62), 60)), 2000, seed=62)
