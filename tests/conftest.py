import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def lib():
    """The product C-ABI library (built in-tree if missing; nvcc cross-compiles without a GPU)."""
    from barnacle_b200 import _ffi, build
    if not os.path.exists(_ffi.LIB_PATH):
        build.build()
    return _ffi.load()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle_ffi
    return oracle_ffi.load()


_SCENES = {}


def load_scene(name):
    """Session-cached host scene (JSON in scenes/)."""
    from barnacle_b200.scene import Scene
    if name not in _SCENES:
        _SCENES[name] = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
    return _SCENES[name]


@pytest.fixture(scope="session")
def scene_loader(lib):
    return load_scene


def scene_aabb(scene):
    d = scene.desc.contents
    n = d.tlas_nodes[0]
    return np.array(n.bounds_min[:], dtype=np.float32), np.array(n.bounds_max[:], dtype=np.float32)


def random_rays(scene, n, seed, tmax=np.inf):
    """Batch (ii) of SURVEY §8(d): origins uniform in the scene AABB, uniform directions (numpy PCG64, seed stated)."""
    from barnacle_b200.scene import RAY_DTYPE
    rng = np.random.Generator(np.random.PCG64(seed))
    lo, hi = scene_aabb(scene)
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["origin"] = (lo + (hi - lo) * rng.random((n, 3), dtype=np.float32)).astype(np.float32)
    z = 1 - 2 * rng.random(n)
    phi = 2 * np.pi * rng.random(n)
    r = np.sqrt(np.maximum(0, 1 - z * z))
    rays["direction"] = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    rays["tmax"] = tmax
    return rays
