"""Small invocations of every device entry point, meant to run under compute-sanitizer (tools/session3.sh):
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize.py
bn_render (wide and binary nodes, path / direct / normal, a frame split into many small waves), bn_trace (closest / any, with the
adversarial NaN-lane rays that take the exact fix-up kernel), bn_render_radiance, bn_render_pssmlt, bn_bvh_build,
bn_render_multi (the same device twice: threads + the peer-read combine kernel), bn_film_to_rgba8_device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("BN_WAVE_PATHS", "8192")     # several waves per frame, small queues

from barnacle_b200 import _ffi                                          # noqa: E402
from barnacle_b200.scene import MultiGpuScene, Scene, make_mlt_params, make_params   # noqa: E402
from conftest import random_rays                                         # noqa: E402
from test_hostsim import _adversarial                                    # noqa: E402

lib = _ffi.load()
for name in ("cbox_bunny", "material_sweep", "bunny_instanced_small"):
    scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
    for binary in (False, True):
        if binary:
            os.environ["BN_BINARY_NODES"] = "1"
        else:
            os.environ.pop("BN_BINARY_NODES", None)
        g = scene.gpu()
        for integrator in (0, 1, 2):
            film, st = g.render(make_params(40, 24, 2, integrator=integrator))
            assert np.isfinite(film).mean() > 0.9
        g.render(make_params(24, 24, 1, flags=_ffi.BN_RENDER_FORCE_EXACT | _ffi.BN_RENDER_TRACE_NULL_SHADOW | _ffi.BN_RENDER_PROFILE))
        g.render_radiance(make_params(16, 16, 2))
        for rays in (random_rays(scene, 3000, seed=5), _adversarial(scene, 2000, seed=6)):
            g.trace(rays)
            tm = rays.copy()
            tm["tmax"] = 40.0
            g.trace(tm, any_hit=True)
        scene.close()
        scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
    print("render / trace ok:", name, flush=True)
os.environ.pop("BN_BINARY_NODES", None)
scene = Scene.Load(os.path.join(ROOT, "scenes", "cbox_mlt.json"), base_dir=ROOT)
film, st = scene.gpu().render_pssmlt(make_mlt_params(32, 32, 1, 6, 3, 0, 4096, 256))
print("pssmlt ok: B =", st.b, flush=True)
m = MultiGpuScene(scene.desc, [0, 0])
m.render(make_params(48, 40, 3), partition=_ffi.BN_PARTITION_SAMPLE)
m.render(make_params(48, 40, 1), partition=_ffi.BN_PARTITION_TILE)
m.close()
print("multi ok", flush=True)
rng = np.random.default_rng(3)
lo = rng.uniform(-10, 10, size=(3000, 3)).astype(np.float32)
boxes = np.concatenate([lo, lo + rng.uniform(0.01, 1.0, size=(3000, 3)).astype(np.float32)], axis=1)
nodes = (_ffi.BnBVHNode * 6000)()
perm = np.zeros(3000, dtype=np.uint32)
import ctypes as C
n = lib.bn_bvh_build(0, boxes.ctypes.data_as(C.POINTER(C.c_float)), 3000, nodes, 6000, perm.ctypes.data_as(C.POINTER(C.c_uint32)), None)
assert n > 0 and sorted(perm.tolist()) == list(range(3000))
print("bvh build ok:", n, "nodes", flush=True)
scene.close()
print("sanitize.py: all entry points ran", flush=True)
