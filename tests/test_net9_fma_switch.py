"""BN_NET9_FMA — the one-flag switch between the two candidate conventions of System.Numerics on .NET 9:
Vector3.Cross / Vector3.Transform with fused multiply-adds (default, 1) or in the .NET <= 8 form with separate roundings (0).
Which one the real runtime uses can only be pinned on a box with `dotnet` (integration/fsharp/ParityDump.fs produces the
dump, tests/test_dotnet_dump.py consumes it).  What is proven here is that the switch IS one flag: the oracle, the host-side
stand-in and the kernels built with -DBN_NET9_FMA=0 agree with each other bit for bit exactly as the default builds do —
and that the two conventions really differ in the bits (so a dump can tell them apart).
CPU half: oracle vs the kernels' device functions on the host (tests/hostsim).  GPU half: oracle vs bn_render / bn_trace."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CPU_SCRIPT = r'''
import hashlib, json, os, sys
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from barnacle_b200.scene import Scene, make_params
from conftest import random_rays
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene
from test_hostsim import HostScene, _build_hostsim, _same_hits, _bits_equal
fma0 = os.environ.get("BN_NET9_FMA") == "0"
assert oracle_ffi.LIB_NAME == ("libbarnacle_oracle_fma0.so" if fma0 else "libbarnacle_oracle.so")
hs = _build_hostsim(ROOT, "_fma0" if fma0 else "", ["BN_NET9_FMA=0"] if fma0 else [])
oracle_ffi.set_portable_math(True)
out = {}
for name in ("cbox_bunny", "material_sweep"):
    scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
    oracle, host = OracleScene(scene.desc), HostScene(hs, scene)
    rays = np.concatenate([oracle.primary_rays(make_params(48, 48, 1)), random_rays(scene, 4000, seed=11)])
    want = oracle.trace(rays)
    _same_hits(scene.desc.contents, host.trace(rays, mode=1)[0], want)
    p = make_params(40, 40, 2)
    rad, film, n_ext, n_sh = host.render(p)
    assert _bits_equal(rad, oracle.render_radiance(p, threads=1))
    out[name] = {"hits": hashlib.sha1(want.tobytes()).hexdigest(), "film": hashlib.sha1(film.tobytes()).hexdigest()}
print("RESULT " + json.dumps(out))
'''

_GPU_SCRIPT = r'''
import json, os, sys
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from barnacle_b200.scene import Scene, make_params
from conftest import random_rays
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene
from test_gpu_trace_parity import assert_closest_equal
oracle_ffi.set_portable_math(True)
import hashlib
out = {}
for name in ("cbox_bunny", "material_sweep"):
    scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
    oracle, gpu = OracleScene(scene.desc), scene.gpu()
    rays = np.concatenate([oracle.primary_rays(make_params(96, 96, 1)), random_rays(scene, 1 << 16, seed=11)])
    assert_closest_equal(scene, gpu.trace(rays), oracle.trace(rays))
    p = make_params(64, 64, 4)
    g, o = gpu.render_radiance(p), oracle.render_radiance(p)
    assert ((g.view(np.uint32) == o.view(np.uint32)) | (np.isnan(g) & np.isnan(o))).all()
    out[name] = hashlib.sha1(g.tobytes()).hexdigest()
print("RESULT " + json.dumps(out))
'''


def _fma0_lib():
    """The product library built with -DBN_NET9_FMA=0 (host stand-in and kernels alike), next to the shipped one."""
    from barnacle_b200 import build
    path = os.path.join(build.LIB_DIR, "lib_fma0.so")
    if not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in build._all_inputs()):
        build.build(defines=["BN_NET9_FMA=0"], out="lib_fma0.so")
    return path


def _run(script, fma0):
    env = dict(os.environ)
    env.pop("BN_NET9_FMA", None)
    env.pop("BN_LIB", None)
    if fma0:
        env.update(BN_NET9_FMA="0", BN_LIB=_fma0_lib())
    r = subprocess.run([sys.executable, "-c", script, ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])


def test_both_conventions_are_self_consistent_and_differ(lib, oracle_lib):
    with_fma, without = _run(_CPU_SCRIPT, False), _run(_CPU_SCRIPT, True)
    for name in with_fma:
        assert with_fma[name]["hits"] != without[name]["hits"], "the switch must reach the traversal arithmetic (Ray.Transform, Triangle.Intersect's crosses)"
        assert with_fma[name]["film"] != without[name]["film"]


@pytest.mark.gpu
def test_kernels_built_without_the_net9_fma_match_the_oracle_built_without_it():
    with_fma, without = _run(_GPU_SCRIPT, False), _run(_GPU_SCRIPT, True)
    assert all(with_fma[k] != without[k] for k in with_fma)
