"""Offline scheduling statistics of the traversal kernel: runs traverse_persistent (the kernel's own code, compiled with
BN_TRAV_STATS) on the host-side warp emulator of tests/hostsim and prints, per phase, how often it ran and how many lanes were
ready — for the committed thresholds and for -D variants — on ray batches shaped like a frame's (primary, bounce-1 and shadow
rays generated with the oracle).  No GPU needed; what it cannot tell is time (use tools/ab.sh on the B200 for that): it
counts warp-steps and a SIMT cost: per-lane work in SASS-instruction units (BN_WORK in traverse.cuh), charged per rendezvous
interval at the MAXIMUM over the 32 lanes plus the intrinsic — what an issue-bound kernel pays for a divergent step.

    python tools/warp_stats.py [scene] [--define BN_STAY_MIN=8 ...]      # several --define groups: one variant each, ';'-separated
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from barnacle_b200.scene import RAY_DTYPE, Scene, make_params  # noqa: E402
from oracle.oracle_ffi import HIT_DTYPE, OracleScene  # noqa: E402

COST = {"N": 75, "T": 90, "E": 70, "S": 50, "vote": 35}   # SASS instructions per phase step / per vote (cuobjdump, rounded)


def build(defines, tag):
    src = os.path.join(ROOT, "tests", "hostsim")
    out = os.path.join(src, "_build", f"libwarp_stats_{tag}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-DBN_TRAV_STATS",
                    *["-D" + d for d in defines], os.path.join(src, "warp_emulator.cpp"),
                    os.path.join(ROOT, "barnacle_b200", "csrc", "cuda", "scene_convert.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.hsw_scene_create.restype = C.c_void_p
    lib.hsw_scene_create.argtypes = [C.c_void_p, C.c_int]
    lib.hsw_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    lib.hsw_stats.argtypes = [C.c_void_p, C.c_int]
    return lib


def frame_like_batches(scene, oracle, n_side=96, seed=1):
    """primary rays; bounce-1 rays (cosine-ish directions off the first hits, in pixel order like the compacted queue);
    shadow rays from the first hits towards points on the emitters (tmax = distance - 1e-3, as PathTracing.fs:45-50)."""
    rng = np.random.default_rng(seed)
    prim = oracle.primary_rays(make_params(n_side, n_side, 1))
    h = oracle.trace(prim)
    hit = h["instance"] >= 0
    p = (prim["origin"][hit] + h["t"][hit, None] * prim["direction"][hit]).astype(np.float32)
    d = rng.normal(size=(hit.sum(), 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b1 = np.zeros(hit.sum(), dtype=RAY_DTYPE)
    b1["origin"], b1["direction"], b1["tmax"] = p, d.astype(np.float32), np.inf
    desc = scene.desc.contents
    lights = [desc.instances[desc.light_instances[k]] for k in range(desc.light_instance_count)]
    lo = np.array([l.bounds_min[:] for l in lights])[rng.integers(0, len(lights), size=len(p))]
    hi = np.array([l.bounds_max[:] for l in lights])[rng.integers(0, len(lights), size=len(p))]
    target = lo + (hi - lo) * rng.random((len(p), 3))
    v = target - p
    dist = np.linalg.norm(v, axis=1)
    sh = np.zeros(len(p), dtype=RAY_DTYPE)
    sh["origin"], sh["direction"], sh["tmax"] = p, (v / dist[:, None]).astype(np.float32), (dist - 1e-3).astype(np.float32)
    return [("primary", prim, False), ("bounce 1", b1, False), ("shadow", sh, True)]


def run(lib, scene, batches):
    h = lib.hsw_scene_create(C.cast(scene.desc, C.c_void_p), 1)
    out = (C.c_ulonglong * 24)()
    rows = []
    for label, rays, any_hit in batches:
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        st = np.zeros(3, dtype=np.uint64)
        lib.hsw_stats(out, 1)
        assert lib.hsw_trace(h, rays.ctypes.data, len(rays), 1 if any_hit else 0, hits.ctypes.data, st.ctypes.data) == 0
        lib.hsw_stats(out, 1)
        v = list(out)
        base = 10 if any_hit else 0
        cnt = dict(zip("NTE", v[base:base + 3]))
        cnt["S"] = v[base + 4]
        lanes = dict(zip("NTE", v[base + 5:base + 8]))
        lanes["S"] = v[base + 9]
        votes = v[base + 3]
        steps = sum(cnt.values())
        cost = sum(cnt[k] * COST[k] for k in cnt) + votes * COST["vote"]
        rows.append((label, len(rays), steps, votes, {k: (cnt[k], lanes[k] / max(cnt[k], 1)) for k in cnt}, sum(lanes.values()) / max(steps, 1), cost / max(len(rays), 1),
                     int(st[2]) / max(len(rays), 1)))
    return rows


def main():
    args = sys.argv[1:]
    name = args[0] if args and not args[0].startswith("--") else "cbox_bunny"
    variants = [("committed", [])]
    for i, a in enumerate(args):
        if a == "--define":
            for group in args[i + 1].split(";"):
                variants.append((group, group.split(",")))
    scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
    oracle = OracleScene(scene.desc)
    batches = frame_like_batches(scene, oracle)
    print(f"scene {name}: " + ", ".join(f"{l} {len(r)} rays" for l, r, _ in batches))
    for tag, defs in variants:
        lib = build(defs, str(abs(hash(tag)) % 10 ** 8))
        print(f"== {tag}")
        for label, n, steps, votes, per, lanes, cost, simt in run(lib, scene, batches):
            ph = "  ".join(f"{k} {c / max(n, 1) * 32:.1f}/ray @{l:.1f}" for k, (c, l) in per.items() if c)
            print(f"  {label:9s} warp-steps*32/ray {steps * 32 / max(n, 1):6.1f}  votes*32/ray {votes * 32 / max(n, 1):5.1f}  lanes/step {lanes:5.2f}  SIMT cost/ray {simt:6.1f} (flat model {cost:5.1f})   [{ph}]")


if __name__ == "__main__":
    main()
