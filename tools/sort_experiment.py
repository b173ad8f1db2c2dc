"""Experiment: how much does ray ordering (octant / octant+Morton) help the traversal kernel?"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from barnacle_b200.scene import Scene, make_params, RAY_DTYPE, HIT_DTYPE
from oracle.oracle_ffi import OracleScene
name = sys.argv[1] if len(sys.argv) > 1 else "cbox_bunny"
scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
g = scene.gpu(); o = OracleScene(scene.desc)
W = 1024
rays = o.primary_rays(make_params(W, W, int(os.environ.get("SORT_SPP", "8"))))            # pixel-major order within sample: coherent
hits = g.trace(rays)
ok = hits["instance"] >= 0
rng = np.random.Generator(np.random.PCG64(3))
sec = np.zeros(ok.sum(), dtype=RAY_DTYPE)
sec["origin"] = (rays["origin"][ok] + hits["t"][ok, None] * rays["direction"][ok]).astype(np.float32)
d = rng.normal(size=(ok.sum(), 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
sec["direction"] = d.astype(np.float32); sec["tmax"] = np.inf
h2 = g.trace(sec); ok2 = h2["instance"] >= 0
ter = np.zeros(ok2.sum(), dtype=RAY_DTYPE)
ter["origin"] = (sec["origin"][ok2] + h2["t"][ok2, None] * sec["direction"][ok2]).astype(np.float32)
d = rng.normal(size=(ok2.sum(), 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
ter["direction"] = d.astype(np.float32); ter["tmax"] = np.inf

def morton(p, bits=5):
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.uint32), 0, (1 << bits) - 1)
    code = np.zeros(len(p), np.uint32)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code

def run(batch, label):
    n = len(batch)
    d_r = torch.from_numpy(batch.view(np.uint8).reshape(n, 28)).cuda()
    d_h = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    best = 1e9
    for _ in range(5):
        ms = g.trace_device(d_r.data_ptr(), n, False, d_h.data_ptr())
        best = min(best, ms)
    print(f"  {label:28s} {n/best/1e6:7.2f} Grays/s  ({best:.3f} ms)")

for nm, b in (("primary", rays), ("secondary", sec), ("tertiary", ter)):
    print(nm, len(b))
    run(b, "as generated")
    octant = (b["direction"][:, 0] > 0).astype(np.uint32) | ((b["direction"][:, 1] > 0).astype(np.uint32) << 1) | ((b["direction"][:, 2] > 0).astype(np.uint32) << 2)
    run(b[np.argsort(octant, kind="stable")], "octant (stable)")
    # what producer-side binning into per-octant chunks would give: a chunk is opened by the first ray that needs it and
    # consumed in the order the chunks were opened
    for CH in (256, 1024, 4096):
        rank = np.zeros(len(b), np.int64)
        for oc in range(8):
            m = octant == oc
            rank[m] = np.arange(m.sum())
        chunk = rank // CH
        key_open = np.zeros(len(b), np.int64)
        for oc in range(8):
            m = np.flatnonzero(octant == oc)
            if len(m) == 0: continue
            first = m[::CH]                      # stream position of the ray that opens each chunk
            key_open[m] = first[chunk[m]]
        run(b[np.lexsort((rank, key_open))], f"octant chunks of {CH}")
    run(b[np.argsort((octant << 9) | morton(b["origin"], 3), kind="stable")], "octant + morton9")
    run(b[np.argsort((octant << 15) | morton(b["origin"]), kind="stable")], "octant + morton15")
    run(b[np.argsort(morton(b["origin"]), kind="stable")], "morton15 only")
    run(b[rng.permutation(len(b))], "random shuffle")
