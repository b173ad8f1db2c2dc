"""PSSMLT throughput (C5): mutations/s and Mrays/s on the GPU vs the CPU restatement."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from barnacle_b200.scene import Scene, make_mlt_params
from barnacle_b200 import _ffi
scene = Scene.Load(os.path.join(ROOT, "scenes", "cbox_mlt.json"), base_dir=ROOT)
i = scene.info
g = scene.gpu()
for chains in (1024, 65536, 262144):
    p = make_mlt_params(i.width, i.height, i.spp, i.max_depth, i.rr_depth, 0, i.n_bootstrap, chains)
    g.render_pssmlt(p)
    t0 = time.perf_counter(); _, st = g.render_pssmlt(p); dt = time.perf_counter() - t0
    print(json.dumps({"chains": chains, "mutations": st.proposed, "accept_rate": st.accepted / st.proposed, "B": st.b, "bootstrap_ms": st.bootstrap_ms, "chains_ms": st.chains_ms,
                      "Mmutations_per_s": st.proposed / st.chains_ms / 1e3, "Mrays_per_s": st.rays / (st.bootstrap_ms + st.chains_ms) / 1e3, "wall_s": dt}))
if "--cpu" in sys.argv:
    from oracle.oracle_ffi import OracleScene, set_portable_math
    set_portable_math(False)
    o = OracleScene(scene.desc)
    p = make_mlt_params(i.width, i.height, 1, i.max_depth, i.rr_depth, 0, 262144, 1024)
    t0 = time.perf_counter(); _, st, _ = o.render_pssmlt(p); dt = time.perf_counter() - t0
    print(json.dumps({"cpu_mutations": st["proposed"], "chain_seconds": st["chain_seconds"], "Mmutations_per_s": st["proposed"] / st["chain_seconds"] / 1e6, "wall_s": dt}))
