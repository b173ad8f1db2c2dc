"""Where the e2e (host-buffer) step spends its time: scene create / render / destroy (debug aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from barnacle_b200.scene import Scene, GpuScene, make_params
name, W, H, SPP = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("cbox_pt", 512, 512, 64)
scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
film = np.zeros(W * H * 3, dtype=np.float32)
p = make_params(W, H, SPP)
for it in range(4):
    t0 = time.perf_counter(); g = GpuScene(scene.desc, 0)
    t1 = time.perf_counter(); _, st = g.render(p, film)
    t2 = time.perf_counter(); g.close()
    t3 = time.perf_counter()
    print(f"iter {it}: create {1e3*(t1-t0):.2f} ms  render {1e3*(t2-t1):.2f} ms (device {st.gpu_ms:.2f})  destroy {1e3*(t3-t2):.2f} ms")
