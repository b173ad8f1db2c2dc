"""Debug aid: traversal phase utilisation (needs a --stats build of the library)."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barnacle_b200 import _ffi
from barnacle_b200.scene import Scene, make_params
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1] if len(sys.argv) > 1 else "cbox_bunny"
scene = Scene.Load(os.path.join(ROOT, "scenes", name + ".json"), base_dir=ROOT)
g = scene.gpu()
W, H, SPP = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1024, 1024, 4)
raw = ctypes.CDLL(_ffi.LIB_PATH)
out = (ctypes.c_ulonglong * 24)()
g.render(make_params(W, H, SPP))
raw.bn_debug_trav_stats(out, 1)
_, st = g.render(make_params(W, H, SPP, flags=_ffi.BN_RENDER_PROFILE))
raw.bn_debug_trav_stats(out, 1)
v = list(out)
for label, base, rays, ms in (("extend", 0, st.extend_rays, st.extend_ms), ("shadow", 10, st.shadow_rays, st.shadow_ms)):
    cnt, sm = v[base:base + 5], v[base + 5:base + 10]
    cnt = [cnt[0], cnt[1], cnt[2], cnt[4], cnt[3]]; sm = [sm[0], sm[1], sm[2], sm[4], sm[3]]
    tot = sum(cnt[:4])
    print(f"{label}: rays={rays} ms={ms:.2f} Grays/s={rays/ms/1e6:.2f} warp-steps={tot} steps/ray(warp-steps*32/rays)={tot*32/max(rays,1):.1f}")
    for k, nm in enumerate(("N", "T", "E", "S")):
        if cnt[k]:
            print(f"   {nm}: {cnt[k]/tot*100:5.1f}% of steps, avg ready lanes {sm[k]/cnt[k]:5.2f}/32, lane-steps/ray {sm[k]/max(rays,1):.2f}")
    print(f"   avg active lanes over all steps: {sm[4]/max(cnt[4],1):.2f}; issue-cycles per warp-step: {ms*1e-3*148*4*1.965e9/max(tot,1):.0f}")
