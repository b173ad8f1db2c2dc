"""Inputs of integration/fsharp/ParityDump.fs (the dump a box with .NET 9 produces from the REFERENCE itself): for each scene a
fixed ray batch (<name>.rays.bin, BnRay records) and a radiance window (<name>.window.txt).  Deterministic (seeds below);
re-running reproduces the committed files byte for byte.   usage: python tests/golden/make_dotnet_inputs.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from barnacle_b200.scene import Scene, make_params          # noqa: E402
from conftest import random_rays                              # noqa: E402
from oracle.oracle_ffi import OracleScene                     # noqa: E402
from test_hostsim import _adversarial                         # noqa: E402

# name -> (scene file, width, height, radiance window x0 y0 x1 y1)
SCENES = {"cbox_ref": ("cbox_ref.json", 1024, 768, (480, 350, 512, 382)), "cbox_bunny": ("cbox_bunny.json", 1024, 1024, (600, 560, 632, 592))}

if __name__ == "__main__":
    out = os.path.join(HERE, "dotnet")
    os.makedirs(out, exist_ok=True)
    for name, (fname, w, h, win) in SCENES.items():
        scene = Scene.Load(os.path.join(ROOT, "scenes", fname), base_dir=ROOT)
        oracle = OracleScene(scene.desc)
        sub = make_params(w, h, 1, rect=(w // 2 - 24, h // 2 - 24, w // 2 + 24, h // 2 + 24))
        rays = np.concatenate([oracle.primary_rays(sub), random_rays(scene, 4096, seed=1234), _adversarial(scene, 2048, seed=1235)])
        rays[len(rays) // 2:]["tmax"][::3] = 60.0            # some finite tmax values (the any-hit dump uses them)
        rays.tofile(os.path.join(out, name + ".rays.bin"))
        with open(os.path.join(out, name + ".window.txt"), "w") as f:
            f.write(f"{w} {h} 2 {win[0]} {win[1]} {win[2]} {win[3]} 8 5\n")
        print(name, len(rays), "rays")
