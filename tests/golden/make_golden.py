#!/usr/bin/env python3
"""Generates the golden fixtures in this directory from the ORACLE (portable-math mode),
so that (a) the CPU suite notices if the oracle's results ever drift and (b) the GPU suite
can compare against committed vectors without trusting a freshly built oracle.

    python tests/golden/make_golden.py

Fixtures (all small):
  film_<scene>.npy       linear fp32 film, [H, W, 3]
  radiance_<scene>.npy   per-path radiance [spp, H, W, 3]
  hits_<scene>.npz       seeded ray batch (rays) + closest hits + any-hit flags
The reference itself has no tests or golden vectors: these pin the C++
restatement, not the F# program.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {"cbox_pt": (48, 48, 4), "cbox_bunny": (48, 48, 4), "material_sweep": (64, 36, 4), "bunny_instanced_small": (64, 36, 2)}


def main():
    from conftest import load_scene, random_rays
    from barnacle_b200.scene import make_params
    from oracle.oracle_ffi import OracleScene, set_portable_math
    set_portable_math(True)
    for name, (w, h, spp) in CASES.items():
        scene = load_scene(name)
        o = OracleScene(scene.desc)
        p = make_params(w, h, spp)
        film, st = o.render(p)
        np.save(os.path.join(HERE, f"film_{name}.npy"), film.reshape(h, w, 3))
        np.save(os.path.join(HERE, f"radiance_{name}.npy"), o.render_radiance(p))
        rays = random_rays(scene, 4096, seed=777)
        closest = o.trace(rays)
        rays_any = rays.copy()
        rays_any["tmax"] = np.where(np.isfinite(closest["t"]), closest["t"], 50.0) * np.float32(1.2)
        rays_any["tmax"][::2] *= np.float32(0.5)
        anyh = o.trace(rays_any, any_hit=True)
        np.savez_compressed(os.path.join(HERE, f"hits_{name}.npz"), rays=rays, closest=closest, rays_any=rays_any, any=anyh["instance"].astype(np.uint8))
        print(name, st["extend_rays"], st["shadow_rays"], float(np.nanmean(film)))


if __name__ == "__main__":
    main()
