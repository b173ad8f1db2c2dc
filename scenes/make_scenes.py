#!/usr/bin/env python3
"""Authors the benchmark / parity scenes in the reference's JSON schema
(Extensions/Scene/Loader.fs) from the reference's two assets.

Run in the build container (needs /root/reference); outputs are committed so the
GPU box — which has no /root/reference — can load them:

    python scenes/make_scenes.py

  scenes/assets/bunny.obj        Stanford bunny: the `v`/`f` records of
                                 Asset/stanford-bunny.obj re-serialised (same decimal
                                 values, trailing zeros stripped, comments dropped)
  scenes/cbox_pt.json            C1: Asset/cbox.json with integrator -> path-tracing 64 spp,
                                 film 512x512, camera aspect 1.0
  scenes/cbox_bunny.json         C2: C1 + bunny mesh instance, 1024x1024, 256 spp
  scenes/material_sweep.json     C3: cbox room + pbr (metallic x roughness) and dielectric (ior) sweep, 1920x1080, 1024 spp
  scenes/bunny_instanced.json    C4: 64x64 bunny instances sharing one BLAS, 3840x2160, 512 spp
  scenes/cbox_mlt.json           C5 scene (pssmlt; a "next" row): C2 geometry, pssmlt integrator
  scenes/cbox_ref.json           Asset/cbox.json as published (1024x768, thin lens, aces, pssmlt 16 spp) — the scene of the
                                 reference's two sample PNGs; used by tests/test_ref_sample_image.py
"""
import copy
import json
import os

REF = "/root/reference/Asset"
OUT = os.path.dirname(os.path.abspath(__file__))


def strip(tok: str) -> str:
    if "." in tok and "e" not in tok.lower():
        tok = tok.rstrip("0").rstrip(".")
        if tok in ("", "-"):
            tok = "0"
    return tok


def write_bunny():
    lines = ["# Stanford bunny (bun_zipper), re-serialised v/f records; 35947 vertices, 69451 faces"]
    with open(os.path.join(REF, "stanford-bunny.obj")) as f:
        for line in f:
            line = line.strip()
            if line.startswith("v "):
                lines.append("v " + " ".join(strip(t) for t in line[2:].split()))
            elif line.startswith("f "):
                lines.append("f " + " ".join(line[2:].split()))
    with open(os.path.join(OUT, "assets", "bunny.obj"), "w") as f:
        f.write("\n".join(lines) + "\n")


def dump(name, scene):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(scene, f, separators=(",", ":"))
        f.write("\n")


def load_cbox():
    with open(os.path.join(REF, "cbox.json"), encoding="utf-8-sig") as f:
        return json.load(f)


def add_instance(scene, primitive, material, keyframe, parent=0):
    """Appends instance + transform + node; returns the node index."""
    scene["instances"].append({"primitive": primitive, "material": material})
    scene["transforms"].append({"keyframes": [keyframe]})
    scene["nodes"].append({"instances": [len(scene["instances"]) - 1], "transform": len(scene["transforms"]) - 1})
    scene["nodes"][parent].setdefault("children", []).append(len(scene["nodes"]) - 1)
    return len(scene["nodes"]) - 1


def make_c1():
    s = load_cbox()
    s["integrator"] = {"type": "path-tracing", "spp": 64, "max-depth": 8}
    s["film"] = {"width": 512, "height": 512, "tone-mapping": "aces"}
    s["camera"]["aspect-ratio"] = 1.0
    return s


BUNNY_URI = "scenes/assets/bunny.obj"


def make_c2():
    s = make_c1()
    s["integrator"]["spp"] = 256
    s["film"] = {"width": 1024, "height": 1024, "tone-mapping": "aces"}
    s["primitives"].append({"type": "mesh", "uri": BUNNY_URI})
    s["materials"].append({"type": "lambertian", "albedo": [0.75, 0.75, 0.75]})
    # bbox y_min = 0.0330 -> scale 250 puts the feet at 8.25 above the origin: translate down to the floor
    add_instance(s, len(s["primitives"]) - 1, len(s["materials"]) - 1,
                 {"scale": [250.0, 250.0, 250.0], "translation": [54.0, -8.25, 115.0]})
    return s


def make_c3():
    s = make_c1()
    s["integrator"]["spp"] = 1024
    s["film"] = {"width": 1920, "height": 1080, "tone-mapping": "aces"}
    s["camera"]["aspect-ratio"] = 1.7778
    # drop the original sphere and cube (instances 7, 8 / nodes 2, 3); keep walls + light
    s["nodes"][0]["children"] = [1, 4]
    sphere = len(s["primitives"])
    s["primitives"].append({"type": "sphere", "radius": 5.0})
    cube = 8
    rough = [0.05, 0.15, 0.3, 0.45, 0.6, 0.8, 1.0]
    metal = [0.0, 0.5, 1.0]
    colors = [[0.9, 0.6, 0.2], [0.75, 0.75, 0.75], [0.3, 0.6, 0.9]]
    for r, m in enumerate(metal):
        for c, ro in enumerate(rough):
            s["materials"].append({"type": "pbr", "albedo": colors[r], "roughness": ro, "metallic": m})
            add_instance(s, sphere, len(s["materials"]) - 1, {"translation": [14.0 + 12.0 * c, 24.0 + 14.0 * r, 50.0]})
    iors = [1.1, 1.3, 1.5, 1.7, 2.0]
    for c, ior in enumerate(iors):
        s["materials"].append({"type": "dielectric", "ior": ior})
        mat = len(s["materials"]) - 1
        add_instance(s, sphere, mat, {"translation": [18.0 + 16.0 * c, 5.0, 95.0]})
        add_instance(s, cube, mat, {"scale": [4.0, 4.0, 4.0], "rotation": [0.0, 0.5 + 0.3 * c, 0.0], "translation": [26.0 + 16.0 * c, 4.0, 120.0]})
    s["materials"].append({"type": "mirror", "albedo": [0.9, 0.9, 0.9]})
    add_instance(s, cube, len(s["materials"]) - 1, {"scale": [30.0, 0.5, 12.0], "translation": [50.0, 0.5, 75.0]})
    return s


def quad(p):  # 4 corners -> inline mesh (two triangles)
    return {"type": "mesh", "vertices": [float(x) for v in p for x in v], "indices": [0, 1, 2, 0, 2, 3]}


def make_c4(n=64, width=3840, height=2160, spp=512):
    X, Y, Z = 400.0, 200.0, 400.0
    s = {"root": 0, "nodes": [{"instances": [0, 1, 2, 3, 4, 5], "children": []}], "instances": [], "transforms": [], "primitives": [],
         "materials": [{"type": "lambertian", "albedo": [0.75, 0.75, 0.75]}, {"type": "lambertian", "albedo": [0.75, 0.25, 0.25]},
                       {"type": "lambertian", "albedo": [0.25, 0.25, 0.75]}],
         "lights": [{"type": "diffuse", "emission": [25.0, 25.0, 25.0]}]}
    walls = [
        ([(0, 0, 0), (X, 0, 0), (X, 0, Z), (0, 0, Z)], 0),      # floor
        ([(0, Y, 0), (0, Y, Z), (X, Y, Z), (X, Y, 0)], 0),      # ceiling
        ([(0, 0, 0), (0, Y, 0), (X, Y, 0), (X, 0, 0)], 0),      # back
        ([(0, 0, Z), (X, 0, Z), (X, Y, Z), (0, Y, Z)], 0),      # front (behind camera)
        ([(0, 0, 0), (0, 0, Z), (0, Y, Z), (0, Y, 0)], 1),      # left
        ([(X, 0, 0), (X, Y, 0), (X, Y, Z), (X, 0, Z)], 2),      # right
    ]
    for p, m in walls:
        s["primitives"].append(quad(p))
        s["instances"].append({"primitive": len(s["primitives"]) - 1, "material": m})
    # light
    s["primitives"].append({"type": "quad"})
    s["instances"].append({"primitive": len(s["primitives"]) - 1, "light": 0})
    s["transforms"].append({"keyframes": [{"scale": [120.0, 1.0, 120.0], "translation": [200.0, 199.5, 200.0]}]})
    s["nodes"].append({"instances": [len(s["instances"]) - 1], "transform": 0})
    s["nodes"][0]["children"].append(1)
    # camera
    s["transforms"].append({"keyframes": [{"rotation": [-0.42, 0.0, 0.0], "translation": [200.0, 150.0, 395.0]}]})
    s["nodes"].append({"transform": 1, "has-camera": True})
    s["nodes"][0]["children"].append(2)
    # bunnies: one primitive, n*n instances
    s["primitives"].append({"type": "mesh", "uri": BUNNY_URI})
    bunny = len(s["primitives"]) - 1
    palette = [[0.75, 0.75, 0.75], [0.8, 0.5, 0.3], [0.4, 0.7, 0.4], [0.5, 0.5, 0.8]]
    first_mat = len(s["materials"])
    for c in palette:
        s["materials"].append({"type": "lambertian", "albedo": c})
    s["materials"].append({"type": "pbr", "albedo": [0.9, 0.8, 0.5], "roughness": 0.3, "metallic": 1.0})
    s["materials"].append({"type": "dielectric", "ior": 1.5})
    nmat = len(s["materials"]) - first_mat
    seed = 12345
    scale = 30.0
    step_x, step_z = (X - 20.0) / n, (Z - 40.0) / n
    for j in range(n):
        for i in range(n):
            seed = (0x00269EC3 + seed * 0x000343FD) & 0xFFFFFFFF      # the reference's LCG (Util/Hash.fs:30-32)
            rot = (seed >> 9) / float(1 << 23) * 6.2831853
            seed = (0x00269EC3 + seed * 0x000343FD) & 0xFFFFFFFF
            mat = first_mat + (seed >> 9) % nmat
            add_instance(s, bunny, mat, {"scale": [scale, scale, scale], "rotation": [0.0, round(rot, 6), 0.0],
                                         "translation": [10.0 + step_x * (i + 0.5), -0.033 * scale, 10.0 + step_z * (j + 0.5)]})
    s["integrator"] = {"type": "path-tracing", "spp": spp, "max-depth": 8}
    s["camera"] = {"type": "pinhole", "fov": 40.0, "aspect-ratio": round(width / height, 4)}
    s["film"] = {"width": width, "height": height, "tone-mapping": "aces"}
    return s


def main():
    write_bunny()
    dump("cbox_pt.json", make_c1())
    dump("cbox_bunny.json", make_c2())
    dump("material_sweep.json", make_c3())
    dump("bunny_instanced.json", make_c4())
    c5 = make_c2()
    # 10 mutations per pixel at 1024x1024 = 10.5 M mutations (BASELINE C5: "10M mutations per GPU");
    # 65 536 chains of 160 mutations: the chain count is the GPU's parallelism (reference default: 1024)
    c5["integrator"] = {"type": "pssmlt", "spp": 10, "max-depth": 8, "n-chains": 65536}
    dump("cbox_mlt.json", c5)
    dump("cbox_ref.json", load_cbox())
    # small variants used by the parity tests (same geometry, tiny films)
    t = make_c4(n=8, width=128, height=72, spp=4)
    dump("bunny_instanced_small.json", t)


if __name__ == "__main__":
    main()
