"""Cornell-box variants whose emitters take the branches the modelled scenes never reach: SphereInstance.Sample / EvalPDF
under rotation + non-uniform scale (Extensions/Primitive/Sphere.fs:90-126), one-sided DiffuseLight (Base/Light.fs:49-53,
`two-sided: false`), an emitter that also carries a material, and several emitters at once so that UniformLightSampler's
uSelect remapping matters (Extensions/LightSampler/Uniform.fs:13-29).  Used by the CPU (hostsim / oracle) and the GPU
parity tests; every scene is cbox_pt.json with its emitter list rewritten."""
import copy
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _base():
    with open(os.path.join(ROOT, "scenes", "cbox_pt.json")) as f:
        s = json.load(f)
    s["integrator"] = {"type": "path-tracing", "spp": 4, "max-depth": 8}
    return s


def _add_instance(s, prim, transform, **inst):
    """Appends a primitive + transform + instance + leaf node under the root; returns the instance index."""
    s["primitives"].append(prim)
    s["transforms"].append({"keyframes": [dict(time=0.0, **transform)]})
    s["instances"].append(dict(primitive=len(s["primitives"]) - 1, **inst))
    s["nodes"].append({"instances": [len(s["instances"]) - 1], "transform": len(s["transforms"]) - 1})
    s["nodes"][0]["children"].append(len(s["nodes"]) - 1)
    return len(s["instances"]) - 1


def sphere_emitter():
    """The quad light is gone; the only emitter is a sphere under rotation and non-uniform scale (two-sided, no material)."""
    s = _base()
    s["instances"][6] = {"primitive": 6, "material": 0}          # the former light quad: a plain grey patch on the ceiling
    s["lights"] = [{"type": "diffuse", "emission": [40.0, 36.0, 30.0]}]
    _add_instance(s, {"type": "sphere", "radius": 3.0}, dict(scale=[2.0, 1.0, 1.5], rotation=[0.4, 1.1, -0.7], translation=[50.0, 62.0, 85.0]), light=0)
    return json.dumps(s)


def one_sided():
    """A one-sided quad (normal -Y: it lights the room, not the ceiling) and a one-sided sphere emitter (far-root hits from
    inside do not occur here, near-root normals face outwards: emission only where wo.z > 0)."""
    s = _base()
    s["lights"] = [{"type": "diffuse", "emission": [60.0, 60.0, 60.0], "two-sided": False},
                   {"type": "diffuse", "emission": [10.0, 25.0, 40.0], "two-sided": False}]
    _add_instance(s, {"type": "sphere", "radius": 2.5}, dict(scale=[1.0, 1.6, 1.0], rotation=[0.0, 0.3, 0.9], translation=[78.0, 50.0, 60.0]), light=1)
    return json.dumps(s)


def two_emitters():
    """Three emitters of different colour: the two-sided quad, a one-sided sphere that also has a material (paths continue
    from it), and a small two-sided triangle mesh — uSelect picks among them and is remapped (Uniform.fs:14-17)."""
    s = _base()
    s["lights"] = [{"type": "diffuse", "emission": [45.0, 45.0, 45.0]},
                   {"type": "diffuse", "emission": [30.0, 8.0, 4.0], "two-sided": False},
                   {"type": "diffuse", "emission": [2.0, 20.0, 6.0]}]
    _add_instance(s, {"type": "sphere", "radius": 4.0}, dict(scale=[1.3, 0.7, 1.0], rotation=[0.5, -0.2, 0.3], translation=[25.0, 55.0, 95.0]), light=1, material=2)
    tri = {"type": "mesh", "vertices": [0.0, 0.0, 0.0, 6.0, 0.0, 1.0, 1.0, 5.0, 0.0, 5.0, 4.0, 3.0], "indices": [0, 1, 2, 1, 3, 2]}
    _add_instance(s, tri, dict(scale=[1.5, 1.0, 1.0], rotation=[0.2, 0.5, 0.1], translation=[70.0, 40.0, 40.0]), light=2)
    return json.dumps(s)


SCENES = {"sphere_emitter": sphere_emitter, "one_sided": one_sided, "two_emitters": two_emitters}


def load(name):
    from barnacle_b200.scene import Scene
    return Scene.LoadString(SCENES[name](), base_dir=ROOT)


def variant_of(scene_json: str, **edits) -> str:
    s = copy.deepcopy(json.loads(scene_json))
    s.update(edits)
    return json.dumps(s)
