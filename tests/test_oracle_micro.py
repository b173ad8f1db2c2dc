"""Pins for the oracle (SURVEY §8c pin 2): hand-checkable micro-scenes with analytically
known answers, and the reference quirks the restatement must reproduce.
CPU only."""
import json

import numpy as np
import pytest

from barnacle_b200.scene import RAY_DTYPE, Scene, make_params
from oracle.oracle_ffi import OracleScene


def scene_json(primitives, instances, nodes, transforms=(), materials=(), lights=(), camera=None, film=(8, 8)):
    return json.dumps({
        "nodes": nodes, "instances": instances, "transforms": list(transforms), "primitives": primitives,
        "materials": list(materials), "lights": list(lights),
        "integrator": {"type": "path-tracing", "spp": 1}, "camera": camera or {"type": "pinhole", "fov": 90.0},
        "film": {"width": film[0], "height": film[1], "tone-mapping": "identity"}})


def rays(*items):
    r = np.zeros(len(items), dtype=RAY_DTYPE)
    for i, (o, d, t) in enumerate(items):
        r[i] = (o, d, t)
    return r


TRI = {"type": "mesh", "vertices": [0, 0, 0, 1, 0, 0, 0, 1, 0], "indices": [0, 1, 2]}
CAM_NODE = {"has-camera": True}


def load(text):
    s = Scene.LoadString(text)
    return s, OracleScene(s.desc)


def test_single_triangle_t_u_v(lib):
    s, o = load(scene_json([TRI], [{"primitive": 0, "material": 0}], [{"instances": [0], "children": [1]}, CAM_NODE], materials=[{"type": "lambertian"}]))
    h = o.trace(rays(((0.25, 0.5, 1.0), (0, 0, -1), np.inf), ((0.25, 0.5, 3.0), (0, 0, -2), np.inf),
                     ((0.7, 0.7, 1.0), (0, 0, -1), np.inf), ((0.25, 0.5, 1.0), (0, 0, -1), 0.5), ((0.25, 0.5, -1.0), (0, 0, -1), np.inf)))
    assert h["instance"].tolist() == [0, 0, -1, -1, -1] and h["primitive"][:2].tolist() == [0, 0]
    assert h["t"][0] == 1.0 and h["u"][0] == 0.25 and h["v"][0] == 0.5       # u along P1-P0, v along P2-P0 (Mesh.fs:62-78)
    assert h["t"][1] == 1.5                                                  # direction is not renormalised
    assert h["t"][3] == 0.5                                                  # miss leaves t at tmax
    a = o.trace(rays(((0.25, 0.5, 1.0), (0, 0, -1), 1.5), ((0.25, 0.5, 1.0), (0, 0, -1), 1.0)), any_hit=True)
    assert a["instance"].tolist() == [1, 0]                                  # strict t' < t (Mesh.fs:48)
    hit, g = o.closest_geom(rays(((0.25, 0.5, 1.0), (0, 0, -1), np.inf)))
    assert hit and np.allclose(g[0], [0.25, 0.5, 0.0]) and np.allclose(g[1], [0, 0, 1])   # n = normalize(e0 x e1)


def test_instance_transform_shares_t(lib):
    tr = [{"keyframes": [{"scale": [2, 2, 2], "translation": [10, 0, 0]}]}]
    s, o = load(scene_json([TRI], [{"primitive": 0, "material": 0}], [{"children": [1, 2]}, {"instances": [0], "transform": 0}, CAM_NODE],
                           transforms=tr, materials=[{"type": "lambertian"}]))
    h = o.trace(rays(((10.5, 1.0, 4.0), (0, 0, -1), np.inf)))
    assert h["instance"][0] == 0 and h["t"][0] == 4.0 and h["u"][0] == 0.25 and h["v"][0] == 0.5
    hit, g = o.closest_geom(rays(((10.5, 1.0, 4.0), (0, 0, -1), np.inf)))
    assert np.allclose(g[0], [10.5, 1.0, 0.0], atol=1e-6)


def test_sphere_roots_and_normal_quirk(lib):
    """Near root flips the normal against the ray; the far root (origin inside) does NOT."""
    tr = [{"keyframes": [{"translation": [0, 0, -5]}]}]
    s, o = load(scene_json([{"type": "sphere", "radius": 2.0}], [{"primitive": 0, "material": 0}],
                           [{"children": [1, 2]}, {"instances": [0], "transform": 0}, CAM_NODE], transforms=tr, materials=[{"type": "lambertian"}]))
    h = o.trace(rays(((0, 0, 0), (0, 0, -1), np.inf), ((0, 0, -4.5), (0, 0, -1), np.inf), ((0, 0, 0), (0, 1, 0), np.inf)))
    assert h["instance"].tolist() == [0, 0, -1] and h["primitive"].tolist()[:2] == [0, 0]
    assert abs(h["t"][0] - 3.0) < 1e-6 and abs(h["t"][1] - 2.5) < 1e-6
    _, g_out = o.closest_geom(rays(((0, 0, 0), (0, 0, -1), np.inf)))
    _, g_in = o.closest_geom(rays(((0, 0, -4.5), (0, 0, -1), np.inf)))
    assert np.allclose(g_out[1], [0, 0, 1], atol=1e-6)       # faces the ray
    assert np.allclose(g_in[1], [0, 0, -1], atol=1e-6)       # outward normal kept: points along the ray
    # eps = 1e-3: a ray starting on the surface does not re-hit it at t ~ 0
    h2 = o.trace(rays(((0, 0, -3), (0, 0, 1), np.inf), ((0, 0, -3), (0, 0, -1), np.inf)))
    assert h2["instance"].tolist() == [-1, 0] and abs(h2["t"][1] - 4.0) < 1e-5


def test_flat_box_self_hit_rejection_and_nan_slab(lib):
    """no origin offset; the 1e-3 tMin of the per-triangle AABB rejects the self hit on an
    axis-aligned (flat-box) triangle.  A.2: 0*inf = NaN in the second slab operand fails the test."""
    s, o = load(scene_json([TRI], [{"primitive": 0, "material": 0}], [{"instances": [0], "children": [1]}, CAM_NODE], materials=[{"type": "lambertian"}]))
    h = o.trace(rays(((0.25, 0.25, 0.0), (0, 0, 1), np.inf), ((0.25, 0.25, 0.0), (0.1, 0, 1), np.inf),
                     ((-1.0, 0.25, 0.0), (1, 0, 0), np.inf)))        # in-plane ray: z lane is (0-0)*inf = NaN
    assert h["instance"].tolist() == [-1, -1, -1]


def test_tie_break_first_visited_wins(lib):
    """Two coincident triangles: strict `t' < t` keeps the first one visited (slot order in the leaf)."""
    prim = {"type": "mesh", "vertices": [0, 0, 0, 1, 0, 0, 0, 1, 0], "indices": [0, 1, 2, 0, 1, 2]}
    s, o = load(scene_json([prim], [{"primitive": 0, "material": 0}], [{"instances": [0], "children": [1]}, CAM_NODE], materials=[{"type": "lambertian"}]))
    h = o.trace(rays(((0.25, 0.25, 1.0), (0, 0, -1), np.inf)))
    assert h["primitive"][0] == 0


def test_emitter_seen_directly_is_exact(lib):
    """depth 0, MIS weight 1 (PathTracing.fs:33-35): every pixel that sees the two-sided quad light
    carries exactly its emission; the light has no material so the path ends there."""
    tr = [{"keyframes": [{"scale": [50, 1, 50], "rotation": [1.5707963267948966, 0, 0], "translation": [0, 0, -5]}]}]
    s, o = load(scene_json([{"type": "quad"}], [{"primitive": 0, "light": 0}], [{"children": [1, 2]}, {"instances": [0], "transform": 0}, CAM_NODE],
                           transforms=tr, lights=[{"type": "diffuse", "emission": [3.0, 2.0, 1.0]}]))
    film, st = o.render(make_params(8, 8, 4))
    assert np.array_equal(film, np.tile(np.array([3, 2, 1], np.float32), (64, 1)))
    assert st["extend_rays"] == 8 * 8 * 4 and st["shadow_rays"] == 0


def test_direct_lighting_estimate_converges(lib):
    """A Lambertian floor under a small two-sided quad light.  max-depth 2 = NEE at the floor plus
    BSDF-sampled emitter hits, i.e. both halves of the balance-heuristic MIS (PathTracing.fs:33-59);
    nothing else is in the scene, so the sum is the direct lighting rho/pi * L * A * cos*cos/d^2
    (point-light approximation of the 1x1 light at d = 4: good to ~1%)."""
    tr = [{"keyframes": [{"scale": [0.5, 1, 0.5], "translation": [0, 4, -6]}]},            # light 1x1 at height 4 above floor point (0,0,-6)
          {"keyframes": [{"scale": [100, 1, 100], "translation": [0, 0, 0]}]},
          {"keyframes": [{"translation": [0, 2, 0]}]}]
    cam = {"type": "pinhole", "fov": 1.0}
    text = scene_json([{"type": "quad"}, {"type": "quad"}], [{"primitive": 0, "light": 0}, {"primitive": 1, "material": 0}],
                      [{"children": [1, 2, 3]}, {"instances": [0], "transform": 0}, {"instances": [1], "transform": 1}, {"transform": 2, "has-camera": True}],
                      transforms=tr, materials=[{"type": "lambertian", "albedo": [0.5, 0.5, 0.5]}], lights=[{"type": "diffuse", "emission": [10, 10, 10]}], camera=cam)
    # camera at (0,2,0) looks down -Z: it sees the floor only if tilted; instead aim rays with the trace API and use render for the estimate:
    s = Scene.LoadString(text)
    o = OracleScene(s.desc)
    # tilt-free check through Li is awkward with a fixed -Z camera, so place the camera looking at the floor via rotation
    sc = json.loads(text)
    sc["transforms"][2] = {"keyframes": [{"rotation": [-0.32175055, 0, 0], "translation": [0, 2, 0]}]}   # atan(2/6): looks at (0,0,-6)
    s2 = Scene.LoadString(json.dumps(sc))
    film, _ = OracleScene(s2.desc).render(make_params(4, 4, 4096, max_depth=2))
    got = film.reshape(4, 4, 3)[1:3, 1:3].mean()
    expect = 0.5 / np.pi * 10.0 * 1.0 * 1.0 / 16.0       # rho/pi * L * A * cos(0)*cos(0) / d^2, d = 4
    assert abs(got / expect - 1) < 0.03, (got, expect)


def test_normal_and_direct_integrators_on_micro_scene(lib):
    """NormalIntegrator: 0.5*(n+1) (Normal.fs:15); DirectIntegrator on a bare emitter: its emission (Direct.fs:16-17)."""
    from barnacle_b200 import _ffi
    tr = [{"keyframes": [{"scale": [50, 1, 50], "rotation": [1.5707963267948966, 0, 0], "translation": [0, 0, -5]}]}]
    s, o = load(scene_json([{"type": "quad"}], [{"primitive": 0, "light": 0}], [{"children": [1, 2]}, {"instances": [0], "transform": 0}, CAM_NODE],
                           transforms=tr, lights=[{"type": "diffuse", "emission": [3.0, 2.0, 1.0]}]))
    film, st = o.render(make_params(8, 8, 2, integrator=_ffi.BN_INTEGRATOR_DIRECT))
    assert np.array_equal(film, np.tile(np.array([3, 2, 1], np.float32), (64, 1))) and st["shadow_rays"] == 0
    film, st = o.render(make_params(8, 8, 2, integrator=_ffi.BN_INTEGRATOR_NORMAL))
    hit, g = o.closest_geom(rays(((0, 0, 0), (0, 0, -1), np.inf)))
    assert hit and np.allclose(np.abs(g[1]), [0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(film, np.tile(0.5 * (g[1] + 1), (64, 1)), atol=1e-6)
    assert st["extend_rays"] == 8 * 8 * 2 and st["shadow_rays"] == 0


def test_loader_errors_match_reference_messages(lib):
    from barnacle_b200._ffi import BarnacleError
    bad = json.loads(scene_json([{"type": "torus"}], [], [{}]))
    with pytest.raises(BarnacleError, match="Unknown primitive type: torus"):   # Loader.fs:116
        Scene.LoadString(json.dumps(bad))
    bad = json.loads(scene_json([TRI], [{"primitive": 0, "material": 0}], [{"instances": [0]}], materials=[{"type": "velvet"}]))
    with pytest.raises(BarnacleError, match="Unknown material type: velvet"):   # Loader.fs:73
        Scene.LoadString(json.dumps(bad))


def test_cbox_defaults_and_tlas_order(lib, scene_loader):
    """Loader defaults (Loader.fs:181-183) and the TLAS permuting the instance array."""
    s = scene_loader("cbox_pt")
    assert (s.info.spp, s.info.max_depth, s.info.rr_depth) == (64, 8, 5) and s.integrator_type == "path-tracing"
    d = s.desc.contents
    perm = s.instance_permutation()
    assert sorted(perm.tolist()) == list(range(9))
    assert d.light_instance_count == 1 and perm[d.light_instances[0]] == 6      # instance 6 is the light quad
    inst = d.instances[d.light_instances[0]]
    assert inst.material_id == -1 and inst.light_id == 0
    assert np.allclose(inst.object_to_world[12:15], [50.0, 81.5, 80.0])        # row-vector convention: translation in M41..M43
    assert np.allclose(np.array(inst.bounds_min[:]), [40, 81.5, 70]) and np.allclose(np.array(inst.bounds_max[:]), [60, 81.5, 90])
