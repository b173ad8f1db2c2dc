"""Function-level pin of the world-space interaction the oracle hands to shading: hit point,
OrthonormalBasis(n) (Primitive.fs:15-23), Triangle.Intersect's unflipped normal (Mesh.fs:76), SpherePrimitive's normal
with its near / far root asymmetry (Sphere.fs:47-70) and LocalGeometry / OrthonormalBasis.Transform with the
un-renormalised n' = t' x b' (Primitive.fs:34-38,57-58) — against a float64 restatement on the randomised
scenes of test_oracle_bruteforce.py (rotations, non-uniform scales, nested transforms)."""
import numpy as np
import pytest

from barnacle_b200.scene import RAY_DTYPE, Scene
from conftest import random_rays
from oracle.oracle_ffi import OracleScene
from test_oracle_bruteforce import _random_scene_json


def _unit(v):
    return v / np.linalg.norm(v)


def _onb(n):  # Primitive.fs:15-23
    axis = np.array([0.0, 1.0, 0.0]) if abs(n[0]) > 0.1 else np.array([1.0, 0.0, 0.0])
    t = _unit(np.cross(n, axis))
    return t, np.cross(n, t)


def ref_interaction(desc, hit, ray):
    inst = desc.instances[int(hit["instance"])]
    o2w = np.array(inst.object_to_world[:], dtype=np.float64).reshape(4, 4)
    w2o = np.array(inst.world_to_object[:], dtype=np.float64).reshape(4, 4)
    o = ray["origin"].astype(np.float64) @ w2o[:3, :3] + w2o[3, :3]        # Ray.Transform (Ray.fs:19-22)
    d = ray["direction"].astype(np.float64) @ w2o[:3, :3]
    p = o + float(hit["t"]) * d
    if inst.prim_kind == 0:
        m = desc.meshes[inst.prim_id]
        k = int(hit["primitive"])

        def vertex(j):
            v = 3 * (m.vertex_offset + desc.triangles[3 * (m.tri_offset + k) + j])
            return np.array([desc.vertices[v], desc.vertices[v + 1], desc.vertices[v + 2]], dtype=np.float64)
        p0, p1, p2 = vertex(0), vertex(1), vertex(2)
        n = _unit(np.cross(p1 - p0, p2 - p0))                                # never flipped towards the ray (Mesh.fs:76)
    else:
        n = _unit(p)
        radius = float(desc.sphere_radii[inst.prim_id])
        a, bq, c = d @ d, -(o @ d), o @ o - radius * radius                  # Sphere.fs:38-47: the near root t0 = c / q is
        disc = radius * radius - (o + bq / a * d) @ (o + bq / a * d)          # taken if it lies beyond eps, else the far one
        q = bq + np.copysign(np.sqrt(max(a * disc, 0.0)), bq)
        near_root = c / q > 1e-3
        if near_root and n @ d > 0:                                          # near root only (Sphere.fs:56-57 vs :66-70)
            n = -n
    t, b = _onb(n)
    tw, bw = _unit(t @ o2w[:3, :3]), _unit(b @ o2w[:3, :3])                  # OrthonormalBasis.Transform (Primitive.fs:34-38)
    return p @ o2w[:3, :3] + o2w[3, :3], np.cross(tw, bw), tw, bw


@pytest.mark.parametrize("seed,n_instances", [(11, 4), (12, 25), (13, 90)])
def test_interaction_frame_matches_restatement(lib, seed, n_instances):
    rng = np.random.default_rng(seed)
    scene = Scene.LoadString(_random_scene_json(rng, n_instances))
    desc = scene.desc.contents
    rays = random_rays(scene, 600, seed=seed)
    pick = rng.integers(0, desc.instance_count, size=len(rays))
    lo = np.array([desc.instances[int(k)].bounds_min[:] for k in pick], dtype=np.float64)
    hi = np.array([desc.instances[int(k)].bounds_max[:] for k in pick], dtype=np.float64)
    d = lo + (hi - lo) * rng.random((len(rays), 3)) - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    oracle = OracleScene(scene.desc)
    hits = oracle.trace(rays)
    kinds = {0: 0, 1: 0}
    far_roots = skew = 0
    for ray, hit in zip(rays, hits):
        if hit["instance"] < 0:
            continue
        one = np.zeros(1, dtype=RAY_DTYPE)
        one[0] = ray
        ok, g = oracle.closest_geom(one)
        assert ok
        p, n, t, b = ref_interaction(desc, hit, ray)
        scale = max(1.0, float(np.abs(p).max()))
        np.testing.assert_allclose(g[0], p, atol=2e-4 * scale)
        np.testing.assert_allclose(g[2], t, atol=2e-4)
        np.testing.assert_allclose(g[3], b, atol=2e-4)
        np.testing.assert_allclose(g[1], n, atol=4e-4)
        inst = desc.instances[int(hit["instance"])]
        kinds[inst.prim_kind] += 1
        far_roots += inst.prim_kind == 1 and bool(g[1] @ ray["direction"] > 0)   # seen from inside: the outward normal is kept (Q6)
        skew += abs(np.linalg.norm(g[1]) - 1) > 1e-3                        # Q7: n' is not a unit vector under non-uniform scale
    assert kinds[0] > 50
    if seed == 12:
        assert kinds[1] > 20 and skew > 0 and far_roots > 0
