"""Device BVH builder (bn_bvh_build, "next" row N2) against the host builder
(bn_host_bvh_build, the restatement of Util/BVH.fs:109-247 that tests/test_bvh_build.py pins):
node array and permutation must be identical BYTE FOR BYTE."""
import ctypes as C
import os

import numpy as np
import pytest

from barnacle_b200 import _ffi

pytestmark = pytest.mark.gpu


def host_build(lib, boxes):
    n = boxes.shape[0]
    nodes = (_ffi.BnBVHNode * (2 * n))()
    perm = np.zeros(n, dtype=np.uint32)
    cnt = lib.bn_host_bvh_build(boxes.ctypes.data_as(C.POINTER(C.c_float)), n, nodes, 2 * n, perm.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert cnt > 0, lib.bn_last_error()
    return bytes(memoryview(nodes))[: cnt * 32], perm


def device_build(lib, boxes):
    n = boxes.shape[0]
    nodes = (_ffi.BnBVHNode * (2 * n))()
    perm = np.zeros(n, dtype=np.uint32)
    ms = C.c_float(0)
    cnt = lib.bn_bvh_build(0, boxes.ctypes.data_as(C.POINTER(C.c_float)), n, nodes, 2 * n, perm.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(ms))
    assert cnt > 0, lib.bn_last_error()
    return bytes(memoryview(nodes))[: cnt * 32], perm, ms.value


def random_boxes(n, seed, extent=100.0, size=2.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    c = (rng.random((n, 3)) * extent).astype(np.float32)
    h = (rng.random((n, 3)) * size).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([c - h, c + h], axis=1), dtype=np.float32)


def check_same(lib, boxes):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    hn, hp = host_build(lib, boxes)
    dn, dp, ms = device_build(lib, boxes)
    assert len(hn) == len(dn), (len(hn) // 32, len(dn) // 32)
    assert np.array_equal(hp, dp)
    assert hn == dn
    return len(hn) // 32, ms


@pytest.mark.parametrize("n", [1, 2, 4, 5, 9, 33, 100, 1000, 4099, 50000])
def test_random_boxes_identical_to_host_builder(lib, n):
    check_same(lib, random_boxes(n, seed=100 + n))


def test_clustered_and_flat_boxes(lib):
    rng = np.random.Generator(np.random.PCG64(7))
    # clusters (skewed bins, empty bins -> NaN SAH costs), zero-thickness boxes (cbox-like quads)
    c = np.concatenate([rng.normal(0, 0.01, (3000, 3)), rng.normal(50, 5, (2000, 3)), rng.normal(-20, 0.5, (1000, 3))]).astype(np.float32)
    h = np.abs(rng.normal(0, 0.2, c.shape)).astype(np.float32)
    h[::3, 1] = 0.0
    check_same(lib, np.concatenate([c - h, c + h], axis=1))


def test_identical_centroids_take_the_median_split(lib):
    # splitExtent == 0 (Util/BVH.fs:150-156): no reordering, halves
    boxes = np.tile(np.array([[1, 2, 3, 4, 5, 6]], dtype=np.float32), (37, 1))
    boxes[:, 3:] += np.arange(37, dtype=np.float32)[:, None] * 0  # same box 37 times
    n_nodes, _ = check_same(lib, boxes)
    assert n_nodes > 1
    # same centroid, different sizes
    half = np.linspace(0.5, 3.0, 23, dtype=np.float32)[:, None] * np.ones((1, 3), np.float32)
    check_same(lib, np.concatenate([-half, half], axis=1))


def test_duplicates_and_signed_zeros(lib):
    rng = np.random.Generator(np.random.PCG64(11))
    base = random_boxes(40, seed=5, extent=4.0, size=1.0)
    boxes = base[rng.integers(0, 40, 2000)]                # heavy duplication: equal extremes everywhere
    boxes = np.round(boxes)                                 # many coordinates exactly 0
    z = boxes == 0
    boxes[z & (rng.random(boxes.shape) < 0.5)] = -0.0       # +0 / -0 mixed: MinNative/MaxNative keep the LAST equal operand
    boxes[:, 3:] = np.maximum(boxes[:, 3:], boxes[:, :3])
    check_same(lib, boxes)


def test_axis_aligned_grid(lib):
    # regular grid: exact ties in centroid extents (split-axis tie rule x >= y >= z) and in SAH costs (first minimum wins)
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    check_same(lib, np.concatenate([g, g + 1], axis=1))


def test_bunny_blas_and_instanced_tlas(lib, scene_loader):
    """The builds the C2 / C4 scenes need: 69 451 bunny triangles, 4 000+ instance boxes."""
    scene = scene_loader("cbox_bunny")
    d = scene.desc.contents
    m = max(range(d.mesh_count), key=lambda i: d.meshes[i].tri_count)
    mm = d.meshes[m]
    v = np.ctypeslib.as_array(d.vertices, shape=(d.vertex_count, 3))[mm.vertex_offset:mm.vertex_offset + mm.vertex_count]
    t = np.ctypeslib.as_array(d.triangles, shape=(d.triangle_count, 3))[mm.tri_offset:mm.tri_offset + mm.tri_count]
    inv = np.argsort(scene.triangle_permutation(m))         # back to file order: build from the original triangle order
    p = v[t[inv]]
    boxes = np.concatenate([p.min(1), p.max(1)], axis=1).astype(np.float32)
    n_nodes, ms = check_same(lib, boxes)
    assert n_nodes == mm.node_count                          # and it is the BLAS the scene holds
    print(f"bunny BLAS: {mm.tri_count} triangles -> {n_nodes} nodes in {ms:.2f} ms on the device")
    n_nodes, ms = check_same(lib, random_boxes(4105, seed=3, extent=2000.0, size=40.0))
    print(f"TLAS-sized build: 4105 boxes -> {n_nodes} nodes in {ms:.2f} ms")


def test_rejects_non_finite_boxes(lib):
    boxes = random_boxes(100, seed=1)
    boxes[17, 4] = np.nan
    nodes = (_ffi.BnBVHNode * 200)()
    rc = lib.bn_bvh_build(0, boxes.ctypes.data_as(C.POINTER(C.c_float)), 100, nodes, 200, None, None)
    assert rc == _ffi.BN_ERR_INVALID and b"finite" in lib.bn_last_error()


def test_large_build(lib):
    n_nodes, ms = check_same(lib, random_boxes(1 << 20, seed=99, extent=1000.0, size=1.0))
    print(f"1 Mi boxes -> {n_nodes} nodes in {ms:.1f} ms on the device")


def test_scene_loaded_with_device_builder_is_identical(lib, root):
    """Scene.Load with every BVHNode.Build on the device: the flattened BnSceneDesc (TLAS, BLAS, permuted
    triangles and instances, alias tables) equals the host-built one byte for byte."""
    from barnacle_b200.scene import Scene
    for name in ("cbox_bunny", "bunny_instanced_small", "bunny_instanced"):
        path = os.path.join(root, "scenes", name + ".json")
        a = Scene.Load(path, base_dir=root)
        b = Scene.Load(path, base_dir=root, build_device=0)
        da, db = a.desc.contents, b.desc.contents
        for cnt, ptr, size in (("tlas_node_count", "tlas_nodes", 32), ("blas_node_count", "blas_nodes", 32), ("instance_count", "instances", 168),
                               ("triangle_count", "triangles", 12), ("alias_count", "alias", 12), ("vertex_count", "vertices", 12)):
            na, nb = getattr(da, cnt), getattr(db, cnt)
            assert na == nb, (name, cnt)
            assert C.string_at(getattr(da, ptr), na * size) == C.string_at(getattr(db, ptr), nb * size), (name, ptr)
        assert np.array_equal(a.instance_permutation(), b.instance_permutation())
        a.close(); b.close()
