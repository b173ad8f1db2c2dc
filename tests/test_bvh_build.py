"""Host-side BVH builder (bn_host_bvh_build, a restatement of Util/BVH.fs:128-247) against
(a) structural invariants (SURVEY §8c pin 3) and (b) the independent numpy restatement in
oracle/bvh_build_np.py."""
import ctypes as C
import os

import numpy as np
import pytest

from barnacle_b200 import _ffi
from oracle import bvh_build_np


def c_build(lib, boxes):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    nodes = (_ffi.BnBVHNode * (2 * n))()
    perm = np.zeros(n, dtype=np.uint32)
    cnt = lib.bn_host_bvh_build(boxes.ctypes.data_as(C.POINTER(C.c_float)), n, nodes, 2 * n, perm.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert cnt > 0, lib.bn_last_error()
    return [nodes[i] for i in range(cnt)], perm


def check_invariants(nodes, perm, boxes):
    n = len(perm)
    assert sorted(perm.tolist()) == list(range(n))          # a permutation
    covered = np.zeros(n, dtype=int)

    def rec(i, depth):
        nd = nodes[i]
        lo, hi = np.array(nd.bounds_min[:], np.float32), np.array(nd.bounds_max[:], np.float32)
        if nd.is_leaf:
            f, c = nd.right_or_offset, nd.count
            assert 1 <= c and (c <= 4 or depth >= 64)        # MaxLeafSize 4 / MaxDepth 64
            covered[f:f + c] += 1
            sub = boxes[perm[f:f + c]]
            assert np.array_equal(lo, sub[:, :3].min(0)) and np.array_equal(hi, sub[:, 3:].max(0))
            return lo, hi, depth
        assert 0 <= nd.split_axis <= 2 and nd.right_or_offset > i + 1
        llo, lhi, dl = rec(i + 1, depth + 1)                  # left child = i + 1 (preorder)
        rlo, rhi, dr = rec(nd.right_or_offset, depth + 1)
        assert np.array_equal(lo, np.minimum(llo, rlo)) and np.array_equal(hi, np.maximum(lhi, rhi))
        return lo, hi, max(dl, dr)

    import sys
    sys.setrecursionlimit(10000)
    _, _, depth = rec(0, 0)
    assert (covered == 1).all()
    return depth


def compare_with_numpy(nodes, perm, boxes):
    ref_nodes, ref_perm = bvh_build_np.build(boxes)
    assert np.array_equal(perm, ref_perm)
    assert len(nodes) == len(ref_nodes)
    for a, b in zip(nodes, ref_nodes):
        assert bool(a.is_leaf) == b["leaf"]
        assert np.array_equal(np.array(a.bounds_min[:], np.float32), b["lo"]) and np.array_equal(np.array(a.bounds_max[:], np.float32), b["hi"])
        if b["leaf"]:
            assert (a.right_or_offset, a.count) == (b["first"], b["count"])
        else:
            assert (a.right_or_offset, a.split_axis) == (b["right"], b["axis"])


def random_boxes(n, seed, degenerate=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.random((n, 3), dtype=np.float32) * 10
    if degenerate:
        c[:, 1] = 3.0                      # zero centroid extent on one axis
        c[n // 2:] = c[n // 2]             # many identical centroids -> median splits
    e = rng.random((n, 3), dtype=np.float32) * 0.2
    return np.concatenate([c - e, c + e], axis=1).astype(np.float32)


@pytest.mark.parametrize("n,seed,deg", [(1, 0, False), (4, 1, False), (5, 2, False), (37, 3, False), (1000, 4, False), (300, 5, True)])
def test_random_boxes(lib, n, seed, deg):
    boxes = random_boxes(n, seed, deg)
    nodes, perm = c_build(lib, boxes)
    check_invariants(nodes, perm, boxes)
    compare_with_numpy(nodes, perm, boxes)


def test_identical_boxes_median_split(lib):
    boxes = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (9, 1))
    nodes, perm = c_build(lib, boxes)
    assert perm.tolist() == list(range(9))   # extent == 0: no reordering (Util/BVH.fs:150-156)
    check_invariants(nodes, perm, boxes)


def test_bunny_blas(lib, scene_loader):
    """Stanford bunny through the scene loader: node count / depth of SURVEY App. C and
    agreement with the numpy restatement."""
    scene = scene_loader("cbox_bunny")
    d = scene.desc.contents
    m = [d.meshes[i] for i in range(d.mesh_count) if d.meshes[i].tri_count == 69451][0]
    assert m.vertex_count == 35947 and m.node_count == 40663
    verts = np.ctypeslib.as_array(d.vertices, (d.vertex_count, 3))[m.vertex_offset:m.vertex_offset + m.vertex_count]
    tris = np.ctypeslib.as_array(d.triangles, (d.triangle_count, 3))[m.tri_offset:m.tri_offset + m.tri_count]
    mesh_index = [i for i in range(d.mesh_count) if d.meshes[i].tri_count == 69451][0]
    perm = scene.triangle_permutation(mesh_index)
    # undo the permutation to get the OBJ order, rebuild with both builders
    orig = np.empty_like(tris)
    orig[perm] = tris
    p = verts[orig]                                   # [n,3,3]
    boxes = np.concatenate([p.min(axis=1), p.max(axis=1)], axis=1).astype(np.float32)
    nodes, perm2 = c_build(lib, boxes)
    assert np.array_equal(perm, perm2) and len(nodes) == 40663
    depth = check_invariants(nodes, perm2, boxes)
    assert depth == 18                                # root = depth 0 ("max depth 18")
    leaves = [nd.count for nd in nodes if nd.is_leaf]
    assert len(leaves) == 20332
    compare_with_numpy(nodes, perm2, boxes)
    blas = [d.blas_nodes[m.node_offset + i] for i in range(m.node_count)]
    assert all(bytes(a) == bytes(b) for a, b in zip(blas, nodes))


def test_alias_table_quirk(lib):
    """AliasTable never creates aliases: alias = i, prob = 1, pdf = w/sum."""
    w = np.array([1, 2, 3, 10], np.float32)
    out = (_ffi.BnAliasEntry * 4)()
    assert lib.bn_host_alias_build(w.ctypes.data_as(C.POINTER(C.c_float)), 4, out) == 0
    for i in range(4):
        assert out[i].alias == i and out[i].prob == 1.0 and out[i].pdf == np.float32(w[i]) / np.float32(16)
    z = np.zeros(3, np.float32)
    out = (_ffi.BnAliasEntry * 3)()
    lib.bn_host_alias_build(z.ctypes.data_as(C.POINTER(C.c_float)), 3, out)
    assert [o.pdf for o in out] == [np.float32(1) / np.float32(3)] * 3
