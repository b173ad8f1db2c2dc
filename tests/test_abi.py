"""The C-ABI library loads and exports every symbol include/barnacle_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from barnacle_b200 import _ffi


def header_symbols(root):
    text = open(os.path.join(root, "include", "barnacle_b200.h")).read()
    return sorted(set(re.findall(r"BN_API\s+[\w\s\*]+?\b(bn_\w+)\s*\(", text)))


def test_header_and_binding_agree(root):
    assert header_symbols(root) == sorted(_ffi.SYMBOLS)


def test_library_exports_every_symbol(lib, root):
    raw = ctypes.CDLL(_ffi.LIB_PATH)
    for name in header_symbols(root):
        assert hasattr(raw, name), f"{name} not exported"


def test_struct_sizes():
    assert ctypes.sizeof(_ffi.BnBVHNode) == 32       # Util/BVH.fs:52 (Size = 32)
    assert ctypes.sizeof(_ffi.BnAliasEntry) == 12    # Util/AliasTable.fs:7-12
    assert ctypes.sizeof(_ffi.BnInstance) == 168
    assert ctypes.sizeof(_ffi.BnRay) == 28
    assert ctypes.sizeof(_ffi.BnHit) == 20
    assert _ffi.BnBVHNode.right_or_offset.offset == 24 and _ffi.BnBVHNode.is_leaf.offset == 28
    assert _ffi.BnBVHNode.split_axis.offset == 29 and _ffi.BnBVHNode.count.offset == 30


def test_no_device_is_an_error_not_a_fallback(lib, scene_loader):
    """Without a GPU the product path must fail loudly (BN_ERR_NO_DEVICE), never fall back."""
    if lib.bn_device_count() > 0:
        return
    import pytest
    scene = scene_loader("cbox_pt")
    with pytest.raises(_ffi.BarnacleError, match="no CUDA device"):
        scene.gpu()


def test_last_error_and_bad_args(lib):
    h = ctypes.c_void_p()
    rc = lib.bn_host_scene_load(b"/nonexistent/scene.json", None, 0.0, ctypes.byref(h))
    assert rc == _ffi.BN_ERR_IO and b"cannot open" in lib.bn_last_error()
    rc = lib.bn_host_scene_load_string(b'{"nodes": [', None, 0.0, ctypes.byref(h))
    assert rc == _ffi.BN_ERR_IO and b"JSON" in lib.bn_last_error()


def test_device_entry_points_fail_loudly_without_a_gpu(lib):
    """bn_bvh_build / bn_measure_l2_read_gbs have no host fallback either."""
    if lib.bn_device_count() > 0:
        return
    import numpy as np
    boxes = np.array([[0, 0, 0, 1, 1, 1]] * 8, dtype=np.float32)
    nodes = (_ffi.BnBVHNode * 16)()
    rc = lib.bn_bvh_build(0, boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 8, nodes, 16, None, None)
    assert rc == _ffi.BN_ERR_NO_DEVICE and b"no CUDA device" in lib.bn_last_error()
    gbs = ctypes.c_double(0)
    assert lib.bn_measure_l2_read_gbs(0, 32 << 20, 4, ctypes.byref(gbs)) == _ffi.BN_ERR_NO_DEVICE


def test_scene_load_ex_with_host_builder_equals_scene_load(lib, root):
    """bn_host_scene_load_ex(build_device = -1) is bn_host_scene_load."""
    import os
    from barnacle_b200.scene import Scene
    path = os.path.join(root, "scenes", "cbox_pt.json")
    a = Scene.Load(path, base_dir=root)
    b = Scene.Load(path, base_dir=root, build_device=-1)
    da, db = a.desc.contents, b.desc.contents
    assert da.tlas_node_count == db.tlas_node_count and da.instance_count == db.instance_count
    assert ctypes.string_at(da.tlas_nodes, da.tlas_node_count * 32) == ctypes.string_at(db.tlas_nodes, db.tlas_node_count * 32)
    assert ctypes.string_at(da.instances, da.instance_count * 168) == ctypes.string_at(db.instances, db.instance_count * 168)
    b.close()
