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
