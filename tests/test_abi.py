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


# ---- the F# binding delivered as source (integration/fsharp/GpuIntegrators.fs) ----------------------------
# It cannot be compiled here (no dotnet), so its struct mirrors are pinned field by field against the ctypes
# mirror above, which test_struct_sizes / the GPU tests pin against the header and the library.
_FS_TYPES = {"uint32": (4, 4), "int": (4, 4), "float32": (4, 4), "float": (8, 8), "uint64": (8, 8), "nativeint": (8, 8),
             "Vector3": (12, 4), "Matrix4x4": (64, 4)}


def _fsharp_structs(root):
    text = open(os.path.join(root, "integration", "fsharp", "GpuIntegrators.fs")).read()
    out = {}
    for m in re.finditer(r"type (Bn\w+) = // (\d+)[^\n]*\n((?:\s+val mutable \w+: \w+[^\n]*\n)+)", text):
        fields = re.findall(r"val mutable (\w+): (\w+)", m.group(3))
        out[m.group(1)] = (int(m.group(2)), fields)
    return out


def _sequential_layout(fields, known):
    """.NET LayoutKind.Sequential == the C rule: natural alignment per field, size rounded to the largest."""
    off, align, offsets = 0, 1, []
    for _, ty in fields:
        size, a = known[ty]
        off = (off + a - 1) // a * a
        offsets.append(off)
        off += size
        align = max(align, a)
    return offsets, (off + align - 1) // align * align, align


def test_fsharp_binding_structs_match_the_c_abi(root):
    structs = _fsharp_structs(root)
    mirror = {"BnInstance": _ffi.BnInstance, "BnMesh": _ffi.BnMesh, "BnMaterial": _ffi.BnMaterial, "BnLight": _ffi.BnLight,
              "BnCamera": _ffi.BnCamera, "BnSceneDesc": _ffi.BnSceneDesc, "BnRenderParams": _ffi.BnRenderParams,
              "BnStats": _ffi.BnStats, "BnMltParams": _ffi.BnMltParams, "BnMltStats": _ffi.BnMltStats}
    assert sorted(structs) == sorted(mirror)
    known = dict(_FS_TYPES)
    for name in ["BnInstance", "BnMesh", "BnMaterial", "BnLight", "BnCamera", "BnSceneDesc", "BnRenderParams", "BnStats",
                 "BnMltParams", "BnMltStats"]:
        quoted, fields = structs[name]
        offsets, size, align = _sequential_layout(fields, known)
        known[name] = (size, align)
        ct = mirror[name]
        assert size == quoted == ctypes.sizeof(ct), name
        assert len(fields) == len(ct._fields_), name
        for (fs_name, _), off, (c_name, _) in zip(fields, offsets, ct._fields_):
            assert off == getattr(ct, c_name).offset, f"{name}.{fs_name} vs {c_name}"
            # same field, modulo naming convention (camelCase vs snake_case; `kind` is `type`, a keyword in F#)
            assert fs_name.lower() == c_name.replace("_", "") or (fs_name, c_name) == ("kind", "type"), f"{name}.{fs_name} vs {c_name}"


def test_fsharp_binding_imports_only_declared_symbols(root):
    text = open(os.path.join(root, "integration", "fsharp", "GpuIntegrators.fs")).read()
    imported = set(re.findall(r"extern \w+ (bn_\w+)\(", text))
    assert imported and imported <= set(_ffi.SYMBOLS)
    assert {"bn_scene_create", "bn_scene_destroy", "bn_render", "bn_render_pssmlt", "bn_last_error"} <= imported


# ---- the C++ caller that mirrors Program.fs (barnacle_b200/csrc/cli/barnacle_gpu.cpp) -----------------------------
def test_cli_mirrors_program_fs_argument_and_error_behaviour(lib, root, tmp_path):
    import subprocess
    exe = os.path.join(root, "barnacle_b200", "lib", "barnacle_gpu")
    assert os.path.exists(exe), "the build step links the CLI next to the library"
    run = lambda *a: subprocess.run([exe, *a], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    r = run("-o", "x.ppm")                                            # Program.fs:24: failwith "Invalid arguments."
    assert r.returncode != 0 and "Invalid arguments." in r.stderr
    r = run("-i", "scenes/does_not_exist.json", "-o", str(tmp_path / "x.ppm"))
    assert r.returncode != 0 and "cannot open" in r.stderr
    if lib.bn_device_count() == 0:                                    # no CPU fallback: loud failure after a successful load
        for scene in ("scenes/cbox_pt.json", "scenes/cbox_mlt.json"):
            r = run("-i", scene, "-o", str(tmp_path / "x.ppm"), "--base-dir", root)
            assert r.returncode != 0 and "Loaded scene from" in r.stdout and "no CUDA device" in r.stderr
            assert not (tmp_path / "x.ppm").exists()
