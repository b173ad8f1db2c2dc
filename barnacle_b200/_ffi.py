"""ctypes mirror of include/barnacle_b200.h (the C-ABI drop-in boundary).

This is the binding a managed host would write with `DllImport`
(INTEGRATION.md shows the F# version).  There is NO fallback: if
`libbarnacle_b200.so` is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BN_LIB") or os.path.join(_HERE, "lib", "libbarnacle_b200.so")  # BN_LIB: experiment builds only

BN_OK = 0
BN_ERR_INVALID, BN_ERR_CUDA, BN_ERR_NO_DEVICE, BN_ERR_IO, BN_ERR_NO_LIGHT = -1, -2, -3, -4, -5
BN_PRIM_MESH, BN_PRIM_SPHERE = 0, 1
BN_MAT_LAMBERTIAN, BN_MAT_MIRROR, BN_MAT_DIELECTRIC, BN_MAT_PBR = 0, 1, 2, 3
BN_CAM_PINHOLE, BN_CAM_THIN_LENS = 0, 1
BN_INTEGRATOR_PATH_TRACING, BN_INTEGRATOR_DIRECT, BN_INTEGRATOR_NORMAL = 0, 1, 2
BN_MLT_GAUSSIAN, BN_MLT_KELEMEN = 0, 1
BN_PARTITION_AUTO, BN_PARTITION_SAMPLE, BN_PARTITION_TILE = 0, 1, 2
BN_RENDER_TRACE_NULL_SHADOW = 1
BN_RENDER_PROFILE = 2
BN_RENDER_FORCE_EXACT = 4


class BnBVHNode(C.Structure):
    _fields_ = [("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3), ("right_or_offset", C.c_int32),
                ("is_leaf", C.c_uint8), ("split_axis", C.c_int8), ("count", C.c_int8), ("visibility_mask", C.c_uint8)]


class BnAliasEntry(C.Structure):
    _fields_ = [("alias", C.c_int32), ("prob", C.c_float), ("pdf", C.c_float)]


class BnInstance(C.Structure):
    _fields_ = [("prim_kind", C.c_uint32), ("prim_id", C.c_uint32), ("material_id", C.c_int32), ("light_id", C.c_int32),
                ("object_to_world", C.c_float * 16), ("world_to_object", C.c_float * 16),
                ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3)]


class BnMesh(C.Structure):
    _fields_ = [("vertex_offset", C.c_uint32), ("vertex_count", C.c_uint32), ("tri_offset", C.c_uint32),
                ("tri_count", C.c_uint32), ("node_offset", C.c_uint32), ("node_count", C.c_uint32),
                ("alias_offset", C.c_uint32), ("reserved", C.c_uint32)]


class BnMaterial(C.Structure):
    _fields_ = [("type", C.c_uint32), ("base_color", C.c_float * 3), ("p0", C.c_float), ("p1", C.c_float)]


class BnLight(C.Structure):
    _fields_ = [("emission", C.c_float * 3), ("two_sided", C.c_uint32)]


class BnCamera(C.Structure):
    _fields_ = [("type", C.c_uint32), ("fov_y", C.c_float), ("aspect_ratio", C.c_float), ("aperture", C.c_float),
                ("focus_distance", C.c_float), ("push_forward", C.c_float), ("camera_to_world", C.c_float * 16)]


class BnSceneDesc(C.Structure):
    _fields_ = [("tlas_nodes", C.POINTER(BnBVHNode)), ("tlas_node_count", C.c_uint32),
                ("instances", C.POINTER(BnInstance)), ("instance_count", C.c_uint32),
                ("light_instances", C.POINTER(C.c_uint32)), ("light_instance_count", C.c_uint32),
                ("meshes", C.POINTER(BnMesh)), ("mesh_count", C.c_uint32),
                ("vertices", C.POINTER(C.c_float)), ("vertex_count", C.c_uint32),
                ("triangles", C.POINTER(C.c_int32)), ("triangle_count", C.c_uint32),
                ("blas_nodes", C.POINTER(BnBVHNode)), ("blas_node_count", C.c_uint32),
                ("alias", C.POINTER(BnAliasEntry)), ("alias_count", C.c_uint32),
                ("sphere_radii", C.POINTER(C.c_float)), ("sphere_count", C.c_uint32),
                ("materials", C.POINTER(BnMaterial)), ("material_count", C.c_uint32),
                ("lights", C.POINTER(BnLight)), ("light_count", C.c_uint32),
                ("camera", BnCamera)]


class BnRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("frame_id", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32), ("flags", C.c_uint32),
                ("interleave_count", C.c_int32), ("interleave_index", C.c_int32), ("integrator", C.c_int32)]


class BnStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_rays_ref", C.c_uint64), ("kernel_launches", C.c_uint64), ("gpu_ms", C.c_double),
                ("extend_ms", C.c_double), ("shade_ms", C.c_double), ("shadow_ms", C.c_double), ("other_ms", C.c_double)]


class BnRay(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3), ("tmax", C.c_float)]


class BnHit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("instance", C.c_int32), ("primitive", C.c_int32)]


class BnHostSceneInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("tone_mapping", C.c_int32), ("integrator", C.c_int32),
                ("spp", C.c_int32), ("max_depth", C.c_int32), ("rr_depth", C.c_int32),
                ("n_bootstrap", C.c_int32), ("n_chains", C.c_int32), ("mutation_strategy", C.c_int32), ("large_step_prob", C.c_float)]


class BnMltParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mutations_per_pixel", C.c_int32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("frame_id", C.c_int32), ("n_bootstrap", C.c_int32), ("n_chains", C.c_int32),
                ("strategy", C.c_int32), ("p0", C.c_float), ("p1", C.c_float), ("large_step_prob", C.c_float),
                ("chain_begin", C.c_int32), ("chain_end", C.c_int32)]


class BnMltStats(C.Structure):
    _fields_ = [("b", C.c_float), ("reserved", C.c_uint32), ("accepted", C.c_uint64), ("proposed", C.c_uint64), ("rays", C.c_uint64),
                ("bootstrap_ms", C.c_double), ("chains_ms", C.c_double)]


assert C.sizeof(BnBVHNode) == 32 and C.sizeof(BnAliasEntry) == 12 and C.sizeof(BnInstance) == 168
assert C.sizeof(BnRay) == 28 and C.sizeof(BnHit) == 20 and C.sizeof(BnMaterial) == 24

# every symbol include/barnacle_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "bn_device_count": (C.c_int, []),
    "bn_last_error": (C.c_char_p, []),
    "bn_release_cached_buffers": (C.c_int, [C.c_int]),
    "bn_measure_l2_read_gbs": (C.c_int, [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "bn_scene_create": (C.c_int, [C.POINTER(BnSceneDesc), C.c_int, C.POINTER(_VP)]),
    "bn_scene_destroy": (None, [_VP]),
    "bn_render": (C.c_int, [_VP, C.POINTER(BnRenderParams), _VP, C.POINTER(BnStats)]),
    "bn_render_device": (C.c_int, [_VP, C.POINTER(BnRenderParams), _VP, _VP, C.POINTER(BnStats)]),
    "bn_multi_scene_create": (C.c_int, [C.POINTER(BnSceneDesc), C.POINTER(C.c_int32), C.c_int32, C.POINTER(_VP)]),
    "bn_multi_scene_destroy": (None, [_VP]),
    "bn_multi_scene_device_count": (C.c_int, [_VP]),
    "bn_render_multi": (C.c_int, [_VP, C.POINTER(BnRenderParams), C.c_int32, _VP, C.POINTER(BnStats)]),
    "bn_multi_partition": (C.c_int, [C.POINTER(BnRenderParams), C.c_int32, C.c_int32, C.c_int32, C.POINTER(BnRenderParams), C.POINTER(C.c_int32)]),
    "bn_trace": (C.c_int, [_VP, _VP, C.c_uint64, C.c_int, _VP]),
    "bn_trace_device": (C.c_int, [_VP, _VP, C.c_uint64, C.c_int, _VP, _VP, C.POINTER(C.c_float)]),
    "bn_render_radiance": (C.c_int, [_VP, C.POINTER(BnRenderParams), _VP]),
    "bn_render_pssmlt": (C.c_int, [_VP, C.POINTER(BnMltParams), _VP, C.POINTER(BnMltStats)]),
    "bn_render_pssmlt_device": (C.c_int, [_VP, C.POINTER(BnMltParams), _VP, _VP, C.POINTER(BnMltStats)]),
    "bn_pssmlt_bootstrap": (C.c_int, [_VP, C.POINTER(BnMltParams), _VP]),
    "bn_debug_render_pssmlt_chains": (C.c_int, [_VP, C.POINTER(BnMltParams), _VP, C.POINTER(BnMltStats), _VP]),
    "bn_debug_order_keys": (C.c_int, [C.c_int, _VP, C.c_uint32, _VP]),
    "bn_film_to_rgba8_device": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP]),
    "bn_bvh_build": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.c_uint32, C.POINTER(BnBVHNode), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]),
    "bn_bvh_build_device": (C.c_int, [C.c_int, _VP, C.c_uint32, _VP, C.c_uint32, _VP, _VP, C.POINTER(C.c_float)]),
    "bn_host_scene_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_float, C.POINTER(_VP)]),
    "bn_host_scene_load_ex": (C.c_int, [C.c_char_p, C.c_char_p, C.c_float, C.c_int, C.POINTER(_VP)]),
    "bn_host_scene_load_string": (C.c_int, [C.c_char_p, C.c_char_p, C.c_float, C.POINTER(_VP)]),
    "bn_host_scene_desc": (C.POINTER(BnSceneDesc), [_VP]),
    "bn_host_scene_info": (None, [_VP, C.POINTER(BnHostSceneInfo)]),
    "bn_host_scene_instance_permutation": (C.POINTER(C.c_uint32), [_VP]),
    "bn_host_scene_triangle_permutation": (C.POINTER(C.c_uint32), [_VP, C.c_uint32]),
    "bn_host_scene_destroy": (None, [_VP]),
    "bn_host_bvh_build": (C.c_int, [C.POINTER(C.c_float), C.c_uint32, C.POINTER(BnBVHNode), C.c_uint32, C.POINTER(C.c_uint32)]),
    "bn_host_alias_build": (C.c_int, [C.POINTER(C.c_float), C.c_uint32, C.POINTER(BnAliasEntry)]),
    "bn_host_film_to_rgba8": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, _VP]),
}

_lib = None


class BarnacleError(RuntimeError):
    """Raised where the reference would `failwith` (non-zero C-ABI status)."""


def load() -> C.CDLL:
    """Load libbarnacle_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BarnacleError(
            f"{LIB_PATH} not found: build it with `python -m barnacle_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != BN_OK:
        msg = load().bn_last_error().decode("utf-8", "replace")
        raise BarnacleError(f"{what or 'barnacle_b200'} failed (status {rc}): {msg}")
