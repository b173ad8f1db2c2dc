"""ctypes loader for the ORACLE (oracle/libbarnacle_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BN_NET9_FMA=0 in the environment selects the restatement built without .NET 9's fused Cross / Transform (Makefile)
_FMA0 = os.environ.get("BN_NET9_FMA") == "0"
LIB_NAME = "libbarnacle_oracle_fma0.so" if _FMA0 else "libbarnacle_oracle.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("instance", "<i4"), ("primitive", "<i4")])
COUNTER_NAMES = ["tlas_nodes", "blas_nodes", "tris_fetched", "tris_box_pass", "inst_visited", "inst_box_pass", "inst_committed", "rays"]

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "barnacle_oracle.cpp")
    deps = [src, os.path.join(_HERE, "..", "include", "barnacle_b200.h"), os.path.join(_HERE, "..", "include", "bn_portable_math.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", LIB_NAME], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.bo_xxhash32_two.restype = C.c_uint32
        lib.bo_xxhash32_two.argtypes = [C.c_uint32, C.c_uint32]
        lib.bo_xxhash32_three.restype = C.c_uint32
        lib.bo_xxhash32_three.argtypes = [C.c_uint32] * 3
        lib.bo_lcg.restype = C.c_float
        lib.bo_lcg.argtypes = [C.POINTER(C.c_uint32)]
        lib.bo_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.bo_atan.restype = C.c_float
        lib.bo_atan.argtypes = [C.c_float]
        lib.bo_scene_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.bo_scene_destroy.argtypes = [C.c_void_p]
        lib.bo_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.bo_closest_geom.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bo_primary_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bo_render_radiance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.bo_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.bo_set_portable_math.argtypes = [C.c_int]
        lib.bo_mlt_sampler_script.argtypes = [C.c_uint32, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.bo_camera_ray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.bo_light_eval_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bo_light_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        lib.bo_material_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bo_material_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        lib.bo_pssmlt_bootstrap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.bo_render_pssmlt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


class OracleScene:
    """CPU restatement of BVHAggregate + UniformLightSampler + PathTracingIntegrator
    over a BnSceneDesc (any object exposing `.desc`, a ctypes pointer to it)."""

    def __init__(self, desc_ptr):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.bo_scene_create(C.cast(desc_ptr, C.c_void_p), C.byref(h))
        if rc != 0:
            raise RuntimeError("bo_scene_create failed")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bo_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def trace(self, rays: np.ndarray, any_hit: bool = False, counters: bool = False, threads: int = 0):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        cnt = np.zeros(8, dtype=np.uint64) if counters else None
        rc = self._lib.bo_trace(self._h, rays.ctypes.data, rays.shape[0], 1 if any_hit else 0, hits.ctypes.data,
                                cnt.ctypes.data if counters else None, threads)
        assert rc == 0
        return (hits, dict(zip(COUNTER_NAMES, (int(x) for x in cnt)))) if counters else hits

    def closest_geom(self, ray: np.ndarray):
        ray = np.ascontiguousarray(ray, dtype=RAY_DTYPE)
        out = np.zeros(12, dtype=np.float32)
        hit = self._lib.bo_closest_geom(self._h, ray.ctypes.data, out.ctypes.data)
        return bool(hit), out.reshape(4, 3)

    def camera_ray(self, width: int, height: int, x: int, y: int, u_pixel, u_lens) -> np.ndarray:
        """CameraBase.GeneratePrimaryRay(resolution, (x, y), uPixel, uLens) as a one-element ray batch."""
        u = np.array([u_pixel[0], u_pixel[1], u_lens[0], u_lens[1]], dtype=np.float32)
        out = np.zeros(1, dtype=RAY_DTYPE)
        self._lib.bo_camera_ray(self._h, width, height, x, y, u.ctypes.data, out.ctypes.data)
        return out

    def light_eval_hit(self, ray: np.ndarray):
        """UniformLightSampler.Eval(ray.Origin, interaction) at the closest hit: (L xyz, pdf), or None if not an emitter."""
        ray = np.ascontiguousarray(ray, dtype=RAY_DTYPE)
        out = np.zeros(4, dtype=np.float32)
        return out if self._lib.bo_light_eval_hit(self._h, ray.ctypes.data, out.ctypes.data) else None

    def light_sample(self, p, usel: float, ulight) -> np.ndarray:
        """UniformLightSampler.Sample(p, uSelect, uLight): (eval.p xyz, eval.L xyz, eval.pdf, wi xyz)."""
        p, ul, out = np.asarray(p, np.float32), np.asarray(ulight, np.float32), np.zeros(10, np.float32)
        self._lib.bo_light_sample(self._h, p.ctypes.data, float(usel), ul.ctypes.data, out.ctypes.data)
        return out

    def primary_rays(self, params) -> np.ndarray:
        n = (params.sample_end - params.sample_begin) * (params.y1 - params.y0) * (params.x1 - params.x0)
        out = np.empty(n, dtype=RAY_DTYPE)
        self._lib.bo_primary_rays(self._h, C.byref(params), out.ctypes.data)
        return out

    def render_radiance(self, params, threads: int = 0) -> np.ndarray:
        ns = params.sample_end - params.sample_begin
        out = np.empty((ns, params.y1 - params.y0, params.x1 - params.x0, 3), dtype=np.float32)
        rc = self._lib.bo_render_radiance(self._h, C.byref(params), out.ctypes.data, threads)
        assert rc == 0, rc
        return out

    def pssmlt_bootstrap(self, params, threads: int = 0) -> np.ndarray:
        w = np.empty(params.n_bootstrap, dtype=np.float32)
        rc = self._lib.bo_pssmlt_bootstrap(self._h, C.byref(params), w.ctypes.data, None, threads)
        assert rc == 0, rc
        return w

    def render_pssmlt(self, params, threads: int = 0):
        """Returns (film [H*W,3], stats dict, accepted-per-chain)."""
        film = np.zeros((params.height * params.width, 3), dtype=np.float32)
        out = np.zeros(5, dtype=np.uint64)
        per_chain = np.zeros(max(params.chain_end - params.chain_begin, 1), dtype=np.uint32)
        rc = self._lib.bo_render_pssmlt(self._h, C.byref(params), film.ctypes.data, out.ctypes.data, per_chain.ctypes.data, threads)
        assert rc == 0, rc
        b = np.array([int(out[0])], dtype=np.uint32).view(np.float32)[0]
        return film, {"B": float(b), "B_bits": int(out[0]), "accepted": int(out[1]), "proposed": int(out[2]), "rays": int(out[3]),
                      "chain_seconds": int(out[4]) * 1e-6}, per_chain

    def render(self, params, threads: int = 0, counters: bool = False):
        """Returns (film [H*W,3], stats dict)."""
        film = np.empty((params.height * params.width, 3), dtype=np.float32)
        st = np.zeros(5, dtype=np.uint64)
        ce = np.zeros(8, dtype=np.uint64) if counters else None
        cs = np.zeros(8, dtype=np.uint64) if counters else None
        rc = self._lib.bo_render(self._h, C.byref(params), film.ctypes.data, st.ctypes.data,
                                 ce.ctypes.data if counters else None, cs.ctypes.data if counters else None, threads)
        assert rc == 0, rc
        stats = {"paths": int(st[0]), "extend_rays": int(st[1]), "shadow_rays": int(st[2]), "shadow_rays_nonnull": int(st[3]),
                 "seconds": int(st[4]) * 1e-6}
        if counters:
            stats["extend_counters"] = dict(zip(COUNTER_NAMES, (int(x) for x in ce)))
            stats["shadow_counters"] = dict(zip(COUNTER_NAMES, (int(x) for x in cs)))
        return film, stats


def material_eval(material, wo, wi) -> np.ndarray:
    """(bsdf.xyz, pdf) of MaterialBase.Eval for a BnMaterial (ctypes struct of the product binding)."""
    wo, wi, out = np.asarray(wo, np.float32), np.asarray(wi, np.float32), np.zeros(4, np.float32)
    load().bo_material_eval(C.addressof(material), wo.ctypes.data, wi.ctypes.data, out.ctypes.data)
    return out


def material_sample(material, wo, ulobe: float, u) -> np.ndarray:
    """(bsdf.xyz, pdf, wi.xyz) of MaterialBase.Sample."""
    wo, u, out = np.asarray(wo, np.float32), np.asarray(u, np.float32), np.zeros(7, np.float32)
    load().bo_material_sample(C.addressof(material), wo.ctypes.data, float(ulobe), u.ctypes.data, out.ctypes.data)
    return out


def mlt_sampler_script(seed_state: int, large_step_prob: float, strategy: int, p0: float, p1: float, script, n_dims: int) -> np.ndarray:
    """Runs MLTSampler (PSSMLT.fs:35-150) through a script of ops (0 StartIteration, 1 Next1D, 2 Accept, 3 Reject); returns the draws."""
    script = np.ascontiguousarray(script, dtype=np.int32)
    out = np.zeros(int((script == 1).sum()), dtype=np.float32)
    k = load().bo_mlt_sampler_script(seed_state, large_step_prob, strategy, p0, p1, script.ctypes.data, len(script), n_dims, out.ctypes.data)
    assert k == len(out), k
    return out


def film_to_rgba8(film: np.ndarray, width: int, height: int, tone: int) -> np.ndarray:
    """Film.PostProcess + Rgba32 (Base/Film.fs:21-30,55-66): W*H*3 float32 -> [H, W, 4] uint8."""
    film = np.ascontiguousarray(film, dtype=np.float32)
    out = np.empty((height, width, 4), dtype=np.uint8)
    lib = load()
    lib.bo_film_to_rgba8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    assert lib.bo_film_to_rgba8(film.ctypes.data, width, height, tone, out.ctypes.data) == 0
    return out


def set_portable_math(on: bool) -> None:
    load().bo_set_portable_math(1 if on else 0)


def num_threads() -> int:
    return load().bo_num_threads()
