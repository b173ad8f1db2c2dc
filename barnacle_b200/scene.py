"""Host-side mirror of the reference interface for the hot path.

Names follow the reference so call sites read like Barnacle's own:

    scene = Scene.Load("scenes/cbox_pt.json")       # Extensions/Scene/Loader.fs:277-281
    scene.Render(0.0, "out.png")                    # Extensions/Scene/Render.fs:10-19

`Scene.Render` does what Render.fs does — Film.Clear, Scene.Traverse, BVHAggregate,
UniformLightSampler (all inside the host-side builder), then the timed
`Integrator.Render`, which here is `GpuPathTracingIntegrator.Render`: one C-ABI
call into the CUDA wavefront (bn_render).  No PyTorch is needed on this path;
torch is only used by the multi-GPU driver (barnacle_b200/multi_gpu.py).
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass

import numpy as np

from . import _ffi
from ._ffi import BarnacleError, BnHostSceneInfo, BnMltParams, BnMltStats, BnRenderParams, BnStats, check

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("instance", "<i4"), ("primitive", "<i4")])
assert RAY_DTYPE.itemsize == 28 and HIT_DTYPE.itemsize == 20

TONE_MAPPING = {0: "identity", 1: "aces", 2: "gamma"}
INTEGRATORS = {0: "normal", 1: "direct", 2: "path-tracing", 3: "pssmlt"}


@dataclass
class RenderStats:
    paths: int
    extend_rays: int
    shadow_rays: int
    shadow_rays_ref: int
    kernel_launches: int
    gpu_ms: float
    extend_ms: float = 0.0
    shade_ms: float = 0.0
    shadow_ms: float = 0.0
    other_ms: float = 0.0

    @property
    def rays(self) -> int:
        return self.extend_rays + self.shadow_rays


def make_params(width, height, spp, max_depth=8, rr_depth=5, frame_id=0, sample_begin=0, sample_end=None,
                rect=None, flags=0, interleave=(1, 0), integrator=0) -> BnRenderParams:
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, width, height)
    return BnRenderParams(width, height, spp, max_depth, rr_depth, frame_id, sample_begin,
                          spp if sample_end is None else sample_end, x0, y0, x1, y1, flags, interleave[0], interleave[1], integrator)


def make_mlt_params(width, height, mutations_per_pixel, max_depth=8, rr_depth=5, frame_id=0, n_bootstrap=4 * 1024 * 1024, n_chains=1024,
                    strategy="Gaussian", large_step_prob=0.5, chain_begin=0, chain_end=None) -> BnMltParams:
    """PSSMLTIntegrator's parameters with the reference's defaults (Loader.fs:189-203)."""
    if strategy == "Gaussian":
        st, p0, p1 = _ffi.BN_MLT_GAUSSIAN, 1e-2, 0.0
    elif strategy == "Kelemen":
        st, p0, p1 = _ffi.BN_MLT_KELEMEN, 1.0 / 1024.0, 1.0 / 16.0
    else:
        raise BarnacleError(f"Unknown mutation strategy: {strategy}")
    return BnMltParams(width, height, mutations_per_pixel, max_depth, rr_depth, frame_id, n_bootstrap, n_chains, st, p0, p1, large_step_prob,
                       chain_begin, n_chains if chain_end is None else chain_end)


class Film:
    """Base/Film.fs: Pixels is Vector3[W*H], row 0 = top (SetPixel flips Y)."""

    def __init__(self, width: int, height: int, tone_mapping: str = "identity"):
        self.ImageWidth, self.ImageHeight, self.ToneMapping = width, height, tone_mapping
        self.Pixels = np.zeros((height * width, 3), dtype=np.float32)

    @property
    def Resolution(self):
        return (self.ImageWidth, self.ImageHeight)

    def Clear(self):
        self.Pixels[:] = 0

    def image(self) -> np.ndarray:
        return self.Pixels.reshape(self.ImageHeight, self.ImageWidth, 3)

    def to_rgba8(self) -> np.ndarray:
        lib = _ffi.load()
        out = np.empty((self.ImageHeight, self.ImageWidth, 4), dtype=np.uint8)
        tm = {v: k for k, v in TONE_MAPPING.items()}[self.ToneMapping]
        check(lib.bn_host_film_to_rgba8(self.Pixels.ctypes.data, self.ImageWidth, self.ImageHeight, tm, out.ctypes.data), "Film.Save")
        return out

    def Save(self, filename: str):
        """Film.Save (Film.fs:55-66): tone-map + clamp -> RGBA8 -> encoder chosen by extension."""
        if filename.endswith(".npy"):
            np.save(filename, self.image())
            return
        from PIL import Image
        Image.fromarray(self.to_rgba8(), "RGBA").save(filename)


class GpuScene:
    """Device-resident scene (bn_scene_create) — the aggregate + light sampler + camera
    the integrator reads, flattened."""

    def __init__(self, desc_ptr, device: int = 0):
        self._lib = _ffi.load()
        h = C.c_void_p()
        check(self._lib.bn_scene_create(desc_ptr, device, C.byref(h)), "bn_scene_create")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bn_scene_destroy(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def _stats(st: BnStats) -> RenderStats:
        return RenderStats(st.paths, st.extend_rays, st.shadow_rays, st.shadow_rays_ref, st.kernel_launches, st.gpu_ms,
                           st.extend_ms, st.shade_ms, st.shadow_ms, st.other_ms)

    def render(self, params: BnRenderParams, film: np.ndarray | None = None):
        """bn_render: host film (W*H*3 fp32, Film.Pixels layout); returns (film, stats)."""
        if film is None:
            film = np.empty((params.height * params.width, 3), dtype=np.float32)
        assert film.dtype == np.float32 and film.size == params.width * params.height * 3 and film.flags.c_contiguous
        st = BnStats()
        check(self._lib.bn_render(self._h, C.byref(params), film.ctypes.data, C.byref(st)), "bn_render")
        return film, self._stats(st)

    def render_device(self, params: BnRenderParams, d_film_ptr: int, stream: int = 0):
        """bn_render_device: film is a device pointer on this scene's device."""
        st = BnStats()
        check(self._lib.bn_render_device(self._h, C.byref(params), C.c_void_p(d_film_ptr), C.c_void_p(stream), C.byref(st)), "bn_render_device")
        return self._stats(st)

    def render_radiance(self, params: BnRenderParams) -> np.ndarray:
        ns = params.sample_end - params.sample_begin
        out = np.empty((ns, params.y1 - params.y0, params.x1 - params.x0, 3), dtype=np.float32)
        check(self._lib.bn_render_radiance(self._h, C.byref(params), out.ctypes.data), "bn_render_radiance")
        return out

    def render_pssmlt(self, params: BnMltParams, film: np.ndarray | None = None):
        """bn_render_pssmlt: accumulates INTO `film` (zeros if None); returns (film, BnMltStats)."""
        if film is None:
            film = np.zeros((params.height * params.width, 3), dtype=np.float32)
        st = BnMltStats()
        check(self._lib.bn_render_pssmlt(self._h, C.byref(params), film.ctypes.data, C.byref(st)), "bn_render_pssmlt")
        return film, st

    def render_pssmlt_device(self, params: BnMltParams, d_film_ptr: int, stream: int = 0) -> BnMltStats:
        st = BnMltStats()
        check(self._lib.bn_render_pssmlt_device(self._h, C.byref(params), C.c_void_p(d_film_ptr), C.c_void_p(stream), C.byref(st)), "bn_render_pssmlt_device")
        return st

    def pssmlt_bootstrap(self, params: BnMltParams) -> np.ndarray:
        w = np.empty(params.n_bootstrap, dtype=np.float32)
        check(self._lib.bn_pssmlt_bootstrap(self._h, C.byref(params), w.ctypes.data), "bn_pssmlt_bootstrap")
        return w

    def trace(self, rays: np.ndarray, any_hit: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        check(self._lib.bn_trace(self._h, rays.ctypes.data, rays.shape[0], 1 if any_hit else 0, hits.ctypes.data), "bn_trace")
        return hits

    def trace_device(self, d_rays: int, n: int, any_hit: bool, d_hits: int, stream: int = 0) -> float:
        ms = C.c_float()
        check(self._lib.bn_trace_device(self._h, C.c_void_p(d_rays), n, 1 if any_hit else 0, C.c_void_p(d_hits), C.c_void_p(stream), C.byref(ms)), "bn_trace_device")
        return ms.value


class MultiGpuScene:
    """The scene on several devices behind ONE C-ABI handle (bn_multi_scene_create): flattened once, uploaded to every device
    by its own host thread.  `render` is bn_render_multi — the drop-in for Integrator.Render on a multi-GPU box: the
    (pixel, sampleId) space is shared out inside the call, the films are combined over NVLink on devices[0]."""

    def __init__(self, desc_ptr, devices):
        self._lib = _ffi.load()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(self._lib.bn_multi_scene_create(desc_ptr, arr, len(self.devices), C.byref(h)), "bn_multi_scene_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bn_multi_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def render(self, params: BnRenderParams, film: np.ndarray | None = None, partition: int = _ffi.BN_PARTITION_AUTO):
        if film is None:
            film = np.empty((params.height * params.width, 3), dtype=np.float32)
        assert film.dtype == np.float32 and film.flags["C_CONTIGUOUS"] and film.size == params.height * params.width * 3
        st = BnStats()
        check(self._lib.bn_render_multi(self._h, C.byref(params), partition, film.ctypes.data, C.byref(st)), "bn_render_multi")
        return film, GpuScene._stats(st)


def multi_partition(params: BnRenderParams, partition: int, n_devices: int, rank: int):
    """(share of device `rank`, empty?) as bn_render_multi computes it (host logic; needs no GPU)."""
    out, empty = BnRenderParams(), C.c_int32(0)
    check(_ffi.load().bn_multi_partition(C.byref(params), partition, n_devices, rank, C.byref(out), C.byref(empty)), "bn_multi_partition")
    return out, bool(empty.value)


class GpuPathTracingIntegrator:
    """Drop-in for PathTracingIntegrator (PathTracing.fs:9-12) whose Render
    (Integrator.fs:46-55) runs on the GPU.  `kind` selects the Li variant:
    "path-tracing" | "direct" (Direct.fs) | "normal" (Normal.fs)."""

    KINDS = {"path-tracing": _ffi.BN_INTEGRATOR_PATH_TRACING, "direct": _ffi.BN_INTEGRATOR_DIRECT, "normal": _ffi.BN_INTEGRATOR_NORMAL}

    def __init__(self, spp: int, max_depth: int = 8, rr_depth: int = 5, kind: str = "path-tracing"):
        self.SamplePerPixel, self.MaxDepth, self.RRDepth = spp, max_depth, rr_depth
        self.kind = kind
        self.FrameId = 0
        self.last_stats: RenderStats | None = None

    def Render(self, gpu_scene: GpuScene, film: Film, flags: int = 0):
        p = make_params(film.ImageWidth, film.ImageHeight, self.SamplePerPixel, self.MaxDepth, self.RRDepth, self.FrameId, flags=flags,
                        integrator=self.KINDS[self.kind])
        _, self.last_stats = gpu_scene.render(p, film.Pixels)
        self.FrameId += 1


class Scene:
    """Base/Scene.fs `Scene` + Loader.fs `Scene.Load` + Render.fs `Scene.Render`."""

    def __init__(self, handle, lib, time_: float = 0.0, reload=None):
        self._h, self._lib = handle, lib
        self._time, self._reload = float(time_), reload   # pose of the flattened scene; how to re-run Scene.Traverse at another t
        info = BnHostSceneInfo()
        lib.bn_host_scene_info(handle, C.byref(info))
        self.info = info
        self.Film = Film(info.width, info.height, TONE_MAPPING[info.tone_mapping])
        self.integrator_type = INTEGRATORS[info.integrator]
        kind = self.integrator_type if self.integrator_type in GpuPathTracingIntegrator.KINDS else "path-tracing"
        self.Integrator = GpuPathTracingIntegrator(info.spp, info.max_depth, info.rr_depth, kind)
        self._gpu: GpuScene | None = None

    @staticmethod
    def Load(filename: str, base_dir: str | None = None, time_: float = 0.0, build_device: int | None = None) -> "Scene":
        """build_device: CUDA ordinal whose BVH builder (bn_bvh_build) builds every BLAS and the TLAS; None = host builder.
        The flattened scene is identical either way."""
        lib = _ffi.load()
        bd = (base_dir or os.getcwd()).encode()

        def load_at(t: float):
            h = C.c_void_p()
            if build_device is None:
                check(lib.bn_host_scene_load(filename.encode(), bd, C.c_float(t), C.byref(h)), "Scene.Load")
            else:
                check(lib.bn_host_scene_load_ex(filename.encode(), bd, C.c_float(t), int(build_device), C.byref(h)), "Scene.Load")
            return h
        return Scene(load_at(time_), lib, time_, load_at)

    @staticmethod
    def LoadString(text: str, base_dir: str | None = None, time_: float = 0.0) -> "Scene":
        lib = _ffi.load()
        bd = (base_dir or os.getcwd()).encode()

        def load_at(t: float):
            h = C.c_void_p()
            check(lib.bn_host_scene_load_string(text.encode(), bd, C.c_float(t), C.byref(h)), "Scene.Load")
            return h
        return Scene(load_at(time_), lib, time_, load_at)

    @property
    def desc(self):
        return self._lib.bn_host_scene_desc(self._h)

    def instance_permutation(self) -> np.ndarray:
        n = self.desc.contents.instance_count
        return np.ctypeslib.as_array(self._lib.bn_host_scene_instance_permutation(self._h), (n,)).copy()

    def triangle_permutation(self, mesh: int) -> np.ndarray:
        n = self.desc.contents.meshes[mesh].tri_count
        return np.ctypeslib.as_array(self._lib.bn_host_scene_triangle_permutation(self._h, mesh), (n,)).copy()

    def gpu(self, device: int = 0) -> GpuScene:
        if self._gpu is None or self._gpu.device != device:
            self._gpu = GpuScene(self.desc, device)
        return self._gpu

    def Render(self, t: float, filename: str | None, device: int = 0) -> float:
        """Render.fs:10-19.  Returns the seconds spent in Integrator.Render (the
        reference's Stopwatch region)."""
        self.Film.Clear()
        self.Traverse(t)
        gpu = self.gpu(device)
        t0 = time.perf_counter()
        if self.integrator_type == "pssmlt":
            i = self.info
            p = make_mlt_params(i.width, i.height, i.spp, i.max_depth, i.rr_depth, 0, i.n_bootstrap, i.n_chains,
                                "Kelemen" if i.mutation_strategy == _ffi.BN_MLT_KELEMEN else "Gaussian", i.large_step_prob)
            _, st = gpu.render_pssmlt(p, self.Film.Pixels)
            if st.b == 0.0:
                print("Warning: all bootstrap samples are zero, exiting...")
            else:
                print(f"Accepted mutation count: {st.accepted}\nProposed mutation count: {st.proposed}\nAcceptance rate: {st.accepted / max(st.proposed, 1):f}")
        else:
            self.Integrator.Render(gpu, self.Film)
        dt = time.perf_counter() - t0
        print(f"Render time: {dt:f} seconds")
        if filename:
            self.Film.Save(filename)
        return dt

    def Traverse(self, t: float):
        """Scene.Traverse(t) (Base/Scene.fs:38-61, called by every Render, Render.fs:10-13): instance transforms at time t,
        then the TLAS and the light list over them.  The flattened scene holds ONE pose, so another t re-runs the host-side
        builder (keyframe evaluation, instance bounds, BVHNode.Build) and drops the device copy of the old pose."""
        if float(t) == self._time:
            return
        if self._reload is None:
            raise BarnacleError(f"Scene.Render at t = {t}: this scene was flattened at t = {self._time} and cannot be re-traversed")
        h = self._reload(float(t))
        if self._gpu is not None:
            self._gpu.close()
            self._gpu = None
        self._lib.bn_host_scene_destroy(self._h)
        self._h, self._time = h, float(t)

    def close(self):
        if self._gpu is not None:
            self._gpu.close()
            self._gpu = None
        if self._h:
            self._lib.bn_host_scene_destroy(self._h)
            self._h = None

    __del__ = close
