#!/usr/bin/env python3
"""bench.py — the hot path's headline measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one pass of the hot path over one batch: one full frame of the named
workload (default C2 = BASELINE.json configs[1]: Cornell box + Stanford bunny,
1024x1024, 256 spp) — raygen -> extend -> shade -> shadow -> accumulate for every
(pixel, sampleId).  The scene is synthetic in the sense of the contract: authored
in the reference's schema by scenes/make_scenes.py, seeds fully determined by
(x, y, sampleId).

metric  Mrays/s = (closest-hit + any-hit rays actually traced) / device time.
value   whole-job rate with the scene already resident in HBM and the film left in
        HBM (bn_render_device), CUDA events, max over ranks.
e2e     the same metric through the reference-facing C-ABI with HOST buffers
        (bn_scene_create from host arrays + bn_render into a host film): host->device
        copy of the flattened scene and device->host read of the film inside the
        timed region.
N > 1   launched by torchrun, one rank per GPU; the fixed frame is split by sample
        index (strong scaling), one NCCL sum-reduce of the fp32 film to rank 0.
--impl reference   the reference's CPU path (the C++ restatement in oracle/; the .NET
        binary cannot run in this image) on all host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene file, width, height, spp, label)
    "C1": ("cbox_pt.json", 512, 512, 64, "C1 cbox path-tracing 512x512x64spp"),
    "C2": ("cbox_bunny.json", 1024, 1024, 256, "C2 cbox+stanford-bunny 1024x1024x256spp"),
    "C3": ("material_sweep.json", 1920, 1080, 1024, "C3 GGX+dielectric sweep 1920x1080x1024spp"),
    "C4": ("bunny_instanced.json", 3840, 2160, 512, "C4 4096 bunny instances 3840x2160x512spp"),
    # "next" row N1: PSSMLT, 10 mutations per pixel = 10.5 M mutations; chains split over the GPUs
    "C5": ("cbox_mlt.json", 1024, 1024, 10, "C5 PSSMLT cbox+bunny 1024x1024, 10.5M mutations, 65536 chains"),
}
MAX_DEPTH, RR_DEPTH = 8, 5
# algorithmic bytes per traced ray in the REFERENCE layout
B_NODE, B_TRI, B_INST, B_IO_EXTEND, B_IO_SHADOW = 32, 48, 152, 48, 32


def algorithmic_bytes_per_ray(c: dict, io: int, cap: bool = False) -> float:
    """32 B per node popped + 48 B per leaf triangle fetched + the instance
    record + ray/hit I/O, all in the REFERENCE layout and visiting order.  The instance
    term is the exact split the oracle counts (24 B bounds per instance visited, + 64 B
    WorldToObject when its AABB passes, + 64 B ObjectToWorld on a committed hit);
    cap=True charges the full 152 B per visit instead (the survey's upper bound)."""
    rays = max(c["rays"], 1)
    inst = B_INST * c["inst_visited"] if cap else 24 * c["inst_visited"] + 64 * c["inst_box_pass"] + 64 * c["inst_committed"]
    return (B_NODE * (c["tlas_nodes"] + c["blas_nodes"]) + B_TRI * c["tris_fetched"] + inst) / rays + io


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for (_, r) in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def host_threads() -> int:
    """The host threads this process may use.  Passed to the oracle EXPLICITLY: torchrun exports OMP_NUM_THREADS=1 to
    every rank, and a CPU arm that follows it runs on one core (round 1's scaling record did exactly that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def oracle_sample(scene, w, h, spp_sample, counters: bool, libm: bool = True, keep_film: bool = False):
    """The CPU restatement on a bounded sample of the workload (all host threads)."""
    from barnacle_b200.scene import make_params
    from oracle.oracle_ffi import OracleScene, set_portable_math
    set_portable_math(not libm)
    threads = host_threads()
    try:
        o = OracleScene(scene.desc)
        p = make_params(w, h, spp_sample, MAX_DEPTH, RR_DEPTH)
        film, st = o.render(p, counters=counters, threads=threads)
    finally:
        set_portable_math(True)
    st["threads"] = threads
    if keep_film:
        st["film"] = film
    return st


def rel_mse(image, reference, eps=1e-2):
    """mean over pixels and channels of (I - R)^2 / (R^2 + eps); pixels that are not finite in either image are left out
    and counted (PSSMLT-style NaN pixels do not occur in path tracing, but a metric must say so rather than hide it)."""
    import numpy as np
    a, b = np.asarray(image, dtype=np.float64).reshape(-1, 3), np.asarray(reference, dtype=np.float64).reshape(-1, 3)
    ok = np.isfinite(a).all(axis=1) & np.isfinite(b).all(axis=1)
    if not ok.any():
        return None, int((~ok).sum())
    return float((((a[ok] - b[ok]) ** 2) / (b[ok] ** 2 + eps)).mean()), int((~ok).sum())


def run_reference(args, scene_file, W, H, SPP, label):
    """--impl reference: the reference's CPU renderer (C++ restatement, libm math, 16x16
    tiles scheduled dynamically over all host threads), each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from barnacle_b200.scene import Scene
    scene = Scene.Load(os.path.join(ROOT, "scenes", scene_file), base_dir=ROOT)
    # bounded sample per step, sized from one calibration pass so that the K timed steps take about a minute whatever the
    # host (rates are spp-independent); --ref-spp pins it
    t_cal = oracle_sample(scene, W, H, 1, False)["seconds"]
    spp_s = args.ref_spp if args.ref_spp > 0 else max(1, min(SPP, int(args.ref_seconds / max(args.steps, 1) / max(t_cal, 1e-3))))
    for _ in range(max(0, args.warmup - 1)):
        oracle_sample(scene, W, H, 1, False)
    rays = secs = paths = 0
    threads = 0
    for _ in range(args.steps):
        st = oracle_sample(scene, W, H, spp_s, False)
        rays += st["extend_rays"] + st["shadow_rays"]
        paths += st["paths"]
        secs += st["seconds"]
        threads = st["threads"]
    val = rays / secs / 1e6
    sample = f"{W}x{H} at {spp_s} of {SPP} spp per step (rates are spp-independent); rays = extend + shadow rays the reference traces"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": label, "max_depth": MAX_DEPTH, "rr_depth": RR_DEPTH},
        "samples_per_s": paths / secs,
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of Barnacle's CPU path (the .NET binary is not runnable in this image)"},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_pssmlt(args, scene_file, W, H, SPP, label):
    """C5: one step = one PSSMLTIntegrator.Render (bootstrap + all chains + film reduce)."""
    import torch
    import torch.distributed as dist
    from barnacle_b200.multi_gpu import render_pssmlt_sharded
    from barnacle_b200.scene import Scene, make_mlt_params
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle.oracle_ffi import OracleScene, set_portable_math
        scene = Scene.Load(os.path.join(ROOT, "scenes", scene_file), base_dir=ROOT)
        i = scene.info
        set_portable_math(False)
        o = OracleScene(scene.desc)
        p = make_mlt_params(W, H, 1, i.max_depth, i.rr_depth, 0, 262144, 1024)   # bounded sample: 1 mutation/pixel, 256 Ki bootstrap
        rays = secs = muts = 0
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _, st, _ = o.render_pssmlt(p, threads=host_threads())
            if k >= args.warmup:
                secs += time.perf_counter() - t0; rays += st["rays"]; muts += st["proposed"]
        set_portable_math(True)
        val = rays / secs / 1e6
        print(json.dumps({"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": label}, "mutations_per_s": muts / secs,
                          "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": host_threads(), "kind": "port",
                                           "sample": "1 of 10 mutations per pixel, 262144 bootstrap paths, 1024 chains per step"},
                          "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = Scene.Load(os.path.join(ROOT, "scenes", scene_file), base_dir=ROOT)
    i = scene.info
    gpu = scene.gpu(local)
    film = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
    base = make_mlt_params(W, H, SPP, i.max_depth, i.rr_depth, 0, i.n_bootstrap, i.n_chains, "Gaussian", i.large_step_prob)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        film.zero_()
        return render_pssmlt_sharded(gpu, base, film, dist if world > 1 else None, stream)

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = muts = acc = 0
    ms = 0.0
    barrier()
    for _ in range(args.steps):
        ev0.record(); st = step(); ev1.record(); torch.cuda.synchronize()
        ms += ev0.elapsed_time(ev1); rays += st.rays; muts += st.proposed; acc += st.accepted
    barrier()
    t_wall1 = time.time()
    host_film = torch.empty(W * H * 3, dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        if rank == 0:
            host_film.copy_(film)
    barrier()
    e2e_s = time.perf_counter() - t0
    # stopped only now (an exiting NVML client can stall the next CUDA calls); its report covers t_wall0..t_wall1, the timed region
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    # the same 10.5 M mutations over four times the chains (N = 1 only; outside the timed regions): from ~10^5 chains up the device
    # integrator runs the chains as a wavefront through the path tracer's traversal kernels instead of one chain per thread
    more = None
    if world == 1 and rank == 0 and not args.no_configs:
        big = make_mlt_params(W, H, SPP, i.max_depth, i.rr_depth, 0, i.n_bootstrap, 4 * i.n_chains, "Gaussian", i.large_step_prob)
        film.zero_()
        render_pssmlt_sharded(gpu, big, film, None, stream)
        ev0.record(); stb = render_pssmlt_sharded(gpu, big, film, None, stream); ev1.record(); torch.cuda.synchronize()
        mb = ev0.elapsed_time(ev1)
        more = {"n_chains": 4 * i.n_chains, "value": stb.rays / mb / 1e3, "unit": "Mrays/s", "ms_per_step": mb, "mutations_per_s": stb.proposed / (mb * 1e-3),
                "acceptance_rate": stb.accepted / max(stb.proposed, 1), "bootstrap_ms": stb.bootstrap_ms, "chains_ms": stb.chains_ms,
                "form": "bootstrap and chains as waves through the wavefront's traversal kernels (CUDA-graph replay per mutation round)"}
    agg = torch.tensor([ms, float(rays), float(muts), float(acc), e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        agg[0], agg[4] = mx[0], mx[4]
    if rank == 0:
        ms, rays, muts, acc, e2e_s = (float(x) for x in agg)
        print(json.dumps({
            "metric": "Mrays/s", "value": rays / ms / 1e3, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "n_chains": i.n_chains, "n_bootstrap": i.n_bootstrap, "parallelism": f"chain-split x{world} + film reduce" if world > 1 else "single GPU",
                       "l2_flush": "film (12.6 MB) rewritten each step; working set (primary samples + scene) is L2-resident by design"},
            "mutations_per_s": muts / (ms * 1e-3), "acceptance_rate": acc / max(muts, 1.0), "gpu_launches": 2 * args.steps * world,
            "e2e": {"value": rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": W * H * 12, "ms_per_step": e2e_s / args.steps * 1e3},
            "clocks": clk, "more_chains": more,
            "roofline": {"bound": "hbm", "kernel": "k_mlt_chains (one Markov chain per thread)", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "per-thread megakernel on the exact per-lane traversal (divergence- and latency-bound, scene in L2); no bytes-per-mutation roofline is claimed for this row"},
            "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (the result is then NOT the named config)")
    ap.add_argument("--ref-spp", type=int, default=0, help="--impl reference: spp of the bounded sample per step (0: sized so that the timed steps take about a minute)")
    ap.add_argument("--ref-seconds", type=float, default=60.0, help="--impl reference: CPU seconds the K timed steps should take together")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary `configs` lines (C1 / C3 / C4 at stated reduced spp)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    scene_file, W, H, SPP, label = WORKLOADS[args.workload]
    if args.spp:
        SPP = args.spp
        label += f" [spp overridden to {SPP}]"
    if args.workload == "C5":
        return run_pssmlt(args, scene_file, W, H, SPP, label)
    if args.impl == "reference":
        return run_reference(args, scene_file, W, H, SPP, label)

    import numpy as np
    import torch
    import torch.distributed as dist
    from barnacle_b200 import _ffi
    from barnacle_b200.multi_gpu import partition, render_sharded, shard_params
    from barnacle_b200.scene import GpuScene, Scene, make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream
    lib = _ffi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(gpu, base, film, steps, warmup, mode):
        """W untimed + K timed steps of one workload on this rank's shard; returns (max-over-ranks ms of the K steps,
        job-wide totals, this rank's per-class ms, wall-clock bracket of the timed region)."""
        def step():
            flush.zero_()  # L2 flush between iterations
            return render_sharded(gpu, base, film, dist if world > 1 else None, stream, mode)
        for _ in range(warmup):
            step()
        barrier()
        t_w0 = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = {"extend": 0, "shadow": 0, "paths": 0, "launches": 0, "extend_ms": 0.0, "shade_ms": 0.0, "shadow_ms": 0.0, "other_ms": 0.0, "render_ms": 0.0}
        my_ms = 0.0
        barrier()
        for _ in range(steps):
            ev0.record()
            st = step()
            ev1.record()
            torch.cuda.synchronize()
            my_ms += ev0.elapsed_time(ev1)   # flush.zero_() is inside: ~0.1 ms of a multi-hundred-ms step, stated in config
            if st is not None:
                tot["extend"] += st.extend_rays; tot["shadow"] += st.shadow_rays; tot["paths"] += st.paths; tot["launches"] += st.kernel_launches
                tot["extend_ms"] += st.extend_ms; tot["shade_ms"] += st.shade_ms; tot["shadow_ms"] += st.shadow_ms; tot["other_ms"] += st.other_ms
                tot["render_ms"] += st.gpu_ms
        barrier()
        t_w1 = time.time()
        agg = torch.tensor([my_ms, tot["extend"], tot["shadow"], tot["paths"], tot["launches"], tot["extend_ms"], tot["shadow_ms"]], dtype=torch.float64, device="cuda")
        total_ms = my_ms
        if world > 1:
            mx = agg.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            total_ms = float(mx[0])
        job = {"extend": float(agg[1]), "shadow": float(agg[2]), "paths": float(agg[3]), "launches": int(agg[4]),
               "extend_ms_sum": float(agg[5]), "shadow_ms_sum": float(agg[6])}   # *_ms_sum: device time of the class summed over ranks
        return total_ms, job, tot, (t_w0, t_w1)

    def l2_peak():
        v = ctypes.c_double(0.0)
        if lib.bn_measure_l2_read_gbs(local, 32 << 20, 20, ctypes.byref(v)) == 0 and v.value > 0:
            return v.value
        return None

    def traversal_roofline(scene, w, h, job, l2_gbs):
        """Extend (closest-hit traversal) against the L2 roofline: rays x algorithmic bytes per ray (instrumented oracle on a
        bounded sample of this scene, REFERENCE layout and visiting order) / device time of the extend launches
        (CUDA events on the launching stream, BN_RENDER_PROFILE; summed over ranks, so the rate is per GPU x ranks)."""
        cw, ch = max(64, w // 8), max(64, h // 8)
        cst = oracle_sample(scene, cw, ch, 2, True, libm=False)
        b_ext = algorithmic_bytes_per_ray(cst["extend_counters"], B_IO_EXTEND)
        b_sh = algorithmic_bytes_per_ray(cst["shadow_counters"], B_IO_SHADOW)
        ext_s, sh_s = job["extend_ms_sum"] * 1e-3 / world, job["shadow_ms_sum"] * 1e-3 / world   # mean per-rank device time
        per_gpu = (job["extend"] / world) * b_ext / ext_s / 1e9 if ext_s > 0 else None           # one GPU's rate vs one GPU's L2
        return {"cw": cw, "ch": ch, "b_ext": b_ext, "b_sh": b_sh, "b_ext_cap": algorithmic_bytes_per_ray(cst["extend_counters"], B_IO_EXTEND, cap=True),
                "achieved": per_gpu, "frac_l2": (per_gpu / l2_gbs) if per_gpu and l2_gbs else None,
                "ext_rays_per_s": (job["extend"] / world) / ext_s if ext_s > 0 else None,
                "sh_rays_per_s": (job["shadow"] / world) / sh_s if sh_s > 0 else None,
                "sh_achieved": (job["shadow"] / world) * b_sh / sh_s / 1e9 if sh_s > 0 else None,
                "counters": {k: cst["extend_counters"][k] / max(cst["extend_counters"]["rays"], 1) for k in cst["extend_counters"] if k != "rays"}}

    scene = Scene.Load(os.path.join(ROOT, "scenes", scene_file), base_dir=ROOT)
    gpu = scene.gpu(local)
    film = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
    base = make_params(W, H, SPP, MAX_DEPTH, RR_DEPTH, flags=_ffi.BN_RENDER_PROFILE)
    # C4 is tile-split over the GPUs (BASELINE.json configs[3]); everything else by sampleId
    mode = "tile" if (args.workload == "C4" and world > 1) else "auto"
    clocks = ClockSampler(local) if rank == 0 else None
    total_ms, job, tot, (t_wall0, t_wall1) = timed_steps(gpu, base, film, args.steps, args.warmup, mode)
    rays_total = job["extend"] + job["shadow"]
    paths_total = job["paths"]
    launches_total = job["launches"]
    value = rays_total / (total_ms * 1e-3) / 1e6

    # ---- N > 1: is the reduced film right?  Rank 0 renders the same seeds on its own GPU (outside every timed region) and
    # compares with the film the last timed step left after the NCCL reduce.
    multi_check = None
    if world > 1:
        barrier()
        if rank == 0:
            reduced = film.clone()
            single = torch.zeros_like(film)
            gpu.render_device(make_params(W, H, SPP, MAX_DEPTH, RR_DEPTH), single.data_ptr(), stream)
            a, b = reduced.cpu().numpy(), single.cpu().numpy()
            v, skipped = rel_mse(a, b)
            multi_check = {"value": v, "eps": 1e-2, "non_finite_pixels_skipped": skipped, "max_abs_diff": float(np.nanmax(np.abs(a - b))),
                           "bit_identical": bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()),
                           "vs": f"the same {W}x{H}x{SPP}spp frame rendered on ONE GPU (rank 0, outside the timed region) against the film left by the "
                                 f"{world}-rank {'tile' if mode == 'tile' or SPP < world else 'sample'}-split step after the NCCL reduce; "
                                 "sample split differs by fp32 reassociation of the per-pixel sum only, tile split not at all"}
        barrier()

    # ---- e2e: through the C-ABI with host buffers (scene upload + film download per step)
    host_film = torch.empty(W * H * 3, dtype=torch.float32).pin_memory()
    host_film_np = host_film.numpy()
    shard = partition(W, H, SPP, world, rank, mode)
    sp = shard_params(make_params(W, H, SPP, MAX_DEPTH, RR_DEPTH), shard)
    d = scene.desc.contents
    h2d = (d.tlas_node_count + d.blas_node_count) * 32 + d.instance_count * 168 + d.vertex_count * 12 + d.triangle_count * 12 + d.alias_count * 12 + \
        d.sphere_count * 4 + d.material_count * 24 + d.light_count * 16 + d.light_instance_count * 4 + ctypes.sizeof(_ffi.BnCamera)
    d2h = W * H * 3 * 4

    # N = 1: bn_scene_create + bn_render.  N > 1: ONE process (rank 0) drives all N devices through bn_multi_scene_create +
    # bn_render_multi — the drop-in for Integrator.Render, whose parallelism lives inside the call (Integrator.fs:46-55): the
    # scene is flattened once and uploaded to every device, the shares rendered side by side, the films combined over NVLink
    # on device 0 and read back.  The other ranks keep their GPUs idle meanwhile (they wait on the CPU-side store, not in an
    # NCCL kernel, which would time-slice with rank 0's work on their devices).
    from barnacle_b200.scene import MultiGpuScene
    part_mode = {"tile": _ffi.BN_PARTITION_TILE, "auto": _ffi.BN_PARTITION_AUTO}[mode]
    full = make_params(W, H, SPP, MAX_DEPTH, RR_DEPTH)

    def e2e_step():
        if world > 1:
            m2 = MultiGpuScene(scene.desc, list(range(world)))   # host arrays -> N devices (flatten once, N uploads in parallel)
            try:
                _, st2 = m2.render(full, host_film_np, part_mode)   # film lands in the host buffer
            finally:
                m2.close()
            return st2
        g2 = GpuScene(scene.desc, local)               # host arrays -> device (flatten + cudaMemcpy H2D)
        try:
            _, st2 = g2.render(sp, host_film_np)            # bn_render: film lands in the host buffer
        finally:
            g2.close()
        return st2

    barrier()
    e2e_rays = 0
    e2e_step_ms = []
    e2e_s = 0.0
    e2e_check = None
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if rank == 0:
        e2e_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ts = time.perf_counter()
            st2 = e2e_step()
            e2e_step_ms.append((time.perf_counter() - ts) * 1e3)
            e2e_rays += st2.extend_rays + st2.shadow_rays
        e2e_s = time.perf_counter() - t0
        if world > 1:
            # the film bn_render_multi returned vs the film the torchrun step left after the NCCL reduce
            a, b = host_film_np.reshape(-1), film.cpu().numpy().reshape(-1)
            e2e_check = {"max_abs_diff_vs_nccl_path": float(np.nanmax(np.abs(a - b))), "bit_identical_to_nccl_path": bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())}
            store.set("bn_e2e_done", "1")
    elif store is not None:
        store.wait(["bn_e2e_done"])
    barrier()
    # stopped only now (an exiting NVML client can stall the next CUDA calls); its report covers t_wall0..t_wall1, the timed region
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    e2e_value = e2e_rays / e2e_s / 1e6 if rank == 0 else 0.0

    l2_gbs = l2_peak() if rank == 0 else None

    # ---- secondary configs: C1 / C3 / C4 at stated (reduced) spp so that the driver's record anchors every config, the
    # weakest one (C4, tile-split at N > 1) included.  Same timing rules as the main line; 1 warm-up + 2 timed steps each.
    SECONDARY_SPP = {"C1": 64, "C2": 32, "C3": 32, "C4": 8}
    configs = {}
    if not args.no_configs:
        for name in ("C1", "C2", "C3", "C4"):
            if name == args.workload:
                continue
            sf, w2, h2, spp_full, label2 = WORKLOADS[name]
            spp2 = min(spp_full, SECONDARY_SPP[name])
            mode2 = "tile" if (name == "C4" and world > 1) or spp2 < world else "sample"
            sc2 = Scene.Load(os.path.join(ROOT, "scenes", sf), base_dir=ROOT)
            g2 = sc2.gpu(local)
            film2 = torch.zeros(w2 * h2 * 3, dtype=torch.float32, device="cuda")
            ms2, job2, _, _ = timed_steps(g2, make_params(w2, h2, spp2, MAX_DEPTH, RR_DEPTH, flags=_ffi.BN_RENDER_PROFILE), film2, 2, 1, mode2)
            if rank == 0:
                rf = traversal_roofline(sc2, w2, h2, job2, l2_gbs)
                configs[name] = {"workload": label2 + (f" [{spp2} of {spp_full} spp per step; rates are spp-independent]" if spp2 != spp_full else ""),
                                 "value": (job2["extend"] + job2["shadow"]) / (ms2 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms2 / 2, "steps": 2, "warmup": 1,
                                 "samples_per_s": job2["paths"] / (ms2 * 1e-3), "parallelism": "single GPU" if world == 1 else f"{mode2}-split x{world} + film reduce",
                                 "extend_rays_per_s_per_gpu": rf["ext_rays_per_s"], "shadow_rays_per_s_per_gpu": rf["sh_rays_per_s"],
                                 "roofline": {"bound": "l2", "achieved": rf["achieved"], "peak": l2_gbs, "unit": "GB/s", "frac": rf["frac_l2"],
                                              "algorithmic_bytes_per_ray": rf["b_ext"]}}
            g2.close()
            sc2.close()
            del film2

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, hbm_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        else:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        rf = traversal_roofline(scene, W, H, job, l2_gbs)
        # rank 0's launches per wave: raygen + accumulate + (extend, shade, shadow) per bounce; 5 per bounce when every
        # traversal launch is followed by a fix-up launch (BN_SEPARATE_FIXUP / a BN_INKERNEL_DRAIN=0 build)
        per_wave = next((2 + k * MAX_DEPTH for k in ((5, 3) if os.environ.get("BN_SEPARATE_FIXUP") else (3, 5)) if tot["launches"] % (2 + k * MAX_DEPTH) == 0), 2 + 3 * MAX_DEPTH)
        n_ext_launches = max(1, tot["launches"] // per_wave * MAX_DEPTH)
        rays_per_launch = tot["extend"] / n_ext_launches
        # ncu figures of the committed kernel on this workload (profiles/ncu_summary.json, made by tools/ncu_to_json.py from the
        # committed capture): per-ray numbers of ONE profiled launch, scaled to this run's average launch
        ncu = {}
        npath = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(npath):
            ncu = json.load(open(npath)).get(args.workload, {})
        ne = ncu.get("extend", {})
        traffic = ne["dram_bytes_per_ray"] * rays_per_launch if ne.get("dram_bytes_per_ray") else None
        roofline = {
            # the scene data of every config is L2-resident (ncu: DRAM traffic is the ray queue in and the hit record out, ~5 % of
            # the algorithmic bytes), so the memory roofline of the traversal is the L2 read bandwidth, measured in this run
            "bound": "l2", "kernel": "k_traverse<closest> (extend: TLAS+BLAS closest-hit traversal)",
            "achieved": rf["achieved"], "peak": l2_gbs, "unit": "GB/s", "frac": rf["frac_l2"],
            "peak_source": "bn_measure_l2_read_gbs: 20 sweeps of a 32 MiB L2-resident buffer, ld.global.cg.v4, CUDA events, this run, rank 0's GPU",
            "traffic": traffic, "traffic_source": (ncu.get("source", "") + f"; {ne['dram_bytes_per_ray']:.1f} DRAM B/ray of the profiled launch x the rays of this run's average extend launch") if traffic else None,
            "lanes_per_inst": ne.get("lanes_per_inst"), "warp_inst_per_ray": ne.get("warp_inst_per_ray"), "ncu_source": ncu.get("source"),
            "algorithmic_bytes_per_launch": rays_per_launch * rf["b_ext"], "avg_launch_ms": tot["extend_ms"] / n_ext_launches,
            "algorithmic_bytes_per_ray": rf["b_ext"], "algorithmic_bytes_per_ray_cap152": rf["b_ext_cap"], "per_ray_counters": rf["counters"],
            "rays_per_launch": rays_per_launch, "rays_per_s": rf["ext_rays_per_s"],
            "kernel_share_of_step": tot["extend_ms"] / tot["render_ms"] if tot["render_ms"] else None,
            "hbm": {"peak": hbm_peak, "peak_source": hbm_src, "frac_algorithmic": (rf["achieved"] / hbm_peak) if rf["achieved"] else None,
                    "frac_dram_traffic": (traffic / (tot["extend_ms"] / n_ext_launches * 1e-3) / 1e9 / hbm_peak) if traffic and tot["extend_ms"] else None,
                    "note": "frac_algorithmic divides traffic served by L1/L2 by a DRAM-copy peak (can exceed 1, says nothing); frac_dram_traffic is ncu's DRAM bytes over the same peak"},
            "shadow": {"algorithmic_bytes_per_ray": rf["b_sh"], "achieved": rf["sh_achieved"], "frac": (rf["sh_achieved"] / l2_gbs) if rf["sh_achieved"] and l2_gbs else None,
                       "rays_per_s": rf["sh_rays_per_s"], "lanes_per_inst": ncu.get("shadow", {}).get("lanes_per_inst"),
                       "warp_inst_per_ray": ncu.get("shadow", {}).get("warp_inst_per_ray")},
            "rank0_class_ms_per_step": {k: tot[k + "_ms"] / args.steps for k in ("extend", "shade", "shadow", "other")},
            "note": "bytes/ray = 32*N_node + 48*N_tri + (24*N_inst_visited + 64*N_inst_boxpass + 64*N_inst_committed) + 48 I/O in the REFERENCE layout (SURVEY 8d), counted by the "
                    f"instrumented oracle on {rf['cw']}x{rf['ch']}x2spp of this scene; achieved = one GPU's extend rays x bytes/ray / its extend device time"}
        cpu = relmse = None
        if not args.no_cpu_baseline and world == 1:
            # ~10-30 s of CPU work on all host threads: about 30 M paths of the workload's film (libm math, as the reference)
            cpu_spp = max(1, min(SPP, -(-30_000_000 // (W * H))))
            cs = oracle_sample(scene, W, H, cpu_spp, False, keep_film=True)
            cpu = {"value": (cs["extend_rays"] + cs["shadow_rays"]) / cs["seconds"] / 1e6, "unit": "Mrays/s", "cores": cs["threads"], "kind": "port",
                   "samples_per_s": cs["paths"] / cs["seconds"],
                   "sample": f"{W}x{H} at {cpu_spp} of {SPP} spp ({cs['seconds']:.1f} s); rays = extend + shadow rays the reference traces",
                   "note": "C++ restatement of Barnacle's CPU path (the .NET binary is not runnable in this image)"}
            # BASELINE's second metric, relMSE of the linear film against the CPU path under identical seeds: the same
            # sample of the workload rendered here on the GPU (outside every timed region) against the film just made
            try:
                gfilm, _ = gpu.render(make_params(W, H, cpu_spp, MAX_DEPTH, RR_DEPTH))
                v, skipped = rel_mse(gfilm, cs["film"])
                relmse = {"value": v, "eps": 1e-2, "non_finite_pixels_skipped": skipped,
                          "vs": f"CPU restatement (libm transcendentals, as the reference) on the same seeds, {W}x{H} at {cpu_spp} spp; "
                                "against the restatement with this library's fixed fp32 transcendentals the film is bit-identical (tests/)"}
            except Exception as e:  # a metric must never take the bench line down with it
                relmse = {"value": None, "error": f"{type(e).__name__}: {e}"}
        if world > 1:
            relmse = multi_check
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": label, "max_depth": MAX_DEPTH, "rr_depth": RR_DEPTH,
                       "parallelism": (f"{'tile' if shard.interleave_count > 1 else 'sample'}-split x{world} + film reduce") if world > 1 else "single GPU",
                       "l2_flush": "256 MiB memset between steps", "wave_paths": int(os.environ.get("BN_WAVE_PATHS", 64 << 20))},
            "samples_per_s": paths_total / (total_ms * 1e-3),
            "rays_per_step": rays_total / args.steps, "paths_per_step": paths_total / args.steps,
            "gpu_launches": launches_total,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                    "rank0_step_ms": [round(x, 2) for x in e2e_step_ms],
                    "path": "bn_scene_create + bn_render (host scene arrays in, host film out)" if world == 1 else
                            f"ONE process: bn_multi_scene_create + bn_render_multi over {world} devices (host scene arrays in, flattened once, uploaded to every device; "
                            "films combined over NVLink by this library's own kernel; host film out) — no torch / NCCL on this path",
                    "check": e2e_check},
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "relmse": relmse, "configs": configs,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
