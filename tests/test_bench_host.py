"""Host-side logic of bench.py that needs no GPU: the algorithmic-bytes formula of SURVEY §8(d), the relMSE metric,
the sharding plan of --gpus N, and the JSON contract of the reference arm (`--impl reference` runs the CPU path only)."""
import json
import os
import subprocess
import sys

import numpy as np

import bench


def test_algorithmic_bytes_formula():
    c = {"tlas_nodes": 30, "blas_nodes": 70, "tris_fetched": 20, "tris_box_pass": 5, "inst_visited": 10, "inst_box_pass": 4,
         "inst_committed": 2, "rays": 10}
    # 32 B per node popped, 48 B per triangle fetched, 24 / +64 / +64 B per instance visited / entered / committed, + I/O
    want = (32 * 100 + 48 * 20 + 24 * 10 + 64 * 4 + 64 * 2) / 10 + bench.B_IO_EXTEND
    assert bench.algorithmic_bytes_per_ray(c, bench.B_IO_EXTEND) == want
    assert bench.algorithmic_bytes_per_ray(c, bench.B_IO_EXTEND, cap=True) == (32 * 100 + 48 * 20 + 152 * 10) / 10 + bench.B_IO_EXTEND
    assert bench.B_IO_EXTEND == 28 + 20 and bench.B_IO_SHADOW == 28 + 4


def test_rel_mse():
    ref = np.full((6, 3), 2.0, dtype=np.float32)
    img = ref.copy()
    assert bench.rel_mse(img, ref) == (0.0, 0)
    img[0] = 3.0                                                   # one of six pixels off by 1: (1 / (4 + eps)) / 6
    v, skipped = bench.rel_mse(img, ref, eps=1e-2)
    assert skipped == 0 and abs(v - (1 / 4.01) / 6) < 1e-12
    img[1, 2] = np.nan
    v2, skipped = bench.rel_mse(img, ref, eps=1e-2)
    assert skipped == 1 and abs(v2 - (1 / 4.01) / 5) < 1e-12
    assert bench.rel_mse(np.full((2, 3), np.nan), np.zeros((2, 3))) == (None, 2)


def test_oracle_sample_keeps_the_film_on_request(scene_loader):
    scene = scene_loader("cbox_pt")
    st = bench.oracle_sample(scene, 32, 32, 1, True, keep_film=True)
    assert st["film"].shape == (32 * 32, 3) and st["paths"] == 32 * 32 and st["extend_counters"]["rays"] == st["extend_rays"]
    assert "film" not in bench.oracle_sample(scene, 32, 32, 1, False)


def test_reference_arm_prints_the_contract_line(root):
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0",
                        "--ref-spp", "1"], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C1")


def test_reference_arm_other_ranks_exit_without_work(root):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_every_host_thread_under_torchrun(root):
    """torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU arm must not follow it (round 1's scaling record ran the
    reference on one core for N >= 2).  The thread count is handed to the oracle explicitly."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--ref-spp", "1"], cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == bench.host_threads() == len(os.sched_getaffinity(0))
    if bench.host_threads() > 1:
        assert d["cpu_baseline"]["cores"] > 1


def test_reference_arm_sizes_its_sample_for_about_a_minute(root):
    """--ref-spp 0 (the default): the bounded sample is sized from a calibration pass; the line says what it was."""
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "5", "--warmup", "1", "--ref-seconds", "4"],
                       cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["steps"] == 5 and "of 64 spp per step" in d["cpu_baseline"]["sample"]
    assert d["ms_per_step"] * 5 < 30e3
