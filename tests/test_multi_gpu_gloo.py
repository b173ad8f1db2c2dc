"""The N > 1 path on CPU: world_size-2 `gloo` run of barnacle_b200.multi_gpu —
partitioning (sample split / tile-row interleave) and the single sum-reduce of the
film to rank 0.  The per-rank renderer here is the oracle standing in for the CUDA
library (this is a test of the host-side sharding logic, not of the kernels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackedScene:
    """Duck-types GpuScene.render_device(params, ptr, stream) on CPU tensors."""

    def __init__(self, scene):
        from oracle.oracle_ffi import OracleScene
        self.o = OracleScene(scene.desc)
        self.films = {}

    def bind(self, film):
        self.films[film.data_ptr()] = film

    def render_device(self, p, ptr, stream=0):
        # interleave is a property of the CUDA library; emulate it with per-row windows
        from barnacle_b200.scene import make_params
        film = self.films[ptr]
        out = np.zeros((p.height * p.width, 3), np.float32)
        rows = range((p.height + 15) // 16)
        cnt, idx = max(p.interleave_count, 1), (p.interleave_index if p.interleave_count > 1 else 0)
        for r in rows:
            if r % cnt != idx:
                continue
            q = make_params(p.width, p.height, p.spp, p.max_depth, p.rr_depth, p.frame_id, p.sample_begin, p.sample_end,
                            rect=(0, r * 16, p.width, min((r + 1) * 16, p.height)))
            f, _ = self.o.render(q)
            out += f
        film.copy_(torch.from_numpy(out.reshape(-1)))
        return None


def _worker(rank, world, port, mode, spp, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from barnacle_b200.multi_gpu import render_sharded
    from barnacle_b200.scene import Scene, make_params
    scene = Scene.Load(os.path.join(ROOT, "scenes", "cbox_pt.json"), base_dir=ROOT)
    fake = OracleBackedScene(scene)
    W, H = 40, 40
    film = torch.zeros(W * H * 3, dtype=torch.float32)
    fake.bind(film)
    render_sharded(fake, make_params(W, H, spp), film, dist, mode=mode)
    if rank == 0:
        np.save(out_path, film.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,spp", [("sample", 4), ("tile", 1), ("auto", 3)])
def test_world2_gloo_film_reduce(tmp_path, mode, spp, lib):
    port = 29500 + os.getpid() % 2000 + {"sample": 0, "tile": 1, "auto": 2}[mode]
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, port, mode, spp, out), nprocs=2, join=True)
    from barnacle_b200.scene import Scene, make_params
    from oracle.oracle_ffi import OracleScene
    scene = Scene.Load(os.path.join(ROOT, "scenes", "cbox_pt.json"), base_dir=ROOT)
    ref, _ = OracleScene(scene.desc).render(make_params(40, 40, spp))
    got = np.load(out).reshape(-1, 3)
    if mode == "tile" or spp < 2:
        assert np.array_equal(got, ref)                 # disjoint tiles: the sum is a gather
    else:
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)   # sample split: fp32 reassociation only


def test_partition_covers_everything_once():
    from barnacle_b200.multi_gpu import owned_pixels, partition
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 5, 8, 256):
            for (w, h) in ((64, 64), (100, 37), (3840, 2160)):
                shards = [partition(w, h, spp, world, r) for r in range(world)]
                work = sum((s.sample_end - s.sample_begin) * owned_pixels(w, h, s) for s in shards if not s.empty)
                assert work == w * h * spp
                if spp >= world:   # sample split: contiguous, ordered, balanced to within one sample
                    assert shards[0].sample_begin == 0 and shards[-1].sample_end == spp
                    assert all(a.sample_end == b.sample_begin for a, b in zip(shards, shards[1:]))
                    sizes = [s.sample_end - s.sample_begin for s in shards]
                    assert max(sizes) - min(sizes) <= 1
