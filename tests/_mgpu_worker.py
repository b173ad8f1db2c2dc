"""torchrun worker for tests/test_gpu_multi.py: every rank renders its shard on its own
GPU, one NCCL reduce, rank 0 compares with a single-GPU render of the whole window."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from barnacle_b200.multi_gpu import render_sharded  # noqa: E402
from barnacle_b200.scene import Scene, make_params  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = Scene.Load(os.path.join(ROOT, "scenes", "cbox_bunny.json"), base_dir=ROOT)
    gpu = scene.gpu(local)
    W, H = 96, 80
    stream = torch.cuda.current_stream().cuda_stream
    for mode, spp in (("sample", 6), ("tile", 1), ("tile", 3)):
        film = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
        render_sharded(gpu, make_params(W, H, spp), film, dist, stream, mode=mode)
        torch.cuda.synchronize()
        if rank == 0:
            ref, _ = gpu.render(make_params(W, H, spp))
            got = film.cpu().numpy().reshape(-1, 3)
            if mode == "tile":
                assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "tile split must be bit-exact"
            else:
                np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK")


if __name__ == "__main__":
    main()
