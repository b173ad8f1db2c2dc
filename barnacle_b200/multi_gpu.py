"""Multi-GPU driver for the hot path: one process per GPU, scene replicated, the
(pixel, sampleId) space sharded, ONE collective — a sum-reduce of the fp32 film to
rank 0 (NCCL over NVLink; gloo on CPU for tests).

The reference shards 16x16 tiles over TPL threads with one writer per pixel
(Base/Integrator.fs:46-54); every (x, y, sampleId) path is independent and fully
determined by its seed (Integrator.fs:35-36), so any partition renders the same
paths.  Two partitions:
  * sample split: rank g renders sampleIds [g*spp/G, (g+1)*spp/G) of every pixel,
    weight 1/spp — perfectly balanced; used when spp >= G.
  * tile split: rank g renders the 16-pixel tile rows g, g+G, g+2G, ... (round-robin
    interleave keeps the load balanced) — disjoint film regions, zero elsewhere, so a
    sum is a gather; used when spp < G.
"""
from __future__ import annotations

from dataclasses import dataclass

from ._ffi import BnRenderParams

TILE = 16  # ProgressiveIntegrator tile size (Integrator.fs:16)


@dataclass(frozen=True)
class Shard:
    sample_begin: int
    sample_end: int
    interleave_count: int = 1   # tile-row interleave (BnRenderParams.interleave_*)
    interleave_index: int = 0

    @property
    def empty(self) -> bool:
        return self.sample_begin >= self.sample_end


def partition(width: int, height: int, spp: int, world: int, rank: int, mode: str = "auto") -> Shard:
    """The part of the (pixel, sampleId) space rank `rank` of `world` renders."""
    assert 0 <= rank < world
    if mode == "auto":
        mode = "sample" if spp >= world else "tile"
    if mode == "sample":
        return Shard((rank * spp) // world, ((rank + 1) * spp) // world)
    if mode == "tile":
        rows = (height + TILE - 1) // TILE
        if rank >= rows:
            return Shard(0, 0)
        return Shard(0, spp, world, rank)
    raise ValueError(f"unknown partition mode {mode!r}")


def shard_params(base: BnRenderParams, shard: Shard) -> BnRenderParams:
    return BnRenderParams(base.width, base.height, base.spp, base.max_depth, base.rr_depth, base.frame_id,
                          shard.sample_begin, shard.sample_end, base.x0, base.y0, base.x1, base.y1, base.flags,
                          shard.interleave_count, shard.interleave_index, base.integrator)


def owned_pixels(width: int, height: int, shard: Shard) -> int:
    """Pixels of the full-frame window this shard renders (host-side bookkeeping)."""
    if shard.interleave_count <= 1:
        return width * height
    rows = range(shard.interleave_index, (height + TILE - 1) // TILE, shard.interleave_count)
    return sum(width * (min((r + 1) * TILE, height) - r * TILE) for r in rows)


def partition_chains(n_chains: int, world: int, rank: int) -> tuple[int, int]:
    """PSSMLT: chains are independent given the (replicated) bootstrap; rank g runs a contiguous range."""
    return (rank * n_chains) // world, ((rank + 1) * n_chains) // world


def render_pssmlt_sharded(gpu_scene, base, film, dist=None, stream: int = 0):
    """PSSMLT across GPUs: every rank runs the whole bootstrap (same B everywhere), its share of the
    chains, splats into its own film (must be zeroed by the caller), one sum-reduce to rank 0."""
    from ._ffi import BnMltParams
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    b, e = partition_chains(base.n_chains, world, rank)
    p = BnMltParams(base.width, base.height, base.mutations_per_pixel, base.max_depth, base.rr_depth, base.frame_id, base.n_bootstrap,
                    base.n_chains, base.strategy, base.p0, base.p1, base.large_step_prob, b, e)
    stats = gpu_scene.render_pssmlt_device(p, film.data_ptr(), stream)
    if world > 1:
        dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
    return stats


def render_sharded(gpu_scene, base: BnRenderParams, film, dist=None, stream: int = 0, mode: str = "auto"):
    """Render this rank's shard into `film` (a torch CUDA float32 tensor of W*H*3 on
    the scene's device; pixels outside the shard are written as 0) and sum-reduce to
    rank 0.  Returns this rank's RenderStats (None for an empty shard)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    shard = partition(base.width, base.height, base.spp, world, rank, mode)
    stats = None
    if shard.empty:
        film.zero_()
    else:
        stats = gpu_scene.render_device(shard_params(base, shard), film.data_ptr(), stream)
    if world > 1:
        dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
    return stats
