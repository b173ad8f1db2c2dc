"""Several devices behind one C-ABI call (bn_multi_scene_create / bn_render_multi, include/barnacle_b200.h): the reference keeps
the machine's parallelism inside Integrator.Render (Base/Integrator.fs:46-55), so the drop-in fans out inside the call.
CPU half: the sharding rule (bn_multi_partition) against barnacle_b200/multi_gpu.partition and its coverage properties.
GPU half (-m gpu): the same device listed twice exercises the whole path (threads, per-device films, the peer-read combine
kernel) on a 1-GPU box; with two real devices the test also runs across NVLink."""
import numpy as np
import pytest

from barnacle_b200 import _ffi
from barnacle_b200.multi_gpu import TILE, partition
from barnacle_b200.scene import MultiGpuScene, make_params, multi_partition
from conftest import load_scene


@pytest.mark.parametrize("w,h,spp", [(64, 48, 8), (100, 70, 3), (33, 17, 1), (512, 512, 64)])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 8])
def test_partition_matches_the_python_driver_and_covers_every_path_once(lib, w, h, spp, n):
    base = make_params(w, h, spp)
    covered = np.zeros((h, spp), dtype=np.int32)            # per (pixel row, sampleId): how many devices render it
    for rank in range(n):
        sp, empty = multi_partition(base, _ffi.BN_PARTITION_AUTO, n, rank)
        want = partition(w, h, spp, n, rank)
        assert empty == want.empty
        if empty:
            continue
        ic = sp.interleave_count if sp.interleave_count > 1 else 1
        assert (sp.sample_begin, sp.sample_end, ic, sp.interleave_index if ic > 1 else 0) == \
               (want.sample_begin, want.sample_end, want.interleave_count, want.interleave_index)
        rows = [y for y in range(h) if ic == 1 or (y // TILE) % ic == sp.interleave_index]
        covered[np.ix_(rows, range(sp.sample_begin, sp.sample_end))] += 1
    assert (covered == 1).all()


def test_partition_modes_and_errors(lib):
    base = make_params(64, 64, 8)
    s, _ = multi_partition(base, _ffi.BN_PARTITION_TILE, 4, 1)
    assert (s.interleave_count, s.interleave_index, s.sample_begin, s.sample_end) == (4, 1, 0, 8)
    s, _ = multi_partition(base, _ffi.BN_PARTITION_SAMPLE, 4, 3)
    assert (s.sample_begin, s.sample_end) == (6, 8) and s.interleave_count <= 1
    sub = make_params(64, 64, 8, sample_begin=2, sample_end=7)    # a window of its own is shared out, not the whole frame
    got = [multi_partition(sub, _ffi.BN_PARTITION_SAMPLE, 2, r)[0] for r in range(2)]
    assert [(g.sample_begin, g.sample_end) for g in got] == [(2, 4), (4, 7)]
    already = make_params(64, 64, 8, interleave=(2, 1))
    with pytest.raises(_ffi.BarnacleError, match="tile-interleaved"):
        multi_partition(already, _ffi.BN_PARTITION_TILE, 2, 0)
    with pytest.raises(_ffi.BarnacleError):
        multi_partition(base, 7, 2, 0)


def test_multi_scene_needs_a_device(lib, scene_loader):
    if lib.bn_device_count() > 0:
        return
    with pytest.raises(_ffi.BarnacleError, match="no CUDA device"):
        MultiGpuScene(scene_loader("cbox_pt").desc, [0, 1])


def _bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spp", [("cbox_pt", 96, 80, 6), ("cbox_bunny", 64, 64, 5)])
def test_render_multi_equals_one_device(name, w, h, spp):
    lib = _ffi.load()
    scene = load_scene(name)
    single, st1 = scene.gpu().render(make_params(w, h, spp))
    lists = [[0], [0, 0], [0, 0, 0]]
    if lib.bn_device_count() >= 2:
        lists += [[0, 1], [1, 0]]
    for devices in lists:
        m = MultiGpuScene(scene.desc, devices)
        try:
            tile, st_t = m.render(make_params(w, h, spp), partition=_ffi.BN_PARTITION_TILE)
            assert _bits_equal(tile, single), devices                 # disjoint pixels: the combine is a gather
            samp, st_s = m.render(make_params(w, h, spp), partition=_ffi.BN_PARTITION_SAMPLE)
            np.testing.assert_allclose(samp, single, rtol=2e-6, atol=1e-7)   # fp32 re-association of the per-pixel sum only
            auto, _ = m.render(make_params(w, h, spp))
            assert _bits_equal(auto, samp if spp >= len(devices) else tile)
            for st in (st_t, st_s):
                assert (st.paths, st.extend_rays, st.shadow_rays) == (st1.paths, st1.extend_rays, st1.shadow_rays)
            again, _ = m.render(make_params(w, h, spp), partition=_ffi.BN_PARTITION_SAMPLE)
            assert _bits_equal(again, samp)                              # fixed device order: the same film every run
        finally:
            m.close()


@pytest.mark.gpu
def test_render_multi_more_devices_than_samples_and_bad_arguments():
    scene = load_scene("cbox_pt")
    single, _ = scene.gpu().render(make_params(64, 40, 1))
    m = MultiGpuScene(scene.desc, [0, 0, 0])
    try:
        film, _ = m.render(make_params(64, 40, 1))                       # 1 spp over 3 devices: tile rows 0,3 | 1 | 2
        assert _bits_equal(film, single)
        with pytest.raises(_ffi.BarnacleError):
            m.render(make_params(64, 40, 1), partition=9)
        with pytest.raises(_ffi.BarnacleError, match="tile-interleaved"):
            m.render(make_params(64, 40, 1, interleave=(2, 0)))
    finally:
        m.close()
    with pytest.raises(_ffi.BarnacleError, match="out of range"):
        MultiGpuScene(scene.desc, [0, 99])
