"""Film.PostProcess + the Rgba32 conversion of Film.Save (Base/Film.fs:21-30,55-66) on the host side of the C ABI
(bn_host_film_to_rgba8) against a float64 numpy restatement, and Film.SetPixel's flipped-Y layout (Film.fs:41-46)
as bn_render's callers see it.  CPU only."""
import ctypes

import numpy as np
import pytest

from barnacle_b200 import _ffi
from barnacle_b200.scene import Film


def ref_post_process(x, tone):  # Film.fs:21-30; ImageSharp's Rgba32(Vector3): x * 255 + 0.5, truncated
    x = x.astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        if tone == "aces":
            x = x * (2.51 * x + 0.03) / (x * (2.43 * x + 0.59) + 0.14)
        elif tone == "gamma":
            x = np.power(x, 1 / 2.2)
    x = np.where(np.isnan(x), 0.0, np.clip(x, 0.0, 1.0))
    return x * 255 + 0.5


@pytest.mark.parametrize("tone", ["identity", "aces", "gamma"])
def test_film_to_rgba8_matches_restatement(lib, tone):
    rng = np.random.default_rng(5)
    w, h = 37, 23
    film = Film(w, h, tone)
    film.Pixels[:] = np.exp(rng.normal(-1.0, 2.0, size=(h * w, 3))).astype(np.float32)   # 1e-4 .. 1e2: both clamps are hit
    film.Pixels[:5] = [[0, 0, 0], [1, 1, 1], [-0.5, 2.0, 0.25], [np.nan, np.inf, 1e-30], [0.18, 0.5, 0.9]]
    got = film.to_rgba8()
    want = ref_post_process(film.Pixels, tone).reshape(h, w, 3)
    assert (got[..., 3] == 255).all()
    # the fp32 evaluation may land on the other side of an 8-bit step only where the exact value is within rounding of it
    diff = np.abs(got[..., :3].astype(np.int64) - np.floor(want).astype(np.int64))
    near_step = np.abs(want - np.round(want)) < 2e-3
    assert (diff[~near_step] == 0).all() and diff.max() <= 1
    if tone == "identity":
        assert got[0, 0, :3].tolist() == [0, 0, 0] and got[0, 1, :3].tolist() == [255, 255, 255]
        assert got[0, 2, :3].tolist() == [0, 255, 64]              # clamps; 0.25 * 255 + 0.5 = 64.25
        assert got[0, 3, :3].tolist() == [0, 255, 0]               # NaN -> 0, +inf -> 1


@pytest.mark.parametrize("tone", ["identity", "aces", "gamma"])
def test_host_film_to_rgba8_equals_the_oracle_restatement(lib, oracle_lib, tone):
    """The product's host-side conversion and the oracle's restatement of Film.PostProcess / Rgba32 (oracle/barnacle_oracle.cpp:
    bo_film_to_rgba8) byte for byte — both call glibc's powf for the gamma curve, as MathF.Pow does on Linux."""
    from oracle import oracle_ffi
    rng = np.random.default_rng(7)
    w, h = 64, 31
    film = Film(w, h, tone)
    film.Pixels[:] = np.exp(rng.normal(-1.0, 2.5, size=(h * w, 3))).astype(np.float32)
    film.Pixels[:6] = [[0, 0, 0], [1, 1, 1], [-0.5, 2.0, 0.25], [np.nan, np.inf, 1e-30], [0.18, 0.5, 0.9], [-np.inf, -0.0, 1e30]]
    want = oracle_ffi.film_to_rgba8(film.Pixels, w, h, {"identity": 0, "aces": 1, "gamma": 2}[tone])
    assert np.array_equal(film.to_rgba8(), want)


def test_film_to_rgba8_rejects_bad_arguments(lib):
    out = (ctypes.c_uint8 * 16)()
    assert lib.bn_host_film_to_rgba8(None, 2, 2, 0, out) == _ffi.BN_ERR_INVALID
    assert b"bad arguments" in lib.bn_last_error()
