"""Accuracy of the bit-reproducible transcendental definitions (include/bn_portable_math.h)
against double-precision libm, and agreement of the oracle's two math modes.  The
measured bounds are the ones DESIGN.md quotes."""
import ctypes

import numpy as np

from oracle import oracle_ffi


def ulp_err(got32, exact64):
    exact32 = exact64.astype(np.float32)
    ulp = np.spacing(np.abs(exact32)).astype(np.float64)
    ulp = np.maximum(ulp, np.finfo(np.float32).tiny)
    return np.abs(got32.astype(np.float64) - exact64) / ulp


def eval_sincos(lib, xs):
    s, c = ctypes.c_float(), ctypes.c_float()
    out = np.empty((len(xs), 2), np.float32)
    for i, x in enumerate(xs):
        lib.bo_sincos(float(x), ctypes.byref(s), ctypes.byref(c))
        out[i] = (s.value, c.value)
    return out


def test_sincos_accuracy(oracle_lib):
    oracle_ffi.set_portable_math(True)
    rng = np.random.Generator(np.random.PCG64(1))
    xs = np.concatenate([rng.uniform(-np.pi, 2 * np.pi + 0.01, 40000), np.linspace(0, 2 * np.pi, 2001), [0.0, np.pi / 2, np.pi, 1.5 * np.pi]]).astype(np.float32)
    got = eval_sincos(oracle_lib, xs)
    x64 = xs.astype(np.float64)
    # absolute error (what matters for directions): < 2.5e-7; ulp error away from zeros of the function: <= 2.5
    assert np.abs(got[:, 0] - np.sin(x64)).max() < 2.5e-7
    assert np.abs(got[:, 1] - np.cos(x64)).max() < 2.5e-7
    big = np.abs(np.sin(x64)) > 0.05
    assert ulp_err(got[big, 0], np.sin(x64[big])).max() <= 2.5
    big = np.abs(np.cos(x64)) > 0.05
    assert ulp_err(got[big, 1], np.cos(x64[big])).max() <= 2.5
    # sin^2 + cos^2 stays 1 to fp32 accuracy
    assert np.abs(got[:, 0].astype(np.float64) ** 2 + got[:, 1].astype(np.float64) ** 2 - 1).max() < 4e-7


def test_atan_accuracy(oracle_lib):
    oracle_ffi.set_portable_math(True)
    rng = np.random.Generator(np.random.PCG64(2))
    xs = np.concatenate([rng.uniform(0, 4, 20000), 10 ** rng.uniform(-6, 6, 20000), -(10 ** rng.uniform(-3, 3, 2000)), [0.0, 1.0, 0.41421357, 2.4142137]]).astype(np.float32)
    got = np.array([oracle_lib.bo_atan(float(x)) for x in xs], np.float32)
    assert ulp_err(got, np.arctan(xs.astype(np.float64))).max() <= 3.0   # measured 2.7 ulp just above tan(pi/8)
    assert np.abs(got - np.arctan(xs.astype(np.float64))).max() < 2e-7
    assert oracle_lib.bo_atan(0.0) == 0.0


def test_libm_mode_is_libm(oracle_lib):
    oracle_ffi.set_portable_math(False)
    try:
        xs = np.linspace(-3, 6, 500).astype(np.float32)
        got = eval_sincos(oracle_lib, xs)
        assert np.array_equal(got[:, 0], np.sin(xs)) or np.abs(got[:, 0] - np.sin(xs.astype(np.float64))).max() < 1.2e-7
    finally:
        oracle_ffi.set_portable_math(True)
