"""The ordering of the live paths between bounces (barnacle_b200/csrc/cuda/ray_sort.cuh) tested on its own: the library's
counting sort (k_sort_hist / k_sort_scan / k_sort_rank through the parity-test entry bn_debug_order_keys) must return a
PERMUTATION of the queue along which the 12-bit keys never decrease — sortedness and bijectivity are the size-independent
properties of the step; the order inside a bin is free.  (The reference has no such step — its paths are independent loop
iterations, Integrator.fs:34-44 — so there is no reference output to compare with; that the film does not depend on the
order is tests/test_gpu_render_parity.py::test_path_ordering_does_not_change_the_film.)"""
import ctypes as C

import numpy as np
import pytest

from barnacle_b200 import _ffi

pytestmark = pytest.mark.gpu


def _order(keys):
    lib = _ffi.load()
    keys = np.ascontiguousarray(keys, dtype=np.uint16)
    perm = np.empty(len(keys), dtype=np.uint32)
    _ffi.check(lib.bn_debug_order_keys(0, keys.ctypes.data_as(C.c_void_p), len(keys), perm.ctypes.data_as(C.c_void_p)), "bn_debug_order_keys")
    return perm


CASES = {
    "uniform": lambda rng, n: rng.integers(0, 4096, n),
    "one bin": lambda rng, n: np.full(n, 1234),
    "two bins": lambda rng, n: rng.integers(0, 2, n) * 4095,
    "already sorted": lambda rng, n: np.sort(rng.integers(0, 4096, n)),
    "reversed": lambda rng, n: np.sort(rng.integers(0, 4096, n))[::-1],
    "clustered": lambda rng, n: np.clip(rng.normal(2000, 30, n), 0, 4095).astype(np.int64),
}


@pytest.mark.parametrize("n", [1, 31, 4095, 4096, 4097, 100_003, 3_000_000])
@pytest.mark.parametrize("case", list(CASES))
def test_ordering_is_a_sorted_permutation(case, n):
    rng = np.random.default_rng(n * 7 + len(case))
    keys = CASES[case](rng, n).astype(np.uint16)
    perm = _order(keys)
    assert perm.max() < n
    seen = np.zeros(n, dtype=bool)
    seen[perm] = True
    assert seen.all(), "not a permutation: some queue index is missing (and another one doubled)"
    along = keys[perm].astype(np.int32)
    assert (np.diff(along) >= 0).all(), "keys decrease along the ordering"


def test_empty_queue_and_bad_arguments():
    lib = _ffi.load()
    assert lib.bn_debug_order_keys(0, None, 0, None) == _ffi.BN_OK
    assert lib.bn_debug_order_keys(0, None, 5, None) == _ffi.BN_ERR_INVALID
    assert lib.bn_debug_order_keys(99, None, 0, None) in (_ffi.BN_ERR_NO_DEVICE, _ffi.BN_OK)
