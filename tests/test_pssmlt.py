"""PSSMLT ("next" row N1, Extensions/Integrator/PSSMLT.fs): the oracle's restatement (CPU tests)
and the CUDA chains against it (GPU tests).  With the portable log/exp/sin/cos definitions a chain
evolves bit-identically on both sides: bootstrap weights, B and every chain's accepted-mutation
count must be EQUAL; the film differs only by the order of the atomic fp32 splats."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_scene
from barnacle_b200 import _ffi
from barnacle_b200.scene import make_mlt_params
from oracle.oracle_ffi import OracleScene, set_portable_math

W, H = 48, 40


def small_params(strategy="Gaussian", **kw):
    return make_mlt_params(W, H, 6, n_bootstrap=4096, n_chains=96, strategy=strategy, **kw)


def test_oracle_pssmlt_is_a_sane_estimator(lib):
    """The MLT film must agree in the mean with the path-traced film of the same scene (both are
    estimators of the same image up to the reference's biases): mean radiance within 15%."""
    from barnacle_b200.scene import make_params
    set_portable_math(True)
    scene = load_scene("cbox_pt")
    o = OracleScene(scene.desc)
    film, st, per_chain = o.render_pssmlt(make_mlt_params(W, H, 64, n_bootstrap=16384, n_chains=128))
    pt, _ = o.render(make_params(W, H, 64))
    assert st["proposed"] == 128 * ((64 * W * H + 127) // 128) and 0.05 < st["accepted"] / st["proposed"] < 0.95
    # A chain whose start state has zero luminance (possible because the bootstrap pick is uniform,
    # SURVEY Q1) computes accept = Min(1, 0/0) = NaN and splats NaN radiance (PSSMLT.fs:351-365): the
    # reference's MLT images carry NaN pixels; Film.Save turns them black.  Restated, not repaired.
    finite = np.isfinite(film).all(axis=1)
    assert st["B"] > 0 and finite.mean() > 0.9
    assert abs(film[finite].mean() / pt[finite].mean() - 1) < 0.15, (film[finite].mean(), pt[finite].mean())
    assert per_chain.sum() == st["accepted"]


def test_oracle_pssmlt_chain_split_is_additive(lib):
    """Chains are independent given the bootstrap: rendering chain ranges separately and adding the
    films equals the full render (the multi-GPU sharding contract)."""
    set_portable_math(True)
    scene = load_scene("cbox_pt")
    o = OracleScene(scene.desc)
    full, st, _ = o.render_pssmlt(small_params(), threads=1)
    a, sa, _ = o.render_pssmlt(small_params(chain_begin=0, chain_end=40), threads=1)
    b, sb, _ = o.render_pssmlt(small_params(chain_begin=40, chain_end=96), threads=1)
    assert sa["accepted"] + sb["accepted"] == st["accepted"] and sa["B_bits"] == st["B_bits"]
    np.testing.assert_allclose(a + b, full, rtol=1e-5, atol=1e-6, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["BN_MLT_WAVEFRONT", "BN_MLT_MEGAKERNEL"])
@pytest.mark.parametrize("strategy", ["Gaussian", "Kelemen"])
@pytest.mark.parametrize("name", ["cbox_pt", "cbox_bunny"])
def test_gpu_pssmlt_matches_oracle(name, strategy, form, monkeypatch):
    """Both forms of the device integrator — a wavefront over chains through the path tracer's traversal kernels (what large
    chain counts and the bootstrap run) and one chain per thread (what small chain counts run) — evolve every chain as the
    oracle does: same bootstrap weights, B, accepted count per chain, ray count."""
    monkeypatch.setenv(form, "1")
    set_portable_math(True)
    scene = load_scene(name)
    g, o = scene.gpu(), OracleScene(scene.desc)
    p = small_params(strategy)
    # phase 1: bootstrap weights bit-exact
    wg, wo = g.pssmlt_bootstrap(p), o.pssmlt_bootstrap(p)
    assert np.array_equal(wg.view(np.uint32), wo.view(np.uint32)) or (np.isnan(wg) == np.isnan(wo)).all() and np.array_equal(wg[~np.isnan(wg)], wo[~np.isnan(wo)])
    # whole render: B, accepted counts per chain, film
    of, ost, o_chain = o.render_pssmlt(p)
    raw = C.CDLL(_ffi.LIB_PATH)
    film = np.zeros((H * W, 3), np.float32)
    st = _ffi.BnMltStats()
    g_chain = np.zeros(96, np.uint32)
    rc = raw.bn_debug_render_pssmlt_chains(g._h, C.byref(p), C.c_void_p(film.ctypes.data), C.byref(st), C.c_void_p(g_chain.ctypes.data))
    assert rc == 0
    assert np.float32(st.b).view(np.uint32) == np.uint32(ost["B_bits"])
    assert np.array_equal(g_chain, o_chain), f"{(g_chain != o_chain).sum()} of 96 chains diverged"
    assert (st.accepted, st.proposed, st.rays) == (ost["accepted"], ost["proposed"], ost["rays"])
    np.testing.assert_allclose(film, of, rtol=2e-5, atol=2e-6, equal_nan=True)      # atomic splat order only; same NaN pixels
    gf, st2 = g.render_pssmlt(p)
    np.testing.assert_allclose(gf, of, rtol=2e-5, atol=2e-6, equal_nan=True)
    assert st2.accepted == ost["accepted"]


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["BN_MLT_WAVEFRONT", "BN_MLT_MEGAKERNEL"])
def test_gpu_pssmlt_chain_shards_add_up(form, monkeypatch):
    monkeypatch.setenv(form, "1")
    set_portable_math(True)
    scene = load_scene("cbox_pt")
    g = scene.gpu()
    full, st = g.render_pssmlt(small_params())
    a, sa = g.render_pssmlt(small_params(chain_begin=0, chain_end=33))
    b, sb = g.render_pssmlt(small_params(chain_begin=33, chain_end=96))
    assert sa.accepted + sb.accepted == st.accepted and sa.b == st.b
    np.testing.assert_allclose(a + b, full, rtol=2e-5, atol=2e-6, equal_nan=True)
