"""Pins the control flow of the oracle's PSSMLTIntegrator ("next" row N1): PSSMLT.fs:172-245 (Li with its fixed seven
dimensions per bounce), :247-273 (bootstrap), :275-377 (a Markov chain: proposal, acceptance, the two splat weights,
Accept / Reject) and :379-414 (B, the uniform-by-Q1 bootstrap pick) restated in Python on top of building blocks that
have their own pins (MLTSampler: test_oracle_mlt_sampler.py; camera, light sampler, materials, traversal: the other
tests/test_oracle_*.py), and compared with the oracle's bootstrap weights, B, per-chain acceptance counts and film."""
import numpy as np
import pytest

from barnacle_b200 import _ffi
from barnacle_b200.scene import make_mlt_params
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene
from test_oracle_camera import M32, xxhash32_three
from test_oracle_li import F, _dot, _fma3, _ray
from test_oracle_mlt_sampler import Lcg, PyMltSampler, _fma


def xxhash32_two(x, y):  # Hash.fs:6-15
    p2, p3, p4, p5 = 2246822519, 3266489917, 668265263, 374761393
    h = (y + p5 + x * p3) & M32
    h = (p4 * (((h << 17) | (h >> 15)) & M32)) & M32
    h = (p2 * (h ^ (h >> 15))) & M32
    h = (p3 * (h ^ (h >> 13))) & M32
    return h ^ (h >> 16)


def mlt_li(oracle, desc, ray, m, max_depth, rr_depth):  # PSSMLT.fs:172-245
    o, d = ray["origin"][0].astype(F), ray["direction"][0].astype(F)
    L, beta = np.zeros(3, F), np.ones(3, F)
    depth, bsdf_pdf = 0, F(0)
    while depth < max_depth:
        r = _ray(o, d, np.inf)
        hit = oracle.trace(r)[0]
        if hit["instance"] < 0:
            break
        _, g = oracle.closest_geom(r)
        p, n, t, b = (g[k].astype(F) for k in range(4))
        inst = desc.instances[int(hit["instance"])]
        to_local = lambda v: np.array([_dot(v, t), _dot(v, b), _dot(v, n)], F)
        to_world = lambda v: (F(v[0]) * t + F(v[1]) * b) + F(v[2]) * n
        if inst.light_id >= 0:
            le = oracle.light_eval_hit(r)
            mis = F(1) if depth == 0 else F(bsdf_pdf * (F(1) / F(le[3] + bsdf_pdf)))
            L = _fma3(beta, le[:3] * mis, L)
        u_light, u_emit = m.next1d(), (m.next1d(), m.next1d())         # seven dimensions on EVERY hit (:201-205)
        u_lobe, u_bsdf, u_rr = m.next1d(), (m.next1d(), m.next1d()), m.next1d()
        if inst.material_id < 0:
            break
        mat = desc.materials[inst.material_id]
        ls = oracle.light_sample(p, u_light, u_emit)
        diff = ls[0:3] - p
        dist = F(np.sqrt(_dot(diff, diff)))
        wo_local = to_local(-d)
        if ls[6] != 0 and oracle.trace(_ray(p, ls[7:10], F(dist - F(1e-3))), any_hit=True)[0]["instance"] == 0:
            e = oracle_ffi.material_eval(mat, wo_local, to_local(ls[7:10]))
            L = _fma3(beta * e[:3], ls[3:6] * F(F(1) / F(e[3] + ls[6])), L)
        bs = oracle_ffi.material_sample(mat, wo_local, u_lobe, u_bsdf)
        bsdf_pdf = bs[3]
        if bsdf_pdf == 0:
            break
        o, d = p, to_world(bs[4:7])
        beta = (beta * bs[:3]) * F(F(1) / bsdf_pdf)
        if depth >= rr_depth:
            thr = min(F(1), max(beta[0], max(beta[1], beta[2])))
            if u_rr < thr:
                beta = beta * F(F(1) / thr)
            else:
                break
        depth += 1
    return L


def sample_path(oracle, desc, prm, m):  # the head shared by :247-273, :284-300, :332-349
    w, h = prm.width, prm.height
    ux, uy = F(m.next1d() * F(w)), F(m.next1d() * F(h))
    px, py = min(w - 1, int(ux)), min(h - 1, int(uy))
    u_lens = (m.next1d(), m.next1d())
    ray = oracle.camera_ray(w, h, px, py, (F(ux - F(px)), F(uy - F(py))), u_lens)
    return mlt_li(oracle, desc, ray, m, prm.max_depth, prm.rr_depth), px, py


def luminance(L):
    return _dot(L, np.array([0.2126, 0.7152, 0.0722], F))


def new_sampler(prm, seed_state):
    return PyMltSampler(seed_state, prm.large_step_prob, prm.strategy, prm.p0, prm.p1, 4 + 7 * prm.max_depth)


@pytest.mark.parametrize("strategy", ["Gaussian", "Kelemen"])
def test_pssmlt_bootstrap_and_chains_match_restatement(scene_loader, oracle_lib, strategy):
    oracle_ffi.set_portable_math(False)
    scene = scene_loader("cbox_pt")
    desc = scene.desc.contents
    oracle = OracleScene(scene.desc)
    w, h = 12, 10
    prm = make_mlt_params(w, h, 1, max_depth=4, rr_depth=2, n_bootstrap=48, n_chains=3, strategy=strategy)
    # ---- phase 1: BootstrapWeights, B (:382-394)
    want_w = oracle.pssmlt_bootstrap(prm, threads=1)
    got_w = np.zeros(prm.n_bootstrap, F)
    for k in range(prm.n_bootstrap):
        m = new_sampler(prm, xxhash32_two(prm.frame_id, k))
        m.start_iteration()
        L, _, _ = sample_path(oracle, desc, prm, m)
        got_w[k] = luminance(L)
    np.testing.assert_allclose(got_w, want_w, rtol=1e-4, atol=1e-7)
    assert (want_w > 0).sum() > 10
    total = F(0)
    for x in want_w:
        total = F(total + x)
    B = F(total / F(prm.n_bootstrap))                                   # Array.average: sequential fp32
    # ---- phase 2: the chains (:275-377)
    want_film, st, want_acc = oracle.render_pssmlt(prm, threads=1)
    assert st["B_bits"] == int(np.array([B]).view(np.uint32)[0])
    per_chain = (prm.mutations_per_pixel * w * h + prm.n_chains - 1) // prm.n_chains
    inv_eff = F(F(1) / F(F(F(per_chain) * F(prm.n_chains)) / F(w * h)))
    inv_b = F(F(1) / B)
    film = np.zeros((h * w, 3), F)

    def splat(px, py, c):  # Film.Accumulate (Film.fs:48-53)
        film[(h - py - 1) * w + px] += c.astype(F)

    got_acc = []
    for chain in range(prm.n_chains):
        sampler = Lcg(xxhash32_two(prm.frame_id, chain))
        bootstrap_id = min(int(F(sampler.next1d() * F(prm.n_bootstrap))), prm.n_bootstrap - 1)   # alias table without aliases
        m = new_sampler(prm, xxhash32_two(prm.frame_id, bootstrap_id))
        m.start_iteration()
        L, px, py = sample_path(oracle, desc, prm, m)
        y = luminance(L)
        m.accept()
        m.inner = Lcg(xxhash32_three(chain, bootstrap_id, prm.frame_id))
        radiance, acc = np.zeros(3, F), 0
        for _ in range(per_chain):
            m.start_iteration()
            Ln, nx, ny = sample_path(oracle, desc, prm, m)
            yn = luminance(Ln)
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = F(yn / y)
                a = ratio if np.isnan(ratio) else min(F(1), ratio)       # MathF.Min propagates NaN (0 / 0 chains, :351)
                w_old = F(F(F(1) - a) / _fma(y, inv_b, F(prm.large_step_prob)))
                radiance = radiance + w_old * L
                w_new = F(F(a + (F(1) if m.large_step else F(0))) / _fma(yn, inv_b, F(prm.large_step_prob)))
            if sampler.next1d() < a:
                acc += 1
                splat(px, py, radiance * inv_eff)
                radiance = w_new * Ln
                px, py, L, y = nx, ny, Ln, yn
                m.accept()
            else:
                if a > 0:
                    splat(nx, ny, F(w_new * inv_eff) * Ln)
                m.reject()
        splat(px, py, radiance * inv_eff)
        got_acc.append(acc)
    assert got_acc == want_acc.tolist()
    assert sum(got_acc) > 0.2 * per_chain * prm.n_chains
    np.testing.assert_allclose(film, want_film, rtol=2e-4, atol=1e-6, equal_nan=True)
