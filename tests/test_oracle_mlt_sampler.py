"""Function-level pin of the oracle's MLTSampler ("next" row N1, PSSMLT.fs:35-150): the lazily mutated primary-sample
vector — start of an iteration with its large-step draw, first touch of a dimension, catch-up after a large step,
back-up / restore on rejection, Gaussian (ErfInv) and Kelemen mutations with their replay of skipped iterations —
against a second restatement written independently in Python (float32 numpy), driven by the same random scripts of
StartIteration / Next1D / Accept / Reject.  A mistake in the bookkeeping (which dimension is re-drawn when) changes the
values completely, so agreement to 2e-6 on thousands of draws pins it; libm mode."""
import numpy as np
import pytest

from oracle import oracle_ffi
from test_oracle_camera import M32

F = np.float32


class Lcg:  # Sampler.Next1D over a given state (Hash.fs:30-32)
    def __init__(self, state):
        self.state = state

    def next1d(self):
        self.state = (0x00269EC3 + self.state * 0x000343FD) & M32
        return F(np.array([(self.state >> 9) | 0x3F800000], dtype=np.uint32).view(F)[0] - F(1))


def _fma(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def erf_inv(x):  # PSSMLT.fs:68-96
    x = min(max(F(x), F(-0.99999)), F(0.99999))
    w = F(-np.log(_fma(x, -x, F(1))))
    if w < 5:
        w = F(w - F(2.5))
        cs = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
    else:
        w = F(np.sqrt(w) - F(3))
        cs = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]
    p = F(cs[0])
    for c in cs[1:]:
        p = _fma(p, w, F(c))
    return F(p * x)


class PyMltSampler:
    def __init__(self, seed_state, large_step_prob, strategy, p0, p1, n_dims):
        self.inner = Lcg(seed_state)
        self.lsp, self.strategy, self.p0, self.p1 = F(large_step_prob), strategy, F(p0), F(p1)
        self.xs = [[F(0), F(0), 0, 0] for _ in range(n_dims)]  # Value, ValueBackUp, LastModification, ModificationBackUp
        self.large_step, self.last_large, self.cur, self.index, self.init = False, 0, 0, 0, 0

    def start_iteration(self):  # :57-60
        self.large_step = self.cur == 0 or self.inner.next1d() < self.lsp
        self.cur += 1
        self.index = 0

    def next1d(self):  # :62-138
        i = self.index
        self.index += 1
        x = self.xs[i]
        if self.init <= i:
            x[:] = [F(0), F(0), 0, 0]
            self.init = i + 1
        if x[2] < self.last_large:
            x[0], x[2] = self.inner.next1d(), self.last_large
        x[1], x[3] = x[0], x[2]
        if self.large_step:
            v = self.inner.next1d()
        elif self.strategy == 0:
            normal = F(np.sqrt(F(2)) * erf_inv(_fma(F(2), self.inner.next1d(), F(-1))))
            v = _fma(normal, F(self.p0 * np.sqrt(F(self.cur - x[2]))), x[0])
        else:
            v = x[0]
            a = F(np.log(F(self.p1 / self.p0)))
            for _ in range(x[2], self.cur):
                u1 = F(self.inner.next1d() - F(0.5))
                u2 = F(F(1) + F(2) * u1) if u1 < 0 else F(F(2) * u1)
                v = F(v + np.copysign(F(self.p1 * np.exp(F(-a * u2))), u1))
        x[0] = F(v - np.floor(v))
        x[2] = self.cur
        return x[0]

    def reject(self):  # :142-146
        for x in self.xs[:self.init]:
            x[0], x[2] = x[1], x[3]
        self.cur -= 1

    def accept(self):  # :148-150
        if self.large_step:
            self.last_large = self.cur


def _random_script(rng, n_iter, n_dims):
    """What a chain does: StartIteration, a path's worth of draws (a varying number of dimensions), Accept or Reject."""
    script = []
    for _ in range(n_iter):
        script.append(0)
        script += [1] * int(rng.integers(1, n_dims + 1))
        script.append(2 if rng.random() < 0.6 else 3)
    return script


@pytest.mark.parametrize("strategy,p0,p1", [(0, 1e-2, 0.0), (1, 1.0 / 1024, 1.0 / 16)])
@pytest.mark.parametrize("large_step_prob", [0.5, 0.1])
def test_mlt_sampler_matches_restatement(oracle_lib, strategy, p0, p1, large_step_prob):
    oracle_ffi.set_portable_math(False)
    rng = np.random.default_rng(7 + strategy)
    n_dims = 4 + 7 * 3
    total = 0
    for seed_state in (0x12345678, 0xDEADBEEF, 1):
        script = _random_script(rng, 120, n_dims)
        got = oracle_ffi.mlt_sampler_script(seed_state, large_step_prob, strategy, p0, p1, script, n_dims)
        m = PyMltSampler(seed_state, large_step_prob, strategy, p0, p1, n_dims)
        want = []
        for op in script:
            if op == 0:
                m.start_iteration()
            elif op == 1:
                want.append(m.next1d())
            elif op == 2:
                m.accept()
            else:
                m.reject()
        want = np.array(want, dtype=F)
        assert ((got >= 0) & (got < 1)).all()
        # values live on the unit circle (x - floor x): compare modulo 1
        diff = np.abs(got - want)
        diff = np.minimum(diff, 1 - diff)
        assert diff.max() <= 2e-6, (int(np.argmax(diff)), float(diff.max()))
        total += len(want)
    assert total > 3000
