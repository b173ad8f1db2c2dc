"""Function-level pin of the oracle's materials.

The GPU shade kernel is compared bit for bit with the oracle, so a transcription error in the oracle's
BSDF code would go unnoticed by every parity test.  Here `material_eval` / `material_sample` of
oracle/barnacle_oracle.cpp are compared with a second restatement of Lambertian.fs, Mirror.fs,
Dielectric.fs and PBR.fs, written independently in float64 numpy from the F# text, on random inputs —
plus the physical identities the formulas must satisfy (cosine pdf, Snell's law, |wi| = 1, the GGX
half-vector pdf integrating to the lobe's weight).  libm mode (what the reference calls); fp32 vs fp64
differences bound the tolerances.
"""
import math

import numpy as np
import pytest

from barnacle_b200 import _ffi
from oracle import oracle_ffi


@pytest.fixture(autouse=True)
def _libm_mode(oracle_lib):
    oracle_ffi.set_portable_math(False)
    yield
    oracle_ffi.set_portable_math(False)


def _material(kind, base, p0=0.0, p1=0.0):
    m = _ffi.BnMaterial()
    m.type = kind
    m.base_color[:] = base
    m.p0, m.p1 = p0, p1
    return m


def _unit(rng, n, upper=None):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if upper is True:
        v[:, 2] = np.abs(v[:, 2])
    return v


# ---- independent float64 restatement of the F# ----------------------------------------------------------

def ref_lambert_eval(base, wo, wi):  # Lambertian.fs:11-16
    if wi[2] * wo[2] < 0 or min(abs(wi[2]), abs(wo[2])) < 1e-6:
        return np.zeros(3), 0.0
    pdf = abs(wi[2]) / math.pi
    return base * pdf, pdf


def ref_cosine(u):  # Lambertian.fs:19-22
    ct, st = math.sqrt(u[0]), math.sqrt(1 - u[0])
    phi = 2 * math.pi * u[1]
    return np.array([st * math.cos(phi), st * math.sin(phi), ct])


def ref_lambert_sample(base, wo, u):  # Lambertian.fs:18-26
    wi = ref_cosine(u)
    pdf = wi[2] / math.pi
    return base * pdf, pdf, (wi if wo[2] > 0 else -wi)


def ref_dielectric_sample(base, ior, wo, ulobe):  # Dielectric.fs:15-31
    eta = 1 / ior if wo[2] > 0 else ior
    cos2 = 1 - eta * eta * (1 - wo[2] * wo[2])
    refl = np.array([-wo[0], -wo[1], wo[2]])
    if cos2 <= 0:
        return base, 1.0, refl
    r0 = ((eta - 1) / (eta + 1)) ** 2
    c = 1 - abs(wo[2])
    r = r0 + (1 - r0) * c ** 5
    if ulobe < r:
        return r * base, r, refl
    return (1 - r) * base, 1 - r, np.array([-wo[0] * eta, -wo[1] * eta, -math.copysign(math.sqrt(cos2), wo[2])])


def ref_pbr_eval(base, metallic, alpha, wo, wi):  # PBR.fs:13-48
    def lam(w):
        s2 = w[0] * w[0] + w[1] * w[1]
        return 0.0 if s2 == 0 else (-1 + math.sqrt(1 + alpha * alpha * s2 / (w[2] * w[2]))) / 2
    wh = (wo + wi) / np.linalg.norm(wo + wi)
    d = alpha * alpha / (math.pi * (1 + (alpha * alpha - 1) * wh[2] ** 2) ** 2)
    g = 1 / (1 + lam(wo) + lam(wi))
    spec = d * g / (4 * abs(wo[2]))
    c5 = (1 - float(wo @ wh)) ** 5
    metal = spec * (base + (1 - base) * c5)
    f = 0.04 + 0.96 * c5
    diffuse = base * abs(wi[2]) / math.pi
    mix = lambda a, b, t: a * (1 - t) + b * t
    bsdf = mix(mix(diffuse, np.full(3, spec), f), metal, metallic)
    pdf = mix(d * abs(wh[2]) / (4 * float(wo @ wh)), abs(wi[2]) / math.pi, 0.5 * (1 - metallic))
    return bsdf, pdf


def ref_pbr_sample_dir(metallic, alpha, wo, ulobe, u):  # PBR.fs:49-63
    if ulobe < 1 - 0.5 * (1 - metallic):
        th = math.atan(alpha * math.sqrt(u[0] / (1 - u[0])))
        ph = 2 * math.pi * u[1]
        wh = np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
        return 2 * float(wo @ wh) * wh - wo
    return ref_cosine(u)


# ---- tests -------------------------------------------------------------------------------------------

def test_lambertian_matches_restatement_and_is_cosine_weighted():
    rng = np.random.default_rng(1)
    base = np.array([0.7, 0.4, 0.2])
    m = _material(_ffi.BN_MAT_LAMBERTIAN, base)
    for wo, wi, u in zip(_unit(rng, 400), _unit(rng, 400), rng.random((400, 2))):
        e = oracle_ffi.material_eval(m, wo, wi)
        b, p = ref_lambert_eval(base, wo.astype(np.float32), wi.astype(np.float32))
        np.testing.assert_allclose(e[:3], b, rtol=2e-6, atol=1e-9)
        assert e[3] == pytest.approx(p, rel=2e-6, abs=1e-9)
        s = oracle_ffi.material_sample(m, wo, 0.5, u)
        b, p, w = ref_lambert_sample(base, wo, u.astype(np.float32))
        np.testing.assert_allclose(s[4:], w, atol=3e-6)
        np.testing.assert_allclose(s[:3], b, rtol=1e-5, atol=1e-7)
        assert s[3] == pytest.approx(p, rel=1e-5, abs=1e-7)
        assert np.sign(s[6]) == np.sign(wo[2])                      # same side as wo
        assert abs(np.linalg.norm(s[4:]) - 1) < 1e-6
    # grazing and opposite-side configurations evaluate to zero (Lambertian.fs:12-13)
    assert not oracle_ffi.material_eval(m, [0, 0.6, 0.8], [0, 0.6, -0.8]).any()
    assert not oracle_ffi.material_eval(m, [0, 1, 5e-7], [0, 0.6, 0.8]).any()


def test_mirror_and_dielectric_match_restatement_and_snell():
    rng = np.random.default_rng(2)
    base = np.array([0.9, 0.8, 0.95])
    mirror = _material(_ffi.BN_MAT_MIRROR, base)
    for ior in (1.1, 1.5, 2.0):
        glass = _material(_ffi.BN_MAT_DIELECTRIC, base, p0=ior)
        n_refr = n_tir = 0
        for wo, ul in zip(_unit(rng, 500), rng.random(500)):
            wo32 = wo.astype(np.float32)
            s = oracle_ffi.material_sample(mirror, wo, ul, [0.3, 0.7])
            np.testing.assert_array_equal(s, np.concatenate([base.astype(np.float32), [1.0], [-wo32[0], -wo32[1], wo32[2]]]))
            assert not oracle_ffi.material_eval(mirror, wo, -wo).any()   # Mirror.fs:11 / Dielectric.fs:13: Eval is zero
            assert not oracle_ffi.material_eval(glass, wo, -wo).any()
            s = oracle_ffi.material_sample(glass, wo, ul, [0.3, 0.7])
            b, p, w = ref_dielectric_sample(base, np.float32(ior), wo32.astype(np.float64), ul)
            # the reflect / refract choice can flip when ulobe sits within rounding of r: skip those (none in practice)
            if abs(p - s[3]) > 1e-4:
                continue
            np.testing.assert_allclose(s[:3], b, rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose(s[4:], w, rtol=2e-5, atol=2e-6)
            if np.sign(s[6]) != np.sign(wo[2]):                      # refracted: Snell with eta' and the far side
                eta = 1 / ior if wo[2] > 0 else ior
                assert math.hypot(s[4], s[5]) == pytest.approx(eta * math.hypot(wo[0], wo[1]), rel=1e-5, abs=1e-6)
                assert abs(np.linalg.norm(s[4:]) - 1) < 1e-5
                n_refr += 1
            elif s[3] == 1.0:
                n_tir += 1
        assert n_refr > 100
        assert n_tir > 0                                                # total internal reflection shows up from inside


@pytest.mark.parametrize("metallic,roughness", [(0.0, 0.4), (0.5, 0.2), (1.0, 0.05), (1.0, 1.0), (0.0, 0.01)])
def test_pbr_matches_restatement(metallic, roughness):
    rng = np.random.default_rng(3)
    base = np.array([0.8, 0.5, 0.3])
    alpha = max(np.float32(roughness) * np.float32(roughness), np.float32(1e-3))  # PBRMaterial.Alpha (PBR.fs:11)
    m = _material(_ffi.BN_MAT_PBR, base, p0=metallic, p1=float(alpha))
    checked = 0
    for wo, wi, ul, u in zip(_unit(rng, 600, upper=True), _unit(rng, 600, upper=True), rng.random(600), rng.random((600, 2))):
        wo32, wi32 = wo.astype(np.float32).astype(np.float64), wi.astype(np.float32).astype(np.float64)
        if min(wo32[2], wi32[2]) < 0.05:
            continue                                                    # grazing: fp32 cancellation in tan^2, not a formula check
        e = oracle_ffi.material_eval(m, wo, wi)
        b, p = ref_pbr_eval(base, metallic, float(alpha), wo32, wi32)
        np.testing.assert_allclose(e[:3], b, rtol=2e-4, atol=1e-7)
        assert e[3] == pytest.approx(p, rel=2e-4, abs=1e-7)
        s = oracle_ffi.material_sample(m, wo, ul, u)
        w = ref_pbr_sample_dir(metallic, float(alpha), wo32, ul, u.astype(np.float32).astype(np.float64))
        np.testing.assert_allclose(s[4:], w, rtol=1e-3, atol=2e-4)     # atan / sincos of small angles in fp32
        # Sample returns Eval at the sampled direction (PBR.fs:64)
        np.testing.assert_array_equal(s[:4], oracle_ffi.material_eval(m, wo, s[4:]))
        checked += 1
    assert checked > 300


def test_pbr_pdf_is_a_density():
    """The mixture pdf (GGX half-vector lobe + cosine lobe) is the density of what Sample generates: over the upper
    hemisphere it integrates to 1 minus the part of the specular lobe that falls below the horizon (the reference does
    not reject those directions: SURVEY Q15), and over the whole sphere — where the |wi.z| / pi term counts twice — to
    1 + the cosine lobe's weight.  Uniform sphere samples; pins D, |wh.z| / (4 wo.wh) and the lobe weights (PBR.fs:18-21,46)."""
    rng = np.random.default_rng(4)
    wo = np.array([0.3, -0.2, 0.0])
    wo[2] = math.sqrt(1 - wo[0] ** 2 - wo[1] ** 2)
    for metallic, roughness in ((0.0, 0.6), (1.0, 0.5), (0.5, 0.8)):
        m = _material(_ffi.BN_MAT_PBR, [0.8, 0.8, 0.8], p0=metallic, p1=roughness * roughness)
        wis = _unit(rng, 40000)
        pdf = np.array([oracle_ffi.material_eval(m, wo, wi)[3] for wi in wis], dtype=np.float64)
        assert np.isfinite(pdf).all() and (pdf >= 0).all()
        w_cos = 0.5 * (1 - metallic)
        whole = 4 * math.pi * pdf.mean()
        upper = 4 * math.pi * np.where(wis[:, 2] > 0, pdf, 0).mean()
        assert whole == pytest.approx(1 + w_cos, rel=0.04), (metallic, roughness, whole)
        assert 0.7 < upper < 1.03, (metallic, roughness, upper)  # rough lobes lose up to ~a quarter below the horizon


@pytest.mark.parametrize("metallic,roughness", [(0.0, 0.4), (1.0, 0.3), (0.5, 0.6), (0.0, 0.8)])
def test_pbr_directional_albedo_furnace(metallic, roughness):
    """White-furnace-style check of PBRMaterial that needs no rendered sphere (the published renders' sphere was made with another
    roughness): the directional albedo a(wo) = integral of Eval(wo, wi).bsdf over the UPPER hemisphere (the bsdf
    carries its cosine, PBR.fs:37-46), computed twice —
      (i)  by quadrature over (theta, phi) of the independent float64 restatement of PBR.fs above, and
      (ii) as the importance-sampling estimate mean(bsdf / pdf) over the oracle's own Sample draws that land in the upper
           hemisphere (what PathTracingIntegrator.Li multiplies the throughput by, PathTracing.fs:61-67).
    (ii) is unbiased for (i) only if Sample's directions are distributed with the density Eval reports AND Eval's bsdf is the
    restated formula: lobe selection, the half-vector pdf's Jacobian, D, G and the Fresnel mixes all enter.  A white base colour
    bounds the energy: a(wo) <= 1 up to the known excess of the un-normalised diffuse + specular mix, asserted as < 1.15."""
    rng = np.random.default_rng(11)
    base = np.array([1.0, 1.0, 1.0])
    alpha = float(max(np.float32(roughness) * np.float32(roughness), np.float32(1e-3)))
    m = _material(_ffi.BN_MAT_PBR, base, p0=metallic, p1=alpha)
    for cos_o in (0.95, 0.6, 0.3):
        wo = np.array([math.sqrt(1 - cos_o * cos_o), 0.0, cos_o])
        # (i) midpoint quadrature in (cos theta, phi), denser near the specular peak than a uniform grid needs: 400 x 720
        ct = (np.arange(400) + 0.5) / 400
        ph = (np.arange(720) + 0.5) / 720 * 2 * math.pi
        quad = 0.0
        for c in ct:
            s = math.sqrt(1 - c * c)
            wis = np.stack([s * np.cos(ph), s * np.sin(ph), np.full_like(ph, c)], axis=1)
            quad += sum(ref_pbr_eval(base, metallic, alpha, wo, wi)[0][0] for wi in wis[::8]) * 8
        quad *= (1 / 400) * (2 * math.pi / 720)
        # (ii) the oracle's Sample / Eval
        est, n = 0.0, 20000
        for ul, u in zip(rng.random(n), rng.random((n, 2))):
            sres = oracle_ffi.material_sample(m, wo, ul, u)
            if sres[6] > 0 and sres[3] > 0:  # upper hemisphere, pdf > 0
                est += sres[0] / sres[3]
        est /= n
        assert est == pytest.approx(quad, rel=0.05), (metallic, roughness, cos_o, est, quad)
        assert 0.0 < quad < 1.15, (metallic, roughness, cos_o, quad)
