"""The only outputs the reference publishes are its two sample renders of Asset/cbox.json
(`Sample - Path Tracing.png`, `Sample - Primary Sample Space Metropolis Light Transport.png`:
1024x768, thin lens, ACES, RGBA8).  tests/golden/ref_sample_blocks16.npz holds the mean 8-bit RGB of
every 16x16 block of both (tests/golden/make_ref_sample_fixture.py); scenes/cbox_ref.json is
Asset/cbox.json as published.

These tests render that scene — oracle in libm mode on the CPU, CUDA path on the GPU — push the
film through Film.Save's tone-map (Film.fs:21-30,55-66) and compare block means with the F#
program's own picture.  It is a statistical pin (the images are 8-bit, of unknown sample count), but
it is a pin on the real reference: camera + thin lens, geometry, light power, Lambertian / dielectric
BSDFs, MIS weights, Russian roulette and ACES + gamma all have to be right for the block means to
agree to ~1 of 255 levels.

Two regions are masked, both measured:
  * the sphere — the published images show a visibly rougher sphere than `roughness: 0.4` of the
    committed cbox.json gives; the best fit is roughness ~0.55 (block error 2.2 levels vs 11.5 at
    0.4), and BOTH sample images agree with each other to 0.4 levels there, so they were rendered
    from an earlier revision of the scene file.  The sphere region is therefore checked with the
    fitted roughness, as a consistency check of PBRMaterial, not as a pin;
  * the two block rows that hold the light's upper / lower edge (a ~2 px edge shift moves those
    block means by ~10 levels; same explanation).
"""
import json
import os

import numpy as np
import pytest

from barnacle_b200.scene import Film, Scene, make_mlt_params, make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, B = 1024, 768, 16
SPHERE = (slice(432 // B, 688 // B), slice(224 // B, 480 // B))
LIGHT_EDGES = (slice(1, 6), slice(25, 39))
FITTED_SPHERE_ROUGHNESS = 0.55


def fixture():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_sample_blocks16.npz"))


def rest_mask():
    m = np.ones((H // B, W // B), bool)
    m[SPHERE] = False
    m[LIGHT_EDGES] = False
    return m


def load_ref_scene(spp, roughness=None, integrator="path-tracing"):
    with open(os.path.join(ROOT, "scenes", "cbox_ref.json")) as f:
        js = json.load(f)
    assert js["film"] == {"width": W, "height": H, "tone-mapping": "aces"}
    if integrator is not None:
        js["integrator"] = {"type": integrator, "spp": spp, "max-depth": js["integrator"].get("max-depth", 8)}
    if roughness is not None:
        for m in js["materials"]:
            if m["type"] == "pbr":
                m["roughness"] = roughness
    return Scene.LoadString(json.dumps(js), base_dir=ROOT)


PRE = 4  # linear pre-average, pixels


def blocks_of(film_pixels):
    """Film.Save's tone-map (ACES + gamma + clamp -> RGBA8) and then block means.  The published
    images are converged (pixel noise ~3 levels); ours are a few spp, and the tone curve is concave, so
    tone-mapping noisy pixels biases block means low (measured: -2.3 levels at 8 spp, -1.2 at 16, -0.5
    at 32).  The linear film is therefore averaged over 4x4 pixels first (16x the samples; the picture
    is smooth at that scale), tone-mapped, and then averaged per block.  NaN pixels (a handful, SURVEY
    Q15) are the reference's black pixels: 0."""
    lin = np.nan_to_num(film_pixels.reshape(H, W, 3).astype(np.float64), nan=0.0, posinf=1e30)
    lin = lin.reshape(H // PRE, PRE, W // PRE, PRE, 3).mean(axis=(1, 3))
    f = Film(W // PRE, H // PRE, "aces")
    f.Pixels[:] = lin.reshape(-1, 3).astype(np.float32)
    a = f.to_rgba8().astype(np.float64)[..., :3]
    k = B // PRE
    return a.reshape(H // B, k, W // B, k, 3).mean(axis=(1, 3))


def test_fixture_is_self_consistent():
    """The two published images are renders of the same scene by two integrators: their block means
    agree to < 1 level everywhere (0.37 measured), including the sphere."""
    fx = fixture()
    d = np.abs(fx["pt"].astype(np.float64) - fx["mlt"])
    assert d.mean() < 0.6 and d[SPHERE].mean() < 0.8
    assert 0 < int(fx["pt_black"]) < 100 and 0 < int(fx["mlt_black"]) < 100  # the reference's NaN pixels


def test_oracle_matches_published_path_tracing_image(lib):
    """Oracle (libm mode = what the reference calls), 8 spp, vs `Sample - Path Tracing.png`."""
    from oracle.oracle_ffi import OracleScene, set_portable_math
    set_portable_math(False)
    try:
        ref = fixture()["pt"].astype(np.float64)
        scene = load_ref_scene(8, FITTED_SPHERE_ROUGHNESS)
        film, _ = OracleScene(scene.desc).render(make_params(W, H, 8))
        d = blocks_of(film) - ref
        m = rest_mask()
        # measured: mean |d| 1.28 levels at 8 spp (0.81 at 32 spp), per-channel bias +0.15 .. +0.45,
        # 98.3 % of the blocks within 6 levels; sphere 3.1 with the fitted roughness (9.9 with 0.4)
        assert np.abs(d[m]).mean() < 2.0, np.abs(d[m]).mean()
        assert np.abs(d[m].mean(axis=0)).max() < 1.0, d[m].mean(axis=0)
        assert (np.abs(d[m]).max(axis=1) < 6).mean() > 0.96
        assert np.abs(d[SPHERE]).mean() < 4.5, np.abs(d[SPHERE]).mean()
    finally:
        set_portable_math(True)


@pytest.mark.gpu
def test_gpu_matches_published_path_tracing_image():
    """CUDA path, 64 spp through bn_render, vs `Sample - Path Tracing.png`; and the committed
    roughness (0.4) must fit the published sphere clearly worse than the fitted one — i.e. the mask
    is justified by the picture, not by us."""
    ref = fixture()["pt"].astype(np.float64)
    m = rest_mask()
    err = {}
    for rough in (FITTED_SPHERE_ROUGHNESS, None):
        scene = load_ref_scene(64, rough)
        film, st = scene.gpu().render(make_params(W, H, 64))
        d = blocks_of(film) - ref
        err[rough] = np.abs(d[SPHERE]).mean()
        assert np.abs(d[m]).mean() < 1.1, np.abs(d[m]).mean()          # oracle at 32 spp: 0.81
        assert np.abs(d[m].mean(axis=0)).max() < 1.5, d[m].mean(axis=0)  # B200 at 64 spp: +0.95 / +0.80 / +0.80 (oracle at 32 spp: +0.43)
        assert (np.abs(d[m]).max(axis=1) < 6).mean() > 0.99            # oracle at 32 spp: 0.998
        scene.close()
    assert err[FITTED_SPHERE_ROUGHNESS] < 4.0 and err[None] > 2 * err[FITTED_SPHERE_ROUGHNESS], err  # oracle: 3.0 vs 9.9


@pytest.mark.gpu
def test_gpu_pssmlt_matches_published_mlt_image():
    """cbox.json as published runs PSSMLT at 16 mutations per pixel; bn_render_pssmlt (4096 chains)
    against `Sample - Primary Sample Space Metropolis Light Transport.png`.  MLT is noisier and its
    normalisation B comes from the bootstrap, so the tolerance is looser."""
    ref = fixture()["mlt"].astype(np.float64)
    scene = load_ref_scene(16, FITTED_SPHERE_ROUGHNESS, integrator="pssmlt")
    film, st = scene.gpu().render_pssmlt(make_mlt_params(W, H, 16, n_bootstrap=1 << 20, n_chains=4096))
    d = blocks_of(film) - ref
    m = rest_mask()
    assert st.b > 0
    # oracle, same parameters: mean |d| 2.5 levels, per-channel bias +1.9 (B from 1 Mi instead of 4 Mi bootstrap paths)
    assert np.abs(d[m]).mean() < 4.0, np.abs(d[m]).mean()
    assert np.abs(d[m].mean(axis=0)).max() < 3.5, d[m].mean(axis=0)
    scene.close()
