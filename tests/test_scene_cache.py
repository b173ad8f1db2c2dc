"""bn_scene_create keeps the last flattened scene (kernels.cu: stage_scene) and reuses it when a new description's arrays are
byte-identical — a host that re-creates its scene every frame from unchanged geometry (the F# binding) skips the conversion,
not the upload.  The cache must hit on identical input, miss on ANY change of the arrays, and never cover the camera."""
import json

import numpy as np
import pytest

import emitter_scenes
from barnacle_b200.scene import Scene, make_params
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


def _render_both(text, p):
    scene = Scene.LoadString(text, base_dir=emitter_scenes.ROOT)
    gpu_film, _ = scene.gpu().render(p)
    want, _ = OracleScene(scene.desc).render(p)
    scene.close()
    return gpu_film, want


@pytest.mark.parametrize("cache", [True, False])
def test_cache_hits_only_on_identical_arrays_and_never_covers_the_camera(cache, monkeypatch):
    if not cache:
        monkeypatch.setenv("BN_NO_SCENE_CACHE", "1")
    oracle_ffi.set_portable_math(True)
    p = make_params(64, 64, 2)
    base = emitter_scenes.sphere_emitter()
    moved = json.loads(base)
    moved["transforms"][-1]["keyframes"][0]["translation"][0] += 7.0      # the emitter sphere elsewhere: other instance arrays
    tweaked = json.loads(base)
    tweaked["materials"][0]["albedo"] = [0.7, 0.75, 0.75]                   # one float of one array
    zoomed = json.loads(base)
    zoomed["camera"]["fov"] = 31.0                                          # same arrays, another camera
    films = {}
    for name, text in (("base", base), ("again", base), ("moved", json.dumps(moved)), ("base3", base), ("tweaked", json.dumps(tweaked)),
                       ("zoomed", json.dumps(zoomed)), ("base4", base)):
        got, want = _render_both(text, p)
        assert _bits_equal(got, want), name                                # whatever the cache did, the film is this scene's
        films[name] = got
    assert _bits_equal(films["base"], films["again"]) and _bits_equal(films["base"], films["base3"]) and _bits_equal(films["base"], films["base4"])
    for other in ("moved", "tweaked", "zoomed"):
        assert not _bits_equal(films["base"], films[other]), other
