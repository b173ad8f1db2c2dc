"""Function-level pin of the oracle's primary-ray generation: Sampler(x, y, sampleId) -> two
Next2D draws -> PinholeCamera / ThinLensCamera.GenerateRay -> CameraBase.GeneratePrimaryRay, against a second
restatement of Sampler.fs, Hash.fs:17-32, Pinhole.fs:12-27, ThinLens.fs:12-35 and Camera.fs:12-20 written
independently in Python / float64 numpy.  The integer part (hash, LCG, the 23-bit float) must agree exactly; the
float part to fp32 rounding.  CPU only, libm mode."""
import json
import math

import numpy as np
import pytest

from barnacle_b200.scene import Scene, make_params
from oracle import oracle_ffi
from oracle.oracle_ffi import OracleScene

M32 = 0xFFFFFFFF


def xxhash32_three(x, y, z):  # Hash.fs:17-28
    p2, p3, p4, p5 = 2246822519, 3266489917, 668265263, 374761393
    rot = lambda h: ((h << 17) | (h >> 15)) & M32
    h = (z + p5 + x * p3) & M32
    h = (p4 * rot(h)) & M32
    h = (h + y * p3) & M32
    h = (p4 * rot(h)) & M32
    h = (p2 * (h ^ (h >> 15))) & M32
    h = (p3 * (h ^ (h >> 13))) & M32
    return h ^ (h >> 16)


class PySampler:  # Sampler.fs:9-16
    def __init__(self, x, y, z):
        self.state = xxhash32_three(x, y, z)

    def next1d(self):  # Hash.LCG, Hash.fs:30-32
        self.state = (0x00269EC3 + self.state * 0x000343FD) & M32
        return float(np.array([(self.state >> 9) | 0x3F800000], dtype=np.uint32).view(np.float32)[0] - np.float32(1))

    def next2d(self):
        return self.next1d(), self.next1d()


def concentric_disk(u):  # ThinLens.fs:12-23
    ux, uy = 2 * u[0] - 1, 2 * u[1] - 1
    if ux == 0 or uy == 0:
        return np.zeros(2)
    if abs(ux) > abs(uy):
        r, th = ux, math.pi / 4 * (uy / ux)
    else:
        r, th = uy, math.pi / 2 - math.pi / 4 * (ux / uy)
    return r * np.array([math.cos(th), math.sin(th)])


def ref_primary_ray(cam, width, height, x, y, sample_id, frame_id=0, spp=1):
    s = PySampler(x, y, frame_id * spp + sample_id)                    # Integrator.fs:35-36
    u_pixel, u_lens = s.next2d(), s.next2d()                            # Integrator.fs:39
    vh = 2 * math.tan(float(np.float32(cam.fov_y)) * math.pi / 360)     # Pinhole.fs:15
    vw = vh * float(np.float32(cam.aspect_ratio))
    loc = np.array([-0.5 * vw + (x + u_pixel[0]) * vw / width, -0.5 * vh + (y + u_pixel[1]) * vh / height, -1.0])
    o, d = np.zeros(3), loc / np.linalg.norm(loc)
    if cam.type == 1 and cam.aperture > 0:                              # ThinLens.fs:25-35
        lens = float(cam.aperture) * concentric_disk(u_lens)
        o2 = o + np.array([lens[0], lens[1], 0.0])
        d = (o + float(cam.focus_distance) * d) - o2
        o, d = o2, d / np.linalg.norm(d)
    m = np.array(cam.camera_to_world[:], dtype=np.float64).reshape(4, 4)
    ow = o @ m[:3, :3] + m[3, :3]                                       # Camera.fs:14
    dw = d @ m[:3, :3]                                                  # Transform(dir) - Translation, Camera.fs:15-18
    dw /= np.linalg.norm(dw)
    return ow + float(cam.push_forward) * dw, dw                       # Camera.fs:19


def _scene(camera, transform):
    return json.dumps({"nodes": [{"children": [1, 2]}, {"instances": [0]}, {"has-camera": True, "transform": 0}],
                       "instances": [{"primitive": 0, "light": 0}], "transforms": [transform], "primitives": [{"type": "quad"}],
                       "materials": [], "lights": [{"type": "diffuse", "emission": [1, 1, 1]}],
                       "integrator": {"type": "path-tracing", "spp": 3}, "camera": camera,
                       "film": {"width": 24, "height": 16, "tone-mapping": "identity"}})


@pytest.mark.parametrize("camera", [
    {"type": "pinhole", "fov": 50.0, "aspect-ratio": 1.5},
    {"type": "pinhole", "fov": 35.0, "aspect-ratio": 1.5, "push-forward": 2.5},
    {"type": "thin-lens", "fov": 40.0, "aspect-ratio": 1.5, "aperture": 0.75, "focus-distance": 7.0, "push-forward": 1.25},
    {"type": "thin-lens", "fov": 40.0, "aspect-ratio": 1.5, "aperture": 0.0, "focus-distance": 7.0},
])
def test_primary_rays_match_restatement(lib, oracle_lib, camera):
    oracle_ffi.set_portable_math(False)
    transform = {"keyframes": [{"translation": [3.0, -2.0, 9.0], "rotation": [0.3, -0.8, 0.15], "scale": [1.0, 1.0, 1.0]}]}
    scene = Scene.LoadString(_scene(camera, transform))
    cam = scene.desc.contents.camera
    w, h, spp = 24, 16, 3
    p = make_params(w, h, spp)
    rays = OracleScene(scene.desc).primary_rays(p).reshape(spp, h, w)
    assert np.all(np.isinf(rays["tmax"]))                               # Li starts at t = +inf (PathTracing.fs:25)
    for s in range(spp):
        for y in range(0, h, 3):
            for x in range(0, w, 5):
                o, d = ref_primary_ray(cam, w, h, x, y, s, spp=spp)
                np.testing.assert_allclose(rays[s, y, x]["origin"], o, rtol=2e-5, atol=2e-5)
                np.testing.assert_allclose(rays[s, y, x]["direction"], d, rtol=2e-5, atol=2e-6)


def test_cbox_camera_matches_restatement(scene_loader, oracle_lib):
    """The published scene's own camera (thin lens, aperture 4, focus 210, push-forward 140, rotated node)."""
    oracle_ffi.set_portable_math(False)
    scene = scene_loader("cbox_pt")
    cam = scene.desc.contents.camera
    assert cam.type == 1 and cam.aperture > 0
    w = h = 32
    rays = OracleScene(scene.desc).primary_rays(make_params(w, h, 2)).reshape(2, h, w)
    for s in range(2):
        for y in range(0, h, 5):
            for x in range(0, w, 7):
                o, d = ref_primary_ray(cam, w, h, x, y, s, spp=2)
                np.testing.assert_allclose(rays[s, y, x]["origin"], o, rtol=2e-5, atol=5e-4)   # coordinates of order 10^2..10^3
                np.testing.assert_allclose(rays[s, y, x]["direction"], d, rtol=2e-5, atol=2e-6)
