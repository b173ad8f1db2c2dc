"""Keyframe interpolation on the host side: Transform.Eval / KeyFrame.Interpolate
(Base/Scene.fs:9-36) with Transform.Decompose / Compose (Util/Transform.fs) as the C++ stand-in for
the managed host evaluates them (barnacle_b200/csrc/host/host_math.hpp), cross-checked against an
independent numpy restatement of the same BCL algorithms written here.  The BCL itself is not in the
image: rounding against a real .NET host is unpinned; what is pinned is the algorithm, the reference's
keyframe selection rule and its crossed Compose arguments."""
import json

import numpy as np
import pytest

from barnacle_b200.scene import Scene

F = np.float32


def rot_xyz(r):
    def rx(a):
        c, s = np.cos(F(a)), np.sin(F(a))
        return np.array([[1, 0, 0, 0], [0, c, s, 0], [0, -s, c, 0], [0, 0, 0, 1]], F)

    def ry(a):
        c, s = np.cos(F(a)), np.sin(F(a))
        return np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], F)

    def rz(a):
        c, s = np.cos(F(a)), np.sin(F(a))
        return np.array([[c, s, 0, 0], [-s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], F)
    return rx(r[0]) @ ry(r[1]) @ rz(r[2])


def keyframe_matrix(kf):  # Loader.fs:20-41: S * R * T, row-vector convention
    S = np.diag(np.array(list(kf.get("scale", [1, 1, 1])) + [1], F))
    R = rot_xyz(kf["rotation"]) if "rotation" in kf else np.eye(4, dtype=F)
    T = np.eye(4, dtype=F)
    T[3, :3] = kf.get("translation", [0, 0, 0])
    return (S @ R @ T).astype(F)


def np_decompose(M):
    """Matrix4x4.Decompose, the generic (non-degenerate, right- or left-handed) branch."""
    tr = M[3, :3].copy()
    basis = [M[i, :3].astype(F).copy() for i in range(3)]
    sc = [F(np.sqrt(F(np.dot(b, b)))) for b in basis]
    order = sorted(range(3), key=lambda i: -sc[i])  # a, b, c by decreasing scale (ties never occur in the cases below)
    for i in order:
        basis[i] = basis[i] / F(np.linalg.norm(basis[i]))
    T = np.eye(4, dtype=F)
    for i in range(3):
        T[i, :3] = basis[i]
    det = np.linalg.det(T.astype(np.float64))
    if det < 0:
        a = order[0]
        sc[a] = -sc[a]
        T[a, :3] = -T[a, :3]
        det = -det
    if (det - 1.0) ** 2 > 1e-4:
        return np.array(sc, F), np.array([0, 0, 0, 1], F), tr
    m = T
    trace = m[0, 0] + m[1, 1] + m[2, 2]
    if trace > 0:
        s = F(np.sqrt(trace + F(1)))
        w = s * F(0.5)
        s = F(0.5) / s
        q = np.array([(m[1, 2] - m[2, 1]) * s, (m[2, 0] - m[0, 2]) * s, (m[0, 1] - m[1, 0]) * s, w], F)
    elif m[0, 0] >= m[1, 1] and m[0, 0] >= m[2, 2]:
        s = F(np.sqrt(F(1) + m[0, 0] - m[1, 1] - m[2, 2]))
        i = F(0.5) / s
        q = np.array([F(0.5) * s, (m[0, 1] + m[1, 0]) * i, (m[0, 2] + m[2, 0]) * i, (m[1, 2] - m[2, 1]) * i], F)
    elif m[1, 1] > m[2, 2]:
        s = F(np.sqrt(F(1) + m[1, 1] - m[0, 0] - m[2, 2]))
        i = F(0.5) / s
        q = np.array([(m[1, 0] + m[0, 1]) * i, F(0.5) * s, (m[2, 1] + m[1, 2]) * i, (m[2, 0] - m[0, 2]) * i], F)
    else:
        s = F(np.sqrt(F(1) + m[2, 2] - m[0, 0] - m[1, 1]))
        i = F(0.5) / s
        q = np.array([(m[2, 0] + m[0, 2]) * i, (m[2, 1] + m[1, 2]) * i, F(0.5) * s, (m[0, 1] - m[1, 0]) * i], F)
    return np.array(sc, F), q, tr


def np_slerp(q1, q2, t):
    c = float(np.dot(q1, q2))
    flip = c < 0
    c = abs(c)
    if c > 1 - 1e-6:
        s1, s2 = 1 - t, (-t if flip else t)
    else:
        om = np.arccos(c)
        s1 = np.sin((1 - t) * om) / np.sin(om)
        s2 = np.sin(t * om) / np.sin(om) * (-1 if flip else 1)
    return (F(s1) * q1 + F(s2) * q2).astype(F)


def np_from_quat(q):
    x, y, z, w = (float(v) for v in q)
    M = np.eye(4)
    M[0, :3] = [1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (z * x - y * w)]
    M[1, :3] = [2 * (x * y - z * w), 1 - 2 * (z * z + x * x), 2 * (y * z + x * w)]
    M[2, :3] = [2 * (z * x + y * w), 2 * (y * z - x * w), 1 - 2 * (y * y + x * x)]
    return M


def np_interpolate(ta, A, tb, B, t):
    ratio = (t - ta) / (tb - ta)
    s1, r1, t1 = np_decompose(A)
    s2, r2, t2 = np_decompose(B)
    tr = t1 * (1 - ratio) + t2 * ratio
    sc = s1 * (1 - ratio) + s2 * ratio
    rot = np_slerp(r1, r2, ratio)
    # Transform.Compose as written: CreateTranslation(scale) * CreateFromQuaternion(rot) * CreateScale(translation)
    T = np.eye(4)
    T[3, :3] = sc
    S = np.diag(list(tr) + [1.0])
    return T @ np_from_quat(rot) @ S


def scene_json(keyframes):
    return json.dumps({
        "nodes": [{"children": [1, 2]}, {"instances": [0], "transform": 0}, {"transform": 1, "has-camera": True}],
        "instances": [{"primitive": 0, "light": 0}],
        "transforms": [{"keyframes": keyframes}, {"keyframes": [{"translation": [0, 0, 5]}]}],
        "primitives": [{"type": "sphere", "radius": 1.0}],
        "materials": [], "lights": [{"type": "diffuse", "emission": [1, 1, 1]}],
        "integrator": {"type": "path-tracing", "spp": 1}, "camera": {"type": "pinhole"},
        "film": {"width": 8, "height": 8, "tone-mapping": "identity"}})


def object_to_world(keyframes, t):
    sc = Scene.LoadString(scene_json(keyframes), time_=t)
    m = np.array(sc.desc.contents.instances[0].object_to_world[:], F).reshape(4, 4)
    sc.close()
    return m


KF = [{"time": 0.0, "scale": [2, 3, 4], "rotation": [0.3, -0.5, 0.2], "translation": [1, 2, 3]},
      {"time": 2.0, "scale": [1.5, 2.5, 5], "rotation": [1.1, 0.4, -0.7], "translation": [-4, 6, 2]},
      {"time": 3.0, "scale": [1, 1, 1], "rotation": [0.0, 2.0, 0.0], "translation": [0, 1, 0]}]


def test_selection_rule(lib):
    """Transform.Eval (Scene.fs:26-36): before the first keyframe -> the first; at or after the last
    -> the last; a single keyframe -> that matrix at any time; keyframes are sorted by time."""
    A, C = keyframe_matrix(KF[0]), keyframe_matrix(KF[2])
    np.testing.assert_allclose(object_to_world(KF, -1.0), A, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(object_to_world(KF, 3.0), C, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(object_to_world(KF, 99.0), C, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(object_to_world(KF[::-1], 99.0), C, rtol=1e-6, atol=1e-6)
    for t in (-5.0, 0.0, 7.0):
        np.testing.assert_allclose(object_to_world(KF[:1], t), A, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("t", [0.5, 1.0, 1.9, 2.25, 2.75])
def test_interpolation_matches_numpy_restatement(lib, t):
    seg = (0, 1) if t < 2.0 else (1, 2)
    A, B = keyframe_matrix(KF[seg[0]]), keyframe_matrix(KF[seg[1]])
    want = np_interpolate(KF[seg[0]]["time"], A, KF[seg[1]]["time"], B, t)
    np.testing.assert_allclose(object_to_world(KF, t), want, rtol=2e-5, atol=2e-5)


def test_compose_quirk_on_a_keyframe(lib):
    """t exactly on a keyframe that is not the last goes through Interpolate with ratio 0, and Compose's
    crossed arguments make the result differ from the keyframe's own matrix: the upper 3x3 is
    R * diag(translation) and the last row is scale * R * diag(translation)."""
    m = object_to_world(KF, 0.0)
    A = keyframe_matrix(KF[0])
    assert not np.allclose(m, A, atol=1e-3)
    s, q, tr = np_decompose(A)
    R = np_from_quat(q)[:3, :3]
    np.testing.assert_allclose(m[:3, :3], R * tr[None, :], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(m[3, :3], (s @ R) * tr, rtol=2e-5, atol=2e-5)


def test_decompose_recovers_srt(lib):
    """Sanity of the decomposition itself: for a pure S*R*T keyframe it returns S, T and a unit
    quaternion whose matrix is R (checked through the numpy restatement and, via the quirk test's
    algebra, through the C++ one)."""
    A = keyframe_matrix(KF[1])
    s, q, tr = np_decompose(A)
    np.testing.assert_allclose(s, KF[1]["scale"], rtol=1e-5)
    np.testing.assert_allclose(tr, KF[1]["translation"], rtol=1e-6)
    np.testing.assert_allclose(np_from_quat(q)[:3, :3], rot_xyz(KF[1]["rotation"])[:3, :3], atol=1e-5)
    assert abs(np.dot(q, q) - 1) < 1e-5


def test_animated_scene_traces_where_the_matrix_says(lib):
    """End to end on the CPU side: the oracle sees the sphere where Transform.Eval(t) put it."""
    from barnacle_b200.scene import RAY_DTYPE
    from oracle.oracle_ffi import OracleScene
    kf = [{"time": 0.0, "translation": [0, 0, 0]}, {"time": 1.0, "translation": [4, 0, 0]}]
    sc = Scene.LoadString(scene_json(kf), time_=1.0)   # last keyframe: plain translation by (4, 0, 0)
    o = OracleScene(sc.desc)
    rays = np.zeros(2, RAY_DTYPE)
    rays["origin"] = [[4, 0, 5], [0, 0, 5]]
    rays["direction"] = [[0, 0, -1], [0, 0, -1]]
    rays["tmax"] = np.inf
    hits = o.trace(rays)
    assert hits["instance"][0] == 0 and abs(hits["t"][0] - 4.0) < 1e-5 and hits["instance"][1] == -1
    sc.close()
