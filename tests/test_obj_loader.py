"""MeshPrimitive.Load (Extensions/Primitive/Mesh.fs:246-279) in the host-side stand-in: `v` / `f` lines only, fan
triangulation of polygons, `a/b/c` tokens cut at the first slash, negative (relative) indices, everything else ignored."""
import json

import numpy as np

from barnacle_b200.scene import Scene

OBJ = """# a comment
mtllib ignored.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vn 0 0 1
vt 0.5 0.5
v 0.5 0.5 1
f 1 2 3 4
f 1/1/1 2/1/1 5/1/1
f -1 -3 -2
   f   2//1   3//1   5//1
"""


def test_obj_polygons_slashes_and_negative_indices(lib, tmp_path):
    (tmp_path / "m.obj").write_text(OBJ)
    text = json.dumps({
        "nodes": [{"children": [1, 2]}, {"instances": [0]}, {"has-camera": True}],
        "instances": [{"primitive": 0, "light": 0}], "transforms": [], "primitives": [{"type": "mesh", "uri": "m.obj"}],
        "materials": [], "lights": [{"type": "diffuse", "emission": [1, 1, 1]}],
        "integrator": {"type": "path-tracing", "spp": 1}, "camera": {"type": "pinhole"},
        "film": {"width": 8, "height": 8, "tone-mapping": "identity"}})
    scene = Scene.LoadString(text, base_dir=str(tmp_path))
    d = scene.desc.contents
    assert d.mesh_count == 1 and d.vertex_count == 5 and d.triangle_count == 5
    verts = np.ctypeslib.as_array(d.vertices, shape=(5, 3))
    np.testing.assert_array_equal(verts, [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]])
    tris = np.ctypeslib.as_array(d.triangles, shape=(5, 3))
    perm = scene.triangle_permutation(0)                      # BLAS slot -> original (file-order) triangle
    original = np.empty_like(tris)
    original[perm] = tris
    # quad 1 2 3 4 -> (0,1,2) (0,2,3); 1/.. 2/.. 5/.. -> (0,1,4); -1 -3 -2 with 5 vertices -> (4,2,3); 2// 3// 5// -> (1,2,4)
    assert original.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [4, 2, 3], [1, 2, 4]]
    # AliasTable over triangle areas (Mesh.fs:181-186): pdf_i = area_i / sum
    alias = [(d.alias[i].alias, d.alias[i].prob, d.alias[i].pdf) for i in range(5)]
    v = verts.astype(np.float64)
    areas = np.array([0.5 * np.linalg.norm(np.cross(v[b] - v[a], v[c] - v[a])) for a, b, c in tris])
    np.testing.assert_allclose([p for _, _, p in alias], areas / areas.sum(), rtol=1e-6)
    assert [a for a, _, _ in alias] == list(range(5)) and all(p == 1.0 for _, p, _ in alias)   # never aliased
