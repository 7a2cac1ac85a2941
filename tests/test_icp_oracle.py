"""CPU checks of the step after the hot path: the ICP / Chamfer oracle against the fixtures captured from the
reference's own ICP_T_S class (oracle/make_golden_icp.py), and the host helpers of the drop-in (mesh loader, surface
sampler).  No GPU needed."""
import glob
import os

import numpy as np
import pytest

from alignsdf_b200 import trimesh_lite as tl
from oracle import icp_oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "icp_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_icp_oracle_reproduces_the_reference(path):
    g = np.load(path)
    moved, norm = icp_oracle.normalize(g["source"], g["target"])
    scale, trans, errors = icp_oracle.run_icp_f(moved, g["target"], max_iter=100)
    assert len(errors) == int(g["n_iter"])
    assert np.allclose(scale, g["scale"], rtol=1e-12, atol=0)
    assert np.allclose(trans, g["trans"], rtol=1e-10, atol=1e-15)
    all_trans, all_scale = icp_oracle.get_trans_scale(scale, trans, norm)
    assert np.allclose(all_trans, g["all_trans"], rtol=1e-10, atol=1e-15)
    assert np.allclose(all_scale, g["all_scale"], rtol=1e-12)
    assert abs(icp_oracle.chamfer(moved * scale + trans, g["target"]) - float(g["chamfer_after"])) <= 1e-12


def _tetra():
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float64)
    f = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], np.int64)
    return tl.Mesh(v, f)


def test_mesh_loader_reads_obj_and_our_ply(tmp_path):
    m = _tetra()
    ply = tmp_path / "t.ply"
    m.export(str(ply))
    back = tl.load(str(ply))
    assert np.array_equal(back.vertices, m.vertices) and np.array_equal(back.faces, m.faces)
    obj = tmp_path / "t.obj"
    with open(obj, "w") as fh:
        fh.write("# comment\n")
        for p in m.vertices:
            fh.write("v %.17g %.17g %.17g\n" % tuple(p))
        fh.write("vn 0 0 1\n")
        fh.write("f 1/1/1 3/2/1 2/3/1\nf 1 2 4\nf 1 4 3\nf -3 -2 -1\n")
    back = tl.load(str(obj))
    assert np.array_equal(back.vertices, m.vertices) and np.array_equal(back.faces, m.faces)
    with open(tmp_path / "q.obj", "w") as fh:                      # a quad is fan-triangulated
        fh.write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    assert np.array_equal(tl.load(str(tmp_path / "q.obj")).faces, [[0, 1, 2], [0, 2, 3]])


def test_surface_samples_lie_on_the_faces_and_follow_their_area():
    m = _tetra()
    rng = np.random.default_rng(5)
    pts, fi = tl.sample_surface(m, 40000, rng)
    assert pts.shape == (40000, 3) and fi.min() >= 0 and fi.max() <= 3
    a, b, c = (m.vertices[m.faces[fi, k]] for k in range(3))
    n = np.cross(b - a, c - a)
    assert np.abs(((pts - a) * n).sum(1)).max() <= 1e-12           # in the face's plane
    # barycentric coordinates inside the triangle
    T = np.stack([b - a, c - a], -1)
    uv = np.einsum("nij,nj->ni", np.linalg.pinv(T), pts - a)
    assert uv.min() >= -1e-12 and (uv.sum(1) <= 1 + 1e-12).all()
    share = np.bincount(fi, minlength=4) / len(fi)
    assert np.abs(share - m.area_faces / m.area).max() <= 0.01
