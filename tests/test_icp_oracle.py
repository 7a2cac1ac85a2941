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


# --- chamfer.py:13-180: the alignment helpers ---------------------------------------------------------------------------
ALIGN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "align_*.npz")))
FLAVOURS = (("refl", dict()), ("rigid", dict(reflection=False)), ("noscale", dict(scale=False, reflection=False)),
            ("notrans", dict(translation=False)))


@pytest.mark.parametrize("path", ALIGN, ids=[os.path.basename(p)[:-4] for p in ALIGN])
def test_alignment_oracle_reproduces_the_reference(path):
    g = np.load(path)
    for tag, kw in FLAVOURS:
        m, _, cost = icp_oracle.procrustes(g["source"], g["paired"], **kw)
        assert np.allclose(m, g["p_%s_matrix" % tag], rtol=1e-12, atol=1e-15), tag
        assert abs(cost - float(g["p_%s_cost" % tag])) <= 1e-14
    m, _, cost = icp_oracle.procrustes_without_rot(g["source"], g["paired"])
    assert np.allclose(m, g["s_matrix"], rtol=1e-12, atol=1e-15) and abs(cost - float(g["s_cost"])) <= 1e-14
    for tag, rot in (("ts", False), ("tr", True)):
        a, b, cost, n_iter = icp_oracle.icp_two_sided(g["source"], g["target"], threshold=float(g["thr"]), max_iterations=int(g["cap"]), rot=rot)
        assert n_iter == int(g["icp_%s_iters" % tag])
        assert np.allclose(a, g["icp_%s_a" % tag], rtol=1e-10, atol=1e-13) and np.allclose(b, g["icp_%s_b" % tag], rtol=1e-10, atol=1e-13)
    total, _, cost, n_iter = icp_oracle.registration_icp(g["source"], g["target"], threshold=float(g["thr"]), max_iterations=int(g["cap"]))
    assert n_iter == int(g["reg_iters"]) and np.allclose(total, g["reg_matrix"], rtol=1e-9, atol=1e-12)


def check_two_sided(g, tag, a, b, cost):
    want = float(g["icp_%s_cost" % tag])
    assert np.allclose(a, g["icp_%s_a" % tag], rtol=1e-7, atol=1e-10) and np.allclose(b, g["icp_%s_b" % tag], rtol=1e-7, atol=1e-10)
    assert abs(cost - want) <= 1e-9 * want


def _host_only(monkeypatch):
    """The drop-in's alignment helpers with the device swapped for the CPU and the CUDA neighbour search for a KD-tree:
    checks the HOST logic (moments, matrices, loop / stopping rules) here; the real path is in test_gpu_metrics.py."""
    import torch
    from scipy.spatial import cKDTree
    from alignsdf_b200.deep_sdf.metrics import chamfer as gch

    def nn(query, ref, want_dist=False):
        d, i = cKDTree(ref.numpy()).query(query.numpy(), 1)
        return (torch.from_numpy(i), torch.from_numpy(d * d)) if want_dist else torch.from_numpy(i)
    monkeypatch.setattr(gch, "_device", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(gch, "nn_search", nn)
    return gch


@pytest.mark.parametrize("path", ALIGN, ids=[os.path.basename(p)[:-4] for p in ALIGN])
def test_alignment_host_logic_matches_the_reference(path, monkeypatch):
    gch = _host_only(monkeypatch)
    g = np.load(path)
    for tag, kw in FLAVOURS:
        m, moved, cost = gch.procrustes(g["source"], g["paired"], **kw)
        assert np.allclose(m, g["p_%s_matrix" % tag], rtol=1e-9, atol=1e-12), tag
        assert abs(cost - float(g["p_%s_cost" % tag])) <= 1e-12
        assert np.allclose(moved, gch.transform_points(g["source"], m), rtol=0, atol=1e-14)
        assert np.array_equal(gch.procrustes(g["source"], g["paired"], return_cost=False, **kw), m)
    m, _, cost = gch.procrustes_without_rot(g["source"], g["paired"])
    assert np.allclose(m, g["s_matrix"], rtol=1e-9, atol=1e-12) and abs(cost - float(g["s_cost"])) <= 1e-12
    for tag, rot in (("ts", False), ("tr", True)):
        a, b, cost = gch.icp(g["source"], g["target"], threshold=float(g["thr"]), max_iterations=int(g["cap"]), rot=rot)
        check_two_sided(g, tag, a, b, cost)
    total, moved, cost = gch.registration_icp(g["source"], g["target"], threshold=float(g["thr"]), max_iterations=int(g["cap"]))
    assert np.allclose(total, g["reg_matrix"], rtol=1e-7, atol=1e-10) and abs(cost - float(g["reg_cost"])) <= 1e-9 * float(g["reg_cost"])
    assert np.allclose(moved, icp_oracle.apply_matrix(g["source"], total), rtol=0, atol=1e-12)


def test_transform_points_conventions(monkeypatch):
    gch = _host_only(monkeypatch)
    p = np.random.default_rng(0).normal(size=(7, 3))
    near = np.eye(4) + 5e-9
    assert np.array_equal(gch.transform_points(p, near), p)                       # chamfer.py:48-50
    m = np.eye(4)
    m[:3, :3] *= 2.0
    m[:3, 3] = [1.0, 2.0, 3.0]
    assert np.allclose(gch.transform_points(p, m), 2 * p + [1.0, 2.0, 3.0], rtol=0, atol=1e-15)
    assert np.allclose(gch.transform_points(p, m, translate=False), 2 * p, rtol=0, atol=1e-15)
    assert gch.transform_points(np.zeros((0, 3)), m).shape == (0, 3)
    with pytest.raises(ValueError):
        gch.transform_points(p, np.eye(3))
    with pytest.raises(ValueError):
        gch.procrustes(p, p[:5])


def test_chamfer_with_rotation_fit_runs_the_one_sided_loop(tmp_path, monkeypatch):
    """chamfer.py:199-203 (optim + rot): a rotated, scaled, shifted copy of a mesh is pulled back onto it."""
    gch = _host_only(monkeypatch)
    m = _tetra()
    sub = tl.Mesh(*_subdivide(m.vertices, m.faces, 3))
    ang = 0.15
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    moved = tl.Mesh(sub.vertices @ R.T * 1.07 + [0.02, -0.01, 0.03], sub.faces)
    sub.export(str(tmp_path / "gt.ply"))
    moved.export(str(tmp_path / "pred.ply"))
    plain = gch.compute_trimesh_chamfer(str(tmp_path / "gt.ply"), str(tmp_path / "pred.ply"), rng=np.random.default_rng(3))
    fitted = gch.compute_trimesh_chamfer(str(tmp_path / "gt.ply"), str(tmp_path / "pred.ply"), optim=True, rot=True,
                                         rng=np.random.default_rng(3))
    rng = np.random.default_rng(3)
    src, _ = tl.sample_surface(tl.load(str(tmp_path / "pred.ply")), 30000, rng)
    tgt, _ = tl.sample_surface(tl.load(str(tmp_path / "gt.ply")), 30000, rng)
    _, aligned, _, _ = icp_oracle.registration_icp(src, tgt)
    want = icp_oracle.chamfer(aligned, tgt)
    assert abs(fitted - want) <= 1e-8 * want and fitted < 0.2 * plain


def _subdivide(v, f, rounds):
    for _ in range(rounds):
        mid = {}
        v = [tuple(p) for p in v]
        out = []

        def m(i, j):
            k = (min(i, j), max(i, j))
            if k not in mid:
                v.append(tuple((np.asarray(v[i]) + np.asarray(v[j])) / 2))
                mid[k] = len(v) - 1
            return mid[k]
        for a, b, c in f:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            out += [[a, ab, ca], [ab, b, bc], [ca, bc, c], [ab, bc, ca]]
        v, f = np.asarray(v, np.float64), np.asarray(out, np.int64)
    return v, f


def test_ply_loader_reads_the_layouts_other_writers_produce(tmp_path):
    """``trimesh.load`` stand-in: our own files (also empty ones), trimesh-style headers with a comment line, vertex
    normals / colours between or after x y z, double coordinates, uint index lists, big-endian and ascii files,
    quads (fan-triangulated), trailing per-face properties and elements after ``face``."""
    m = _tetra()
    own = tmp_path / "own.ply"
    m.export(str(own))
    back = tl.load(str(own))
    assert back.vertices.dtype == np.float64 and back.faces.dtype == np.int64
    assert np.array_equal(back.vertices, m.vertices) and np.array_equal(back.faces, m.faces)
    tl.export_ply(str(tmp_path / "empty.ply"), np.zeros((0, 3)), np.zeros((0, 3), np.int64))
    e = tl.load(str(tmp_path / "empty.ply"))
    assert e.vertices.shape == (0, 3) and e.faces.shape == (0, 3)
    tl.export_ply(str(tmp_path / "points.ply"), m.vertices, np.zeros((0, 3), np.int64))
    assert np.array_equal(tl.load(str(tmp_path / "points.ply")).vertices, m.vertices)

    v32 = m.vertices.astype("<f4")
    nrm = np.arange(12, dtype="<f4").reshape(4, 3)
    rgba = np.arange(16, dtype=np.uint8).reshape(4, 4)
    # trimesh-style: comment, normals and colours after the coordinates, uint indices, a per-face colour after the list
    vrec = np.zeros(4, dtype=[("p", "<f4", (3,)), ("n", "<f4", (3,)), ("c", "u1", (4,))])
    vrec["p"], vrec["n"], vrec["c"] = v32, nrm, rgba
    frec = np.zeros(4, dtype=[("k", "u1"), ("i", "<u4", (3,)), ("c", "u1", (3,))])
    frec["k"], frec["i"] = 3, m.faces
    head = ("ply\nformat binary_little_endian 1.0\ncomment https://github.com/mikedh/trimesh\nelement vertex 4\n"
            "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
            "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
            "element face 4\nproperty list uchar uint vertex_indices\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
            "element edge 1\nproperty int vertex1\nproperty int vertex2\nend_header\n")
    with open(tmp_path / "rich.ply", "wb") as fh:
        fh.write(head.encode() + vrec.tobytes() + frec.tobytes() + np.array([0, 1], "<i4").tobytes())
    r = tl.load(str(tmp_path / "rich.ply"))
    assert np.array_equal(r.vertices, v32.astype(np.float64)) and np.array_equal(r.faces, m.faces)

    # big-endian doubles, a quad and a triangle (variable-length records)
    head = ("ply\nformat binary_big_endian 1.0\nelement vertex 4\nproperty double x\nproperty double y\nproperty double z\n"
            "element face 2\nproperty list uchar int vertex_index\nend_header\n")
    body = m.vertices.astype(">f8").tobytes() + b"\x04" + np.array([0, 1, 2, 3], ">i4").tobytes() \
        + b"\x03" + np.array([0, 2, 1], ">i4").tobytes()
    with open(tmp_path / "big.ply", "wb") as fh:
        fh.write(head.encode() + body)
    b = tl.load(str(tmp_path / "big.ply"))
    assert np.array_equal(b.vertices, m.vertices) and np.array_equal(b.faces, [[0, 1, 2], [0, 2, 3], [0, 2, 1]])

    # ascii, colours BEFORE the coordinates
    with open(tmp_path / "ascii.ply", "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 4\nproperty uchar red\nproperty float x\n"
                 "property float y\nproperty float z\nelement face 3\nproperty list uchar int vertex_indices\nend_header\n")
        for k, p in enumerate(m.vertices):
            fh.write("%d %r %r %r\n" % (k, float(p[0]), float(p[1]), float(p[2])))
        fh.write("3 0 2 1\n4 0 1 3 2\n3 1 2 3\n")
    a = tl.load(str(tmp_path / "ascii.ply"))
    assert np.array_equal(a.vertices, m.vertices)
    assert np.array_equal(a.faces, [[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]])

    for bad, text in (("nohdr.ply", b"ply\nformat ascii 1.0\n"), ("fmt.ply", b"ply\nformat binary_middle_endian 1.0\nend_header\n"),
                      ("noxyz.ply", b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float a\nend_header\n1\n")):
        with open(tmp_path / bad, "wb") as fh:
            fh.write(text)
        with pytest.raises(ValueError):
            tl.load(str(tmp_path / bad))
