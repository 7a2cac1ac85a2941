"""GPU parity of the step after the hot path: nearest-neighbour search, scale / translation ICP
and Chamfer distance against the oracle and the fixtures captured from the reference's own ICP_T_S."""
import glob
import os

import numpy as np
import pytest
import torch

from alignsdf_b200 import mesh as amesh, synthetic, trimesh_lite as tl
from alignsdf_b200.deep_sdf.metrics import chamfer as gchamfer
from alignsdf_b200.deep_sdf.metrics.icp_trans_scale import ICP_T_S, nn_search
from oracle import icp_oracle

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "icp_*.npz")))
DEV = torch.device("cuda")


@pytest.mark.parametrize("nq,nr", [(1, 1), (37, 5), (1000, 3001), (4096, 30000)])
def test_nearest_neighbour_is_exact(nq, nr):
    from sklearn.neighbors import KDTree
    rng = np.random.default_rng(nq * 7 + nr)
    q, r = rng.normal(size=(nq, 3)), rng.normal(size=(nr, 3))
    r[nr // 2] = r[0]                                                  # a duplicate: ties go to the smallest index
    idx, d2 = nn_search(torch.from_numpy(q).to(DEV), torch.from_numpy(r).to(DEV), want_dist=True)
    dist, ref = KDTree(r).query(q)
    got = idx.cpu().numpy()
    assert np.array_equal(np.where(got == nr // 2, 0, got), np.where(ref[:, 0] == nr // 2, 0, ref[:, 0]))
    assert not (got == nr // 2).any() or nr < 2
    assert np.allclose(d2.cpu().numpy(), dist[:, 0] ** 2, rtol=1e-12, atol=1e-300)
    assert nn_search(torch.zeros((0, 3), dtype=torch.float64, device=DEV), torch.from_numpy(r).to(DEV)).numel() == 0


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_icp_matches_the_reference_fixture(path):
    g = np.load(path)
    icp = ICP_T_S(tl.Mesh(g["source"], np.zeros((0, 3), np.int64)), tl.Mesh(g["target"], np.zeros((0, 3), np.int64)))
    icp.normalize_points()
    icp.run_icp_f(max_iter=100)
    assert len(icp.errors) == int(g["n_iter"])
    assert abs(icp.errors[-1] - float(g["final_error"])) <= 1e-12
    assert np.allclose(icp.scale, g["scale"], rtol=1e-9, atol=0)
    assert np.allclose(icp.trans, g["trans"], rtol=1e-8, atol=1e-12)
    all_trans, all_scale = icp.get_trans_scale()
    assert np.allclose(all_trans, g["all_trans"], rtol=1e-8, atol=1e-12) and np.allclose(all_scale, g["all_scale"], rtol=1e-9)
    cd = gchamfer.chamfer_points(icp.points_source * icp.scale + icp.trans, icp.points_target)
    assert abs(cd - float(g["chamfer_after"])) <= 1e-9 * max(1.0, float(g["chamfer_after"]))
    # the tree-rebuilding variant (icp_trans_scale.py:115-186) converges to the same alignment here
    icp2 = ICP_T_S(tl.Mesh(g["source"], np.zeros((0, 3), np.int64)), tl.Mesh(g["target"], np.zeros((0, 3), np.int64)))
    icp2.normalize_points()
    icp2.run_icp(max_iter=10)
    assert abs(float(np.asarray(icp2.scale).reshape(-1)[0]) - float(g["scale"][0])) <= 5e-3


ALIGN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "align_*.npz")))


@pytest.mark.parametrize("path", ALIGN, ids=[os.path.basename(p)[:-4] for p in ALIGN])
def test_alignment_helpers_match_the_reference_fixture(path):
    """chamfer.py:61-180 through the GPU neighbour search and device moments, against the outputs captured from the
    reference's own functions (oracle/make_golden_icp.py)."""
    g = np.load(path)
    for tag, kw in (("refl", dict()), ("rigid", dict(reflection=False)), ("noscale", dict(scale=False, reflection=False)),
                    ("notrans", dict(translation=False))):
        m, moved, cost = gchamfer.procrustes(g["source"], g["paired"], **kw)
        assert np.allclose(m, g["p_%s_matrix" % tag], rtol=1e-9, atol=1e-12), tag
        assert abs(cost - float(g["p_%s_cost" % tag])) <= 1e-12
        assert np.allclose(moved, icp_oracle.apply_matrix(g["source"], m), rtol=0, atol=1e-13)
    m, _, cost = gchamfer.procrustes_without_rot(g["source"], g["paired"])
    assert np.allclose(m, g["s_matrix"], rtol=1e-9, atol=1e-12) and abs(cost - float(g["s_cost"])) <= 1e-12
    thr, cap = float(g["thr"]), int(g["cap"])
    for tag, rot in (("ts", False), ("tr", True)):
        a, b, cost = gchamfer.icp(g["source"], g["target"], threshold=thr, max_iterations=cap, rot=rot)
        want = float(g["icp_%s_cost" % tag])
        assert np.allclose(a, g["icp_%s_a" % tag], rtol=1e-7, atol=1e-10) and np.allclose(b, g["icp_%s_b" % tag], rtol=1e-7, atol=1e-10)
        assert abs(cost - want) <= 1e-9 * want
    total, moved, cost = gchamfer.registration_icp(g["source"], g["target"], threshold=thr, max_iterations=cap)
    assert np.allclose(total, g["reg_matrix"], rtol=1e-7, atol=1e-10) and abs(cost - float(g["reg_cost"])) <= 1e-9 * float(g["reg_cost"])
    assert np.allclose(moved, icp_oracle.apply_matrix(g["source"], total), rtol=0, atol=1e-12)


def test_chamfer_with_rotation_fit(tmp_path):
    """chamfer.py:199-203 (optim + rot) through compute_trimesh_chamfer, against the oracle on the same samples."""
    rng = np.random.default_rng(2)
    d = rng.normal(size=(400, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    import scipy.spatial
    hull = scipy.spatial.ConvexHull(d)                                  # a closed triangle mesh on the unit sphere
    gt = tl.Mesh(d * [0.08, 0.06, 0.05], hull.simplices.astype(np.int64))
    ang = 0.12
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    pred = tl.Mesh(gt.vertices @ R.T * 1.05 + [0.004, -0.003, 0.002], gt.faces)
    gt.export(str(tmp_path / "gt.ply"))
    pred.export(str(tmp_path / "pred.ply"))
    plain = gchamfer.compute_trimesh_chamfer(str(tmp_path / "gt.ply"), str(tmp_path / "pred.ply"), rng=np.random.default_rng(3))
    fitted = gchamfer.compute_trimesh_chamfer(str(tmp_path / "gt.ply"), str(tmp_path / "pred.ply"), optim=True, rot=True,
                                              rng=np.random.default_rng(3))
    rng = np.random.default_rng(3)
    src, _ = tl.sample_surface(tl.load(str(tmp_path / "pred.ply")), 30000, rng)
    tgt, _ = tl.sample_surface(tl.load(str(tmp_path / "gt.ply")), 30000, rng)
    _, aligned, _, _ = icp_oracle.registration_icp(src, tgt)
    want = icp_oracle.chamfer(aligned, tgt)
    assert abs(fitted - want) <= 1e-7 * want and fitted < plain


def test_eval_mode_aligns_the_mesh_to_the_ground_truth_on_disk(tmp_path, monkeypatch):
    """utils/mesh.py:385-395 through the drop-in call: the written hand mesh is the ICP-aligned one, (trans, scale)
    equal the oracle's on the same surface samples, and the object mesh is moved by the hand's transform (:186-194)."""
    N = 48
    dec = synthetic.make_decoder(0)
    s = synthetic.make_sample(0).to(DEV)
    plain = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results, None,
                                               s.specs, str(tmp_path / "7_plain"), N=N)
    # ground truth = the predicted hand, scaled and shifted (what the ICP must find), as <root>/obman/test/mesh_hand/7.obj
    gt_dir = tmp_path / "data" / "obman" / "test" / "mesh_hand"
    gt_dir.mkdir(parents=True)
    true_scale, true_shift = 1.1, np.array([0.02, -0.03, 0.01])
    hv = np.asarray(plain["hand"].vertices, np.float64)
    with open(gt_dir / "7.obj", "w") as fh:
        for p in hv * true_scale + true_shift:
            fh.write("v %.17g %.17g %.17g\n" % tuple(p))
        for f in np.asarray(plain["hand"].faces) + 1:
            fh.write("f %d %d %d\n" % tuple(f))
    monkeypatch.setattr(amesh, "DATA_ROOT", str(tmp_path / "data"))
    monkeypatch.setattr(amesh, "ICP_RNG", np.random.default_rng(11))
    prefix = str(tmp_path / "7")
    res = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results, None,
                                             s.specs, prefix, N=N, eval_mode=True)
    # oracle on the same samples
    rng = np.random.default_rng(11)
    src, _ = tl.sample_surface(plain["hand"], 30000, rng)
    tgt, _ = tl.sample_surface(tl.load(str(gt_dir / "7.obj")), 30000, rng)
    moved, norm = icp_oracle.normalize(src, tgt)
    scale, trans, errors = icp_oracle.run_icp_f(moved, tgt, max_iter=100)
    o_trans, o_scale = icp_oracle.get_trans_scale(scale, trans, norm)
    assert abs(float(o_scale[0]) - true_scale) <= 2e-3 and np.abs(o_trans - true_shift).max() <= 2e-3
    aligned = tl.load(prefix + "_hand.ply")
    want = hv * o_scale + o_trans
    assert np.abs(aligned.vertices - want).max() <= 1e-6            # f32 PLY of the aligned vertices
    assert np.abs(np.asarray(res["hand"].vertices) - want).max() <= 1e-9
    ov = np.asarray(plain["obj"].vertices, np.float64)
    assert np.abs(tl.load(prefix + "_obj.ply").vertices - (ov * o_scale + o_trans)).max() <= 1e-6
