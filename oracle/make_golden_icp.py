"""Generate tests/golden/icp_*.npz by running the REAL reference's ICP_T_S class, and tests/golden/align_*.npz by
running the REAL reference's chamfer.py alignment helpers (procrustes, procrustes_without_rot, icp).  TEST INFRASTRUCTURE.

Runs only in the authoring container (needs /root/reference).  Shims: empty modules for the packages that are not installed
(trimesh, plyfile, skimage, ...: the class only touches trimesh in sample_mesh / export, which are not called -- the
seeded point clouds are injected), and ``np.float``
(removed from NumPy >= 1.24; icp_trans_scale.py:40,88 still use it).  Asserts oracle/icp_oracle.py reproduces the
reference's result -- this PINS the oracle -- and stores inputs + the reference's outputs.

    python oracle/make_golden_icp.py
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import icp_oracle  # noqa: E402


def clouds(seed, n_s, n_t, scale, shift, noise):
    """A bumpy closed surface sampled twice: the target is a scaled, shifted, re-sampled, noisy copy."""
    g = np.random.default_rng(seed)

    def surf(n):
        d = g.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = 0.08 * (1.0 + 0.3 * np.sin(5 * d[:, 0]) * np.cos(3 * d[:, 1]) + 0.2 * d[:, 2] ** 2)
        return d * r[:, None] * np.array([1.0, 0.7, 0.5])
    src = surf(n_s) + g.normal(scale=noise, size=(n_s, 3))
    tgt = surf(n_t) * scale + np.asarray(shift) + g.normal(scale=noise, size=(n_t, 3))
    return src, tgt


CASES = [dict(name="icp_a", seed=1, n_s=1500, n_t=1300, scale=1.25, shift=(0.02, -0.01, 0.03), noise=2e-4),
         dict(name="icp_b", seed=2, n_s=900, n_t=1100, scale=0.8, shift=(-0.05, 0.04, 0.0), noise=1e-3),
         dict(name="icp_c", seed=3, n_s=2000, n_t=2000, scale=1.0, shift=(0.0, 0.0, 0.0), noise=5e-4)]


def main():
    warnings.filterwarnings("ignore")
    for name in ("trimesh", "plyfile", "skimage", "skimage.measure", "lmdb", "chumpy"):   # not installed; never called here
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    if not hasattr(np, "float"):
        np.float = float                                        # icp_trans_scale.py:40,88
    sys.path.insert(0, REF)
    from deep_sdf.metrics.icp_trans_scale import ICP_T_S        # the reference's own class
    out_dir = os.path.join(ROOT, "tests", "golden")
    for c in CASES:
        src, tgt = clouds(c["seed"], c["n_s"], c["n_t"], c["scale"], c["shift"], c["noise"])
        mesh_s = types.SimpleNamespace(vertices=src.copy())
        mesh_t = types.SimpleNamespace(vertices=tgt.copy())
        ref = ICP_T_S(mesh_s, mesh_t)
        # sample_mesh (:19-30) minus the trimesh sampling: the normalisation on the injected clouds
        ref.offset_source = ref.points_source.mean(0)
        ref.scale_source = np.sqrt(((ref.points_source - ref.offset_source) ** 2).sum() / len(ref.points_source))
        ref.offset_target = ref.points_target.mean(0)
        ref.scale_target = np.sqrt(((ref.points_target - ref.offset_target) ** 2).sum() / len(ref.points_target))
        ref.points_source = (ref.points_source - ref.offset_source) / ref.scale_source * ref.scale_target + ref.offset_target
        ref.run_icp_f(max_iter=100)
        all_trans, all_scale = ref.get_trans_scale()
        moved, norm = icp_oracle.normalize(src, tgt)
        assert np.array_equal(moved, ref.points_source)
        scale, trans, errors = icp_oracle.run_icp_f(moved, tgt, max_iter=100)
        o_trans, o_scale = icp_oracle.get_trans_scale(scale, trans, norm)
        assert np.allclose(scale, np.asarray(ref.scale).reshape(-1), rtol=1e-12, atol=0), (scale, ref.scale)
        assert np.allclose(trans, np.asarray(ref.trans).reshape(1, 3), rtol=1e-10, atol=1e-15)
        assert np.allclose(o_trans, all_trans, rtol=1e-10, atol=1e-15) and np.allclose(o_scale, all_scale, rtol=1e-12)
        cd = icp_oracle.chamfer(moved * scale + trans, tgt)
        np.savez_compressed(os.path.join(out_dir, c["name"] + ".npz"), source=src, target=tgt,
                            scale=np.asarray(ref.scale, np.float64).reshape(-1),
                            trans=np.asarray(ref.trans, np.float64).reshape(1, 3),
                            all_scale=np.asarray(all_scale, np.float64).reshape(-1),
                            all_trans=np.asarray(all_trans, np.float64).reshape(1, 3),
                            n_iter=len(errors), final_error=errors[-1], chamfer_after=cd)
        print(c["name"], "iterations", len(errors), "scale", float(np.asarray(ref.scale).reshape(-1)[0]),
              "final error", errors[-1], "chamfer", cd)
    align_cases(out_dir)


def rotation(axis, angle):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


ALIGN = [dict(name="align_a", seed=11, n=1200, scale=1.15, shift=(0.01, -0.02, 0.015), angle=0.20, noise=2e-4, thr=1e-5, cap=20),
         dict(name="align_b", seed=12, n=800, scale=0.9, shift=(-0.03, 0.01, 0.0), angle=0.05, noise=1e-3, thr=1e-9, cap=4),
         dict(name="align_c", seed=13, n=1500, scale=1.0, shift=(0.0, 0.0, 0.0), angle=0.0, noise=5e-4, thr=1e-12, cap=20)]


def align_cases(out_dir):
    """chamfer.py:61-180 of the reference on seeded clouds; ``thr`` below the file's 1e-5 default makes the loops run
    more than two rounds (the clouds are in metres, costs ~1e-6); ``cap`` stops align_b before the two-sided loop has
    shrunk both clouds onto single points (scale is free on both sides: with this cloud it collapses after 9 rounds,
    and where it lands is rounding noise -- nothing to compare)."""
    from deep_sdf.metrics import chamfer as ref              # the reference's own module (trimesh stubbed, unused here)
    for c in ALIGN:
        src, tgt = clouds(c["seed"], c["n"], c["n"] + 100, c["scale"], c["shift"], c["noise"])
        tgt = tgt @ rotation((1.0, 2.0, -1.0), c["angle"]).T
        g = np.random.default_rng(c["seed"] + 100)
        paired = (c["scale"] * src) @ rotation((0.3, -1.0, 0.5), 2 * c["angle"] + 0.1).T + np.asarray(c["shift"]) \
            + g.normal(scale=c["noise"], size=src.shape)
        out = dict(source=src, target=tgt, paired=paired, thr=c["thr"], cap=c["cap"])
        # one matched step, all three flavours
        for tag, kw in (("refl", dict()), ("rigid", dict(reflection=False)), ("noscale", dict(scale=False, reflection=False)),
                        ("notrans", dict(translation=False))):
            m, t, cost = ref.procrustes(src, paired, **kw)
            om, ot, ocost = icp_oracle.procrustes(src, paired, **kw)
            assert np.allclose(om, m, rtol=1e-12, atol=1e-15) and np.allclose(ot, t, rtol=1e-12, atol=1e-15)
            assert abs(ocost - cost) <= 1e-14 * max(1.0, abs(cost))
            out["p_%s_matrix" % tag], out["p_%s_cost" % tag] = m, cost
        m, t, cost = ref.procrustes_without_rot(src, paired)
        om, ot, ocost = icp_oracle.procrustes_without_rot(src, paired)
        assert np.allclose(om, m, rtol=1e-12, atol=1e-15) and np.allclose(ot, t, rtol=1e-12, atol=1e-15)
        out["s_matrix"], out["s_cost"] = m, cost
        # the two-sided loop, without and with rotation
        for tag, rot in (("ts", False), ("tr", True)):
            ta, tb, cost = ref.icp(src, tgt, threshold=c["thr"], max_iterations=c["cap"], rot=rot)
            oa, ob, ocost, n_iter = icp_oracle.icp_two_sided(src, tgt, threshold=c["thr"], max_iterations=c["cap"], rot=rot)
            assert np.allclose(oa, ta, rtol=1e-10, atol=1e-13) and np.allclose(ob, tb, rtol=1e-10, atol=1e-13), tag
            assert abs(ocost - cost) <= 1e-12 * max(1.0, abs(cost))
            out["icp_%s_a" % tag], out["icp_%s_b" % tag], out["icp_%s_cost" % tag], out["icp_%s_iters" % tag] = ta, tb, cost, n_iter
        # the one-sided loop (trimesh.registration.icp restated; unpinned): stored from the ORACLE for regression only
        total, moved, cost, n_iter = icp_oracle.registration_icp(src, tgt, threshold=c["thr"], max_iterations=c["cap"])
        out["reg_matrix"], out["reg_cost"], out["reg_iters"] = total, cost, n_iter
        np.savez_compressed(os.path.join(out_dir, c["name"] + ".npz"), **out)
        print(c["name"], "two-sided iterations", int(out["icp_ts_iters"]), int(out["icp_tr_iters"]), "one-sided", n_iter,
              "costs", float(out["icp_ts_cost"]), float(out["icp_tr_cost"]), cost)


if __name__ == "__main__":
    main()
