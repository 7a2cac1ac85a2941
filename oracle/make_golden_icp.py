"""Generate tests/golden/icp_*.npz by running the REAL reference's ICP_T_S class.  TEST INFRASTRUCTURE.

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


if __name__ == "__main__":
    main()
