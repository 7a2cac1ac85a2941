"""Generate tests/golden/viz_label.npz: the label-visualisation files of the REAL reference on a small labelled mesh.
TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference).

    python oracle/make_golden_viz.py
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import install_shims  # noqa: E402


def main():
    warnings.filterwarnings("ignore")
    install_shims()
    import utils.mesh as ref_mesh                       # the reference's module
    rng = np.random.default_rng(12)
    pts = torch.from_numpy(rng.normal(scale=0.3, size=(37, 3)).astype(np.float32))
    labels = torch.from_numpy(rng.integers(0, 6, 37).astype(np.float32))
    faces = rng.integers(0, 37, (50, 3)).astype(np.int32)
    offset, scale = np.array([[0.01, -0.02, 0.03]]), np.array([1.25])
    out = dict(points=pts.numpy(), labels=labels.numpy(), faces=faces, offset=offset, scale=scale)
    with tempfile.TemporaryDirectory() as td:
        for tag, off, sc in (("plain", None, None), ("moved", offset, scale)):
            ref_mesh.write_verts_label_to_obj(pts, labels, os.path.join(td, "a.obj"), off, sc)
            ref_mesh.write_color_labeled_ply(pts, faces, labels, os.path.join(td, "a.ply"), off, sc)
            out[f"obj_{tag}"] = np.frombuffer(open(os.path.join(td, "a.obj"), "rb").read(), np.uint8)
            out[f"ply_{tag}"] = np.frombuffer(open(os.path.join(td, "a.ply"), "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "viz_label.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
