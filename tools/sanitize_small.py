"""Small invocations of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, mesh as amesh, synthetic  # noqa: E402

dev = torch.device("cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dec = synthetic.make_decoder(0)
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
for path in ("f8", "f16", "auto", "simt"):
    h, o, _, box = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path=path)
    torch.cuda.synchronize()
    print(path, float(h.min()), float(o.min()), box.tolist())
with tempfile.TemporaryDirectory() as td:
    res = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results, None,
                                             s.specs, os.path.join(td, "x"), N=N)
    print({k: (None if m is None else m.faces.shape) for k, m in res.items()})
# second sample of the decoder: the bounding-box pass runs on the single-product kind + exact shell
s2 = synthetic.make_batch(2)[1].to(dev)
vols = amesh.sdf_volumes(dec, s2.latent, s2.mano_results, s2.obj_results, s2.specs, N, keep_pass1=False)
print("fast bbox pass:", vols["bound"].kinds_used)
print("stats", engine.STATS)
