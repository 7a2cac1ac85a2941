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
# marching cubes on a volume wide enough for the interior-tile path of mc_classify (128-point rows, 17 planes)
ax = [torch.linspace(-1, 1, n, device=dev) for n in (40, 24, 264)]
g = torch.meshgrid(*ax, indexing="ij")
vol = (g[0] ** 2 + g[1] ** 2 + 0.3 * g[2] ** 2).sqrt() - 0.7 + 0.05 * torch.sin(9 * g[2])
mc = engine.marching_cubes(vol.contiguous(), 0.0, [0.1] * 3, want_keys=True)
print("mc", tuple(mc["verts"].shape), tuple(mc["faces"].shape))
# nearest neighbour + PixelAlign
from alignsdf_b200.deep_sdf.metrics.icp_trans_scale import nn_search  # noqa: E402
a, b = torch.rand(777, 3, device=dev, dtype=torch.float64), torch.rand(2050, 3, device=dev, dtype=torch.float64)
print("nn", int(nn_search(a, b).sum()))
sp = synthetic.make_sample(3, pixel_align=(10, 12)).to(dev)
bp = engine.get_engine(dec, dev).bind(sp.latent, sp.specs, sp.mano_results, sp.obj_results, cam_intr=sp.cam_intr)
hp, op, _ = bp.eval_points(torch.rand(100, 3, device=dev) * 3 - 1.5)
print("pixel align", float(hp.min()), float(op.max()))
