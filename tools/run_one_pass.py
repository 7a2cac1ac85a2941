"""A few launches of the product kernel over one 256^3 grid (ncu target).
    ncu --set full --clock-control none --import-source on -k regex:tc_eval -s 1 -c 1 -o gpurun_out/prof python tools/run_one_pass.py [N] [path] [init]
path: f8 (fp16 + 2 e4m3, the kind the bench decoder runs), f16 (3 x fp16), auto (adds the calibration launches)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
path = sys.argv[2] if len(sys.argv) > 2 else "f8"
init = sys.argv[3] if len(sys.argv) > 3 else "default"
dev = torch.device("cuda")
dec = synthetic.make_decoder(0, init=init)
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
if path == "f1":          # the bounding-box pass on the single-product kind (after one calibration)
    from alignsdf_b200 import _lib
    bound._calibrate(); bound.verify()
    q = engine.make_query(_lib.QUERY_GRID_REFERENCE, N, 0, N ** 3, 2.0 / (N - 1), (-1.0, -1.0, -1.0), bbox_mask=3)
    for _ in range(3):
        bound.fast_bbox_pass(engine.LEVEL_KIND[engine.get_engine(dec, dev).level], q, N ** 3, engine.new_bbox(dev),
                             engine.get_engine(dec, dev).fast_tau())
else:
    for _ in range(3):
        bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path=path)
torch.cuda.synchronize()
print("stats", engine.STATS)
