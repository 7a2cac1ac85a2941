"""A few launches of the product kernel over one 256^3 grid (ncu target).
    ncu --set full --clock-control none --import-source on -k regex:tc3_eval -s 1 -c 1 -o gpurun_out/prof python tools/run_one_pass.py [N] [path]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
path = sys.argv[2] if len(sys.argv) > 2 else "auto"
dev = torch.device("cuda")
dec = synthetic.make_decoder(0)
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
for _ in range(3):
    bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path=path)
torch.cuda.synchronize()
print("fallbacks", engine.FALLBACKS)
