"""Fast bounding-box pass: error bound of the single-product kind, threshold, size of the ambiguous shell, timing.
    python tools/fast_bbox_probe.py [N] [init] [gain]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import _lib, engine, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
init = sys.argv[2] if len(sys.argv) > 2 else "default"
gain = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dev = torch.device("cuda")
dec = synthetic.make_decoder(0, init=init, out_gain=gain)
s = synthetic.make_sample(0).to(dev)
eng = engine.get_engine(dec, dev)
bound = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
bound._calibrate()
lvl = bound.verify()
print(f"level {lvl} calib {eng.calib}")
kind = engine.LEVEL_KIND[min(lvl, 1)]
n = N ** 3
q = engine.make_query(_lib.QUERY_GRID_REFERENCE, N, 0, n, 2.0 / (N - 1), (-1.0, -1.0, -1.0), bbox_mask=3)


def timed(f, reps=3):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def exact():
    box = engine.new_bbox(dev)
    bound.launch_tc(kind, q, n, False, box)
    return box


want = exact()
print(f"exact pass: {timed(exact):.2f} ms")
for f in (1.0, 2.0, 4.0):
    tau = eng.fast_tau() * f
    a0 = engine.STATS["fast_bbox_ambiguous"]

    def fast():
        box = engine.new_bbox(dev)
        bound.fast_bbox_pass(kind, q, n, box, tau)
        return box
    got = fast()
    amb = engine.STATS["fast_bbox_ambiguous"] - a0
    print(f"tau {tau:.3e} ({f} x default): equal {torch.equal(got, want)}, ambiguous entries {amb} = {amb / n:.4f} of the grid, "
          f"fast pass {timed(fast):.2f} ms, redone {engine.STATS['fast_bbox_redone']}")
