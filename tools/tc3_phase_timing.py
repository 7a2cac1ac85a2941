"""Phase timing of the v3 tcgen05 kernel (CTA pair 0).   python tools/tc3_phase_timing.py [N]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import _lib, engine, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
dec = synthetic.make_decoder(0)
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
for _ in range(2):
    bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc3")
tc2 = bound._tc3_for(2.0)
q = _lib.Query()
q.mode, q.N, q.begin, q.end, q.voxel = 0, N, 0, N ** 3, 2.0 / (N - 1)
q.origin[:] = [-1.0, -1.0, -1.0]
hand = torch.empty(N ** 3, device=dev); obj = torch.empty(N ** 3, device=dev)
dbg = torch.zeros(512, dtype=torch.int64, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
REPS = int(os.environ.get("ASDF_TC3_REPS", 1))
for _ in range(REPS - 1):       # long enough to reach the power-capped steady state
    _lib.check(_lib.lib().asdf_tc3_eval_debug(_lib.ptr(bound.engine.tc3_static), _lib.ptr(tc2.sample), C.byref(q),
                                              _lib.ptr(hand), _lib.ptr(obj), None, _lib.ptr(status), _lib.stream_ptr(dev), _lib.ptr(dbg)), "dbg")
dbg.zero_()
e0.record()
_lib.check(_lib.lib().asdf_tc3_eval_debug(_lib.ptr(bound.engine.tc3_static), _lib.ptr(tc2.sample), C.byref(q),
                                          _lib.ptr(hand), _lib.ptr(obj), None, _lib.ptr(status), _lib.stream_ptr(dev), _lib.ptr(dbg)), "dbg")
e1.record()
torch.cuda.synchronize()
d = dbg.cpu().tolist()
items = ((N ** 3 + 255) // 256 + 73) // 74 * 2
print(f"flags={os.environ.get('ASDF_TC3_DEBUG_FLAGS', 0)} N={N}  kernel {e0.elapsed_time(e1):.2f} ms  items/cluster ~{items}  (UMMA floor 33664 cyc/item)")
print(f"issuer: total {d[0] / items:.0f}/item | wait A {d[1] / items:.0f} | wait ring {d[2] / items:.0f} | wait acc-free {d[3] / items:.0f}")
ph = d[8:24]
for l in range(4):
    print(f"  layer {l}: epilogue waits for accumulator {ph[l] / items:8.0f}  hold-wait {ph[4 + l] / items:8.0f}  epilogue work {ph[8 + l] / items:8.0f}")
print(f"  final: {ph[12] / items:.0f}")

if not int(os.environ.get("ASDF_TC3_DEBUG_FLAGS", 0)) & 32:
    sys.exit(0)
import numpy as np
fine = np.asarray(d[32:32 + 266], dtype=np.float64) / items
ring, await_, accw = fine[:126].reshape(14, 9), fine[126:252].reshape(14, 9), fine[252:266]
np.set_printoptions(linewidth=200, suppress=True, precision=0)
print("issuer ring waits per N block (rows) x fill (P tile, chunk 0..7), cycles per item:")
print(ring)
print("issuer A waits per N block x fill:")
print(await_)
print("issuer accumulator-free waits per N block:", accw)
