"""Phase timing of the tcgen05 kernel (CTA pair 0): where do the cycles of one 256^3 pass go?

    python tools/tc_phase_timing.py [N]
Prints the UMMA issuer's wait cycles (A operand vs weight tiles) and the epilogue warp's time per
phase, from clock64() stamps recorded by the kernel when asdf_tc_desc.debug_dev is set."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from alignsdf_b200 import engine, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
dec = synthetic.make_decoder(0)
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
for _ in range(2):
    bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc")
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
dbg[15] = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # experiment flags (bit0: skip weight copies)
bound.tc.desc.debug_dev = dbg.data_ptr()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc")
e1.record()
torch.cuda.synchronize()
bound.tc.desc.debug_dev = None
d = dbg.cpu().tolist()
items = ((N ** 3 + 127) // 128 + 73) // 74 * 2
print(f"N={N}  kernel {e0.elapsed_time(e1):.2f} ms  items/cluster ~{items}")
print(f"UMMA issuer: total {d[0]} cyc ({d[0] / items:.0f}/item), wait A {d[1]} ({d[1] / items:.0f}/item), "
      f"wait weights {d[2]} ({d[2] / items:.0f}/item), floor 24576/item")
names = ["param load M0", "gen x1", "param load rest", "wait L1", "epi1", "wait L2 (both)", "epi2 (both)",
         "wait L3 (both)", "epi3 (both)", "final+store"]
for n, v in zip(names, d[4:14]):
    print(f"  epilogue {n:18s} {v / items:9.0f} cyc/item")
