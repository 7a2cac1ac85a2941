"""Where does the end-to-end time go?  GPU busy/idle timeline of the drop-in loop and of the pipelined batch API
(torch.profiler / CUPTI: no nsys in the image).  python tools/e2e_trace.py [N] [samples] [mode: loop|pipe]"""
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, mesh as amesh, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "pipe"
dev = torch.device("cuda")
dec = synthetic.make_decoder(0, init="default")
hs = synthetic.make_batch(S)
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def run(n):
    if mode == "pipe":
        amesh.create_meshes_pipelined(dec, hs[:n], [os.path.join(tmp, f"p{i % 2}") for i in range(n)], N=N, device=dev)
    else:
        for i in range(n):
            s = hs[i].to(dev)
            amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results, None,
                                               s.specs, os.path.join(tmp, f"l{i % 2}"), N=N)
    torch.cuda.synchronize()


run(2)
t0 = time.perf_counter()
run(S)
wall = time.perf_counter() - t0
print(f"{mode}: {wall / S * 1e3:.2f} ms per sample without the profiler")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(S)
out = os.path.join(tmp, "trace.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
k = sorted([(e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev
            if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e])
if not k:
    raise SystemExit("no GPU events in the trace")
t_begin, t_end = k[0][0], max(b for _, b, _ in k)
busy, cur_a, cur_b, gaps = 0.0, k[0][0], k[0][1], []
for a, b, name in k[1:]:
    if a > cur_b:
        busy += cur_b - cur_a
        gaps.append((a - cur_b, cur_b - t_begin, name))
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
busy += cur_b - cur_a
print(f"GPU span {(t_end - t_begin) / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms, idle {(t_end - t_begin - busy) / 1e3:.1f} ms "
      f"({len(gaps)} gaps)")
for g, at, name in sorted(gaps, reverse=True)[:25]:
    print(f"  idle {g / 1e3:7.3f} ms at +{at / 1e3:8.2f} ms, ended by {name[:70]}")
tot = {}
for a, b, name in k:
    tot[name[:60]] = tot.get(name[:60], 0.0) + (b - a)
for name, t in sorted(tot.items(), key=lambda x: -x[1])[:14]:
    print(f"  {t / 1e3 / S:8.3f} ms/sample  {name}")
