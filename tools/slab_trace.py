"""GPU busy / idle timeline of the z-slab end-to-end call on every rank (torch.profiler; run under torchrun):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/slab_trace.py [N] [samples]"""
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, slab, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "e2e"        # e2e: the public call (files written); dev: reconstruct_slab only
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dec = synthetic.make_decoder(0, init="default")
hs = [s.to(dev) for s in synthetic.make_batch(S)]
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


bounds = [engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results) for s in hs]


def run(n):
    for i in range(n):
        s = hs[i % S]
        if mode == "dev":
            slab.reconstruct_slab(slab.gpu_backend(bounds[i % S], N, spread=True), N, rank, world, spread=True)
            continue
        slab.create_mesh_combined_decoder_slab(True, True, False, dec, s.latent, s.mano_results, s.obj_results, None,
                                               s.specs, os.path.join(tmp, f"r{rank}_{i % 2}"), N=N, spread=True)
    dist.barrier(); torch.cuda.synchronize()


run(S)
run(3)
t0 = time.perf_counter()
run(S)
wall = (time.perf_counter() - t0) / S
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(S)
out = os.path.join(tmp, f"trace{rank}.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
k = sorted([(e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev
            if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e])
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
tot = {}
for a, b, name in k:
    tot[name[:56]] = tot.get(name[:56], 0.0) + (b - a)
lines = [f"rank {rank}: {wall * 1e3:.2f} ms per sample un-profiled; GPU span {(t_end - t_begin) / 1e3 / S:.2f} ms / sample, "
         f"busy {busy / 1e3 / S:.2f}, idle {(t_end - t_begin - busy) / 1e3 / S:.2f} ({len(gaps)} gaps)"]
for g, at, name in sorted(gaps, reverse=True)[:int(os.environ.get("TRACE_GAPS", "12"))]:
    lines.append(f"   idle {g / 1e3:7.3f} ms at +{at / 1e3:8.2f} ms, ended by {name[:60]}")
for name, t in sorted(tot.items(), key=lambda x: -x[1])[:8]:
    lines.append(f"   {t / 1e3 / S:8.3f} ms/sample  {name}")
for r in range(world):
    if r == rank:
        print("\n".join(lines), flush=True)
    dist.barrier()
dist.destroy_process_group()
