"""Phase timing of the z-slab path (run under torchrun):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/slab_profile.py [N]
Every phase is bracketed by device synchronisations, so the sum is larger than the un-instrumented step (printed
first); the split shows where a rank's time goes."""
import dataclasses
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, slab, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dec = synthetic.make_decoder(0, init="default")
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
plain = slab.gpu_backend(bound, N)
REPS = 6


def run(be, reps):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        slab.reconstruct_slab(be, N, rank, world)
    dist.barrier(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


run(plain, 3)
t_plain = run(plain, REPS)
T = {}


def timed(name, f):
    def g(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return g


be = dataclasses.replace(plain, **{n: timed(n, getattr(plain, n)) for n in ("pass1", "regrid", "pass2", "mc_count", "mc_emit")})
for n in ("reduce_bbox", "exchange_halo", "gather_pieces", "stitch"):
    setattr(slab, n, timed(n, getattr(slab, n)))
t_inst = run(be, REPS)
print(f"rank {rank}/{world}: {1e3 * t_plain:.2f} ms per step ({1e3 * t_inst:.2f} instrumented) | "
      + " ".join(f"{k} {1e3 * v / REPS:.2f}" for k, v in T.items()), flush=True)
dist.destroy_process_group()
