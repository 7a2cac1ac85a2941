"""Phase timing of the z-slab path (run under torchrun):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/slab_profile.py [N]"""
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
dec = synthetic.make_decoder(0)
s = synthetic.make_sample(0).to(dev)
bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
be = slab.gpu_backend(bound, N)
T = {}


def tick(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()


orig = dict(reduce_bbox=slab.reduce_bbox, exchange_halo=slab.exchange_halo, gather_pieces=slab.gather_pieces,
            stitch=slab.stitch)


def wrap(name):
    f = orig[name]

    def g(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f(*a, **k)
        tick(name, t0)
        return r
    setattr(slab, name, g)


for n in orig:
    wrap(n)
ev, mc = be.eval, be.mc


def ev2(*a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = ev(*a); tick("eval", t0); return r


def mc2(*a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = mc(*a); tick("mc", t0); return r


be = slab.Backend(ev2, mc2, dev)
REPS = 6
for it in range(3 + REPS):
    if it == 3:
        T.clear()
        dist.barrier(); torch.cuda.synchronize()
        t_all = time.perf_counter()
    fields = slab.two_pass_slab(be, N, rank, world)
    slab.mesh_slab(be, fields, N, rank, world)
dist.barrier(); torch.cuda.synchronize()
total = (time.perf_counter() - t_all) / REPS
print(f"rank {rank}/{world}: {1e3 * total:.2f} ms per step | " + " ".join(f"{k} {1e3 * v / REPS:.2f}" for k, v in T.items()), flush=True)
dist.destroy_process_group()
