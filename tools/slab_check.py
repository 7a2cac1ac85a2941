"""Where does a z-slab reconstruction differ from the single-GPU one?  (run under torchrun)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/slab_check.py [N] [init]
Compares, stage by stage: lattice, per-slab fields, stitched raw mesh, final mesh."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import engine, mesh as amesh, slab, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
init = sys.argv[2] if len(sys.argv) > 2 else "default"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dec = synthetic.make_decoder(0, init=init)
s = synthetic.make_sample(0).to(dev)
eng = engine.get_engine(dec, dev)
bound = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
res = slab.reconstruct_slab(slab.gpu_backend(bound, N), N, rank, world, keep_fields=True)
vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N, keep_pass1=False)
g = res["grid"].cpu()
ok_grid = float(g[0]) == float(vols["voxel"]) and torch.equal(g[1:4], vols["origin"])
z0, z1 = res["z0"], res["z1"]
dh = (res["hand"] - vols["hand"][z0:z1]).abs().max().item() if z1 > z0 else 0.0
do = (res["obj"] - vols["obj"][z0:z1]).abs().max().item() if z1 > z0 else 0.0
print(f"rank {rank}: planes [{z0},{z1}) level {eng.level} kinds {sorted(bound.kinds_used)} grid equal {ok_grid} "
      f"max|dfield| hand {dh:.3e} obj {do:.3e}", flush=True)
if rank == 0:
    vs, org = float(vols["voxel"]), vols["origin"].tolist()
    for tag in ("hand", "obj"):
        full = engine.marching_cubes(vols[tag], 0.0, [vs] * 3, org, want_keys=True)
        v, p, f = res["meshes"][tag]
        same_shape = v.shape == full["verts"].shape and f.shape == full["faces"].shape
        print(f"  {tag}: slab V {tuple(v.shape)} F {tuple(f.shape)} | single V {tuple(full['verts'].shape)} F {tuple(full['faces'].shape)}",
              flush=True)
        if same_shape:
            bad_v = (v != full["verts"]).any(1).nonzero().flatten()
            bad_p = (p != full["points"]).any(1).nonzero().flatten()
            bad_f = (f != full["faces"]).any(1).nonzero().flatten()
            print(f"  {tag}: verts differ {bad_v.numel()} (first {bad_v[:3].tolist()}), points differ {bad_p.numel()}, "
                  f"faces differ {bad_f.numel()} (first {bad_f[:3].tolist()})", flush=True)
            if bad_p.numel():
                i = int(bad_p[0])
                print("   ", p[i].tolist(), full["points"][i].tolist(), v[i].tolist(), org, flush=True)
dist.barrier()
dist.destroy_process_group()
