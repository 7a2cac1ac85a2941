"""z-slab sharding on real GPUs (needs >= 2 devices; skipped otherwise): the slab-parallel
reconstruction must be bit-identical to the single-GPU one."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from alignsdf_b200 import slab, synthetic
        dec = synthetic.make_decoder(0)
        s = synthetic.make_sample(0).to(torch.device("cuda", rank))
        res = slab.create_mesh_combined_decoder_slab(True, True, False, dec, s.latent, s.mano_results,
                                                     s.obj_results, None, s.specs, os.path.join(out_dir, "slab"), N=48)
        if rank == 0:
            np.savez(os.path.join(out_dir, "slab.npz"), hv=res["hand"].vertices, hf=res["hand"].faces,
                     ov=res["obj"].vertices, of=res["obj"].faces)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slab_reconstruction_equals_single_gpu(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, 29700 + os.getpid() % 1000, str(tmp_path)), nprocs=world, join=True)
    from alignsdf_b200 import mesh as amesh, synthetic
    dec = synthetic.make_decoder(0)
    s = synthetic.make_sample(0).to(torch.device("cuda", 0))
    res = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                             None, s.specs, str(tmp_path / "single"), N=48)
    m = np.load(tmp_path / "slab.npz")
    assert np.array_equal(m["hf"], res["hand"].faces) and np.array_equal(m["hv"], res["hand"].vertices)
    assert np.array_equal(m["of"], res["obj"].faces) and np.array_equal(m["ov"], res["obj"].vertices)
    for tag in ("hand", "obj"):
        assert open(tmp_path / f"slab_{tag}.ply", "rb").read() == open(tmp_path / f"single_{tag}.ply", "rb").read()
