"""world_size-2 gloo test (CPU) of the z-slab sharding logic: bbox all-reduce, halo all-gather,
variable-length gather and key-based stitching.  The numerical kernels are replaced by the oracle
(the product backend needs a GPU); the result must equal the single-process oracle bit for bit."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alignsdf_b200 import slab
from oracle import alignsdf_oracle as orc
from oracle import mc_oracle as mo
from tests import helpers


def _oracle_backend(dec, sample, N):
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    cfg = orc.decoder_cfg(dec)

    def ev(begin, end, voxel, origin, mask):
        if end <= begin:
            e = torch.zeros(0)
            return e, e.clone(), None
        with torch.no_grad():
            h, o, _ = orc.eval_volume(sd, cfg, sample.latent, sample.specs, sample.mano_results,
                                      sample.obj_results, N, voxel, origin, "reference", 2 ** 18, begin, end)
        box = None
        if mask:
            box = torch.tensor([slab.INT_MAX] * 3 + [-1] * 3 + [slab.INT_MAX] * 3 + [-1] * 3, dtype=torch.int32)
            for bi, vals in enumerate((h, o)):
                if not (mask >> bi) & 1:
                    continue
                idx = torch.nonzero(vals < 0)[:, 0] + begin
                if idx.numel():
                    ijk = torch.stack([idx // (N * N), (idx // N) % N, idx % N], 1)
                    box[6 * bi:6 * bi + 3] = ijk.min(0).values.int()
                    box[6 * bi + 3:6 * bi + 6] = ijk.max(0).values.int()
        return h, o, box

    from alignsdf_b200 import mesh as amesh
    vs1 = 2.0 / (N - 1)
    zeros = lambda: torch.zeros(slab.N_FLAGS, dtype=torch.int32)

    def pass1(begin, end, mask):
        _, _, box = ev(begin, end, vs1, [-1.0, -1.0, -1.0], mask)
        if box is None:                                   # empty slab (more ranks than planes)
            box = torch.tensor([slab.INT_MAX] * 3 + [-1] * 3 + [slab.INT_MAX] * 3 + [-1] * 3, dtype=torch.int32)
        return box, zeros()

    def regrid(box, mask):
        mn, mx = amesh._bbox_to_minmax(box, bool(mask & 1), bool(mask & 2))
        voxel, origin = amesh._regrid(mn, mx, N, vs1)
        return torch.cat([voxel.reshape(1), origin]).float()

    def pass2(begin, end, grid):
        h, o, _ = ev(begin, end, float(grid[0]), grid[1:4].tolist(), 0)
        return h, o, zeros()

    def mc_count(vol, grid, z0):
        try:
            v, f, k = mo.marching_cubes(vol.numpy(), 0.0, [float(grid[0])] * 3, (z0, 0, 0), (N, N, N))
        except ValueError:
            v, f, k = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), np.zeros(0, np.uint64)
        return torch.tensor([len(v), len(f), 0, 0, 1], dtype=torch.int64), (v, f, k)

    def mc_emit(handle, nv, nt, nseg):
        v, f, k = handle
        return torch.from_numpy(v), torch.from_numpy(f), torch.from_numpy(k.astype(np.int64))

    return slab.Backend(pass1, regrid, pass2, mc_count, mc_emit, torch.device("cpu"))


def _redo_worker(rank, world, port, name, out_dir):
    """A backend whose flags reject the kernel kind on ONE rank only, once: every rank must repeat the sample."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        meta, g, dec, sample = helpers.load_case(name)
        N = meta["N"]
        be = _oracle_backend(dec, sample, N)
        calls = dict(pass1=0, decide=[])
        inner1 = be.pass1

        def pass1(begin, end, mask):
            calls["pass1"] += 1
            box, flags = inner1(begin, end, mask)
            if calls["pass1"] == 1 and rank == 1:
                flags[0] = 1                           # "an activation left the e4m3 range" on rank 1's slab only
            return box, flags

        def decide(flags):
            calls["decide"].append(list(flags))
            return flags[0] != 0                       # repeat iff ANY rank raised the word

        be.pass1, be.decide = pass1, decide
        res = slab.reconstruct_slab(be, N, rank, world, keep_fields=True)
        assert calls["pass1"] == 2 and [f[0] for f in calls["decide"]] == [1, 0], calls
        if rank == 0:
            np.savez(os.path.join(out_dir, "redo.npz"), hand_f=res["meshes"]["hand"][2].numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_a_rejected_kernel_kind_on_one_rank_repeats_the_sample_on_all(tmp_path):
    name = "sep_both9_n24"
    mp.spawn(_redo_worker, args=(2, 29500 + os.getpid() % 2000 + 31, name, str(tmp_path)), nprocs=2, join=True)
    meta, g, dec, sample = helpers.load_case(name)
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    with torch.no_grad():
        res = orc.two_pass_field(sd, orc.decoder_cfg(dec), sample.latent, sample.specs, sample.mano_results,
                                 sample.obj_results, meta["N"])
    _, f, _ = mo.marching_cubes(res["hand"].numpy(), 0.0, [float(res["voxel"])] * 3)
    assert np.array_equal(np.load(tmp_path / "redo.npz")["hand_f"], f)


def _worker(rank, world, port, name, out_dir, spread):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        meta, g, dec, sample = helpers.load_case(name)
        N = meta["N"]
        be = _oracle_backend(dec, sample, N)
        be.relief = 2 if world == 3 else (1 if spread else 0)          # uneven slabs must give the identical result
        be.relief_ranks = 2 if spread else 1
        res = slab.reconstruct_slab(be, N, rank, world, keep_fields=True, spread=spread)
        grid = res["grid"].numpy()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), voxel=grid[0], origin=grid[1:4], hand=res["hand"].numpy(),
                 z0=res["z0"], z1=res["z1"])
        meshes = res["meshes"]
        assert (meshes is not None) == (rank == 0 or (spread and rank == 1))
        if meshes is not None:                      # hand on rank 0, object on rank 1 when the tail is spread
            assert sorted(meshes) == (["hand", "obj"] if not spread else [("hand", "obj")[rank]])
            np.savez(os.path.join(out_dir, f"mesh{rank}.npz"),
                     **{f"{t}_{n}": a.numpy() for t in meshes for n, a in zip(("v", "p", "f"), meshes[t])})
        dist.barrier()          # rank 0 may still be receiving the gathered pieces
    finally:
        dist.destroy_process_group()


def test_slab_planes_partition():
    for N, w in ((256, 8), (24, 2), (10, 4), (3, 8)):
        cuts = [slab.slab_planes(N, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    # relief: still a contiguous partition, the relieved ranks thinner, the others even among themselves
    for N, w, relief, rr in ((256, 8, 3, 1), (256, 2, 2, 1), (48, 4, 3, 1), (10, 4, 5, 1), (256, 8, 1, 2), (64, 2, 1, 2)):
        cuts = [slab.slab_planes(N, r, w, relief, rr) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        rr_eff = min(rr, w - 1)
        assert all(b - a == max(N // w - relief, 1) for a, b in cuts[:rr_eff])
        rest = [b - a for a, b in cuts[rr_eff:]]
        assert max(rest) - min(rest) <= 1
    assert slab.default_relief(256, 8) == 2 and slab.default_relief(256, 2) == 1 and slab.default_relief(32, 8) == 0
    assert slab.default_relief(256, 8, spread=True) == 1


@pytest.mark.parametrize("world,spread", [(2, False), (3, False), (2, True)])
def test_two_rank_slab_reconstruction_equals_single_process(tmp_path, world, spread):
    name = "sep_both9_n24"
    port = 29500 + os.getpid() % 2000 + world + 7 * spread
    mp.spawn(_worker, args=(world, port, name, str(tmp_path), spread), nprocs=world, join=True)
    meta, g, dec, sample = helpers.load_case(name)
    N = meta["N"]
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    # identical re-grid on every rank, equal to the reference's golden values
    for p in parts:
        assert np.float32(p["voxel"]) == g["new_voxel"] and np.array_equal(p["origin"], g["new_origin"])
    hand = np.concatenate([p["hand"] for p in parts], 0)
    assert np.abs(hand - g["pass2_hand"]).max() <= 1e-6
    m = dict(np.load(tmp_path / "mesh0.npz"))
    if spread:
        m.update(np.load(tmp_path / "mesh1.npz"))
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    with torch.no_grad():
        res = orc.two_pass_field(sd, orc.decoder_cfg(dec), sample.latent, sample.specs, sample.mano_results,
                                 sample.obj_results, N)
    for tag in ("hand", "obj"):
        v, f, _ = mo.marching_cubes(res[tag].numpy(), 0.0, [float(res["voxel"])] * 3)
        assert np.array_equal(m[f"{tag}_f"], f)
        assert np.array_equal(m[f"{tag}_v"], v)
        assert np.array_equal(m[f"{tag}_p"], (res["origin"].numpy()[None] + v).astype(np.float32))
