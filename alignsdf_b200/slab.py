"""z-slab sharding of one reconstruction across the GPUs of a box.

The reference only has sample-level parallelism (dist_reconstruct.py:63-84: one subprocess per
GPU, no communication).  This module adds the north star's second mode: grid axis 0 is cut into
``world`` contiguous slabs, one process per GPU (torchrun / torch.distributed):

  pass 1   each rank evaluates its slab and reduces a local bbox   -> ONE all_reduce(MIN) of 12 ints
  re-grid  identical arithmetic on every rank                       (utils/mesh.py:249-254)
  pass 2   each rank evaluates its slab of the refit grid
  halo     every rank publishes its first plane (both fields)       -> ONE all_gather of [2,N,N] f32
  MC       each rank meshes [z0, z1] (its slab + the neighbour's first plane) with global keys
  gather   vertex / face / key lists go to rank 0                   -> all_gather of counts + ONE packed gather
  stitch   rank 0 merges duplicate boundary vertices by key; the result is bit-identical to the
           single-GPU mesh (same vertex order = key order, same face order = cell order).

The numerical kernels are injected (``Backend``) so the collective / stitching logic can be tested
on CPU with gloo (tests/test_slab_gloo.py uses the oracle as backend); ``GpuBackend`` is the
product backend.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

INT_MAX = 2 ** 31 - 1


def slab_planes(N: int, rank: int, world: int, relief: int = 0):
    """Planes [z0, z1) of axis 0 owned by ``rank`` (as even as possible, contiguous).  ``relief`` planes are
    taken off rank 0 and spread over the others: rank 0 also gathers and stitches the mesh pieces, and with
    back-to-back samples that tail would otherwise make every other rank wait at the next collective."""
    if relief <= 0 or world < 2:
        base, rem = divmod(N, world)
        z0 = rank * base + min(rank, rem)
        return z0, z0 + base + (1 if rank < rem else 0)
    first = max(N // world - relief, 1)
    if rank == 0:
        return 0, first
    base, rem = divmod(N - first, world - 1)
    r = rank - 1
    z0 = first + r * base + min(r, rem)
    return z0, z0 + base + (1 if r < rem else 0)


def default_relief(N: int, world: int) -> int:
    """Planes to take off rank 0 (see slab_planes): the gather + stitch tail is worth ~3.4 planes of two
    passes at any N (both scale with N^2), shared with the other ranks -> 3.4 (world-1)/world, if slabs are thick
    enough for it not to matter otherwise."""
    if world < 2 or N // world < 16:
        return 0
    return int(round(3.4 * (world - 1) / world))


@dataclass
class Backend:
    """eval(begin, end, voxel, origin, bbox_mask) -> (hand [n], obj [n], box int32[12] | None);
    mc(vol [m,N,N], voxel, origin, index0_offset) -> (verts [V,3] f32, points [V,3] f32, faces [F,3] i32, keys [V] i64)"""
    eval: callable
    mc: callable
    device: torch.device
    relief: int = 0            # planes taken off rank 0 (slab_planes)


def reduce_bbox(box: torch.Tensor, group=None) -> torch.Tensor:
    """Global bbox from per-rank {min x3, max x3} x2 with a single MIN all-reduce (max is negated)."""
    b = box.clone().to(torch.int32)
    sign = torch.tensor([1, 1, 1, -1, -1, -1] * 2, dtype=torch.int32, device=b.device)
    b = b * sign
    dist.all_reduce(b, op=dist.ReduceOp.MIN, group=group)
    return b * sign


def exchange_halo(first_planes: torch.Tensor, rank: int, world: int, group=None):
    """all_gather of every rank's first plane; returns the next rank's plane (None on the last)."""
    buf = [torch.empty_like(first_planes) for _ in range(world)]
    dist.all_gather(buf, first_planes.contiguous(), group=group)
    return buf[rank + 1] if rank + 1 < world else None


def gather_pieces(pieces, rank: int, world: int, group=None):
    """Gather every rank's mesh pieces on rank 0 with two collectives.

    pieces: list (one per surface) of (verts [V,3] f32, faces [F,3] i32, keys [V] i64) on this rank.
    Returns, on rank 0, a list over ranks of such lists; None elsewhere.  One all_gather of the
    counts, then one gather of a single packed byte buffer (padded to the largest rank)."""
    dev = pieces[0][0].device
    counts = torch.tensor([[p[0].shape[0], p[1].shape[0]] for p in pieces], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    allc = torch.stack(allc).cpu()                              # [world, n_surfaces, 2]
    nbytes = ((allc[:, :, 0] * (12 + 8) + allc[:, :, 1] * 12 + 7) // 8 * 8).sum(1)   # 8-byte aligned blocks
    cap = max(int(nbytes.max()), 16)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    def as_bytes(t):
        flat_t = torch.empty(t.numel(), dtype=t.dtype, device=t.device)     # fresh unit-stride storage
        flat_t.copy_(t.reshape(-1))
        return flat_t.view(torch.uint8)
    off = 0
    for p in pieces:                                                         # keys first (8-byte aligned)
        for t in (p[2], p[0], p[1]):
            b = as_bytes(t)
            buf[off:off + b.numel()] = b
            off += b.numel()
        off = (off + 7) // 8 * 8
    recv = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, recv, dst=0, group=group)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        off, lst = 0, []
        for s_i in range(len(pieces)):
            nv, nf = int(allc[r, s_i, 0]), int(allc[r, s_i, 1])
            k = recv[r][off:off + nv * 8].view(torch.int64); off += nv * 8
            v = recv[r][off:off + nv * 12].view(torch.float32).view(nv, 3); off += nv * 12
            f = recv[r][off:off + nf * 12].view(torch.int32).view(nf, 3); off += nf * 12
            off = (off + 7) // 8 * 8
            lst.append((v, f, k))
        out.append(lst)
    return out


def stitch(parts):
    """parts: per rank (verts, faces, keys) tensors of ONE surface (any device).  Merge duplicate
    boundary vertices by key; vertex order = key order, face order = rank (== cell) order.
    Returns (verts [V,3] f32, faces [F,3] i32)."""
    keys = torch.cat([p[2] for p in parts])
    if keys.numel() == 0:
        dev = keys.device
        return torch.zeros((0, 3), dtype=torch.float32, device=dev), torch.zeros((0, 3), dtype=torch.int32, device=dev)
    verts = torch.cat([p[0] for p in parts])
    uk, inv = torch.unique(keys, sorted=True, return_inverse=True)
    # any representative of a duplicated key will do: duplicates are bit-identical by construction
    first = torch.empty(uk.numel(), dtype=torch.int64, device=keys.device)
    first[inv] = torch.arange(keys.numel(), device=keys.device)
    faces, off = [], 0
    for p in parts:
        faces.append(inv[p[1].long() + off])
        off += p[2].shape[0]
    return verts[first], torch.cat(faces).to(torch.int32)


def two_pass_slab(backend: Backend, N: int, rank: int, world: int, hand_branch=True, obj_branch=True,
                  group=None):
    """Both evaluation passes on this rank's slab.  Returns dict(hand, obj [nz,N,N], voxel, origin,
    z0, z1) -- ``voxel``/``origin`` identical on every rank."""
    from .mesh import _bbox_to_minmax, _regrid
    z0, z1 = slab_planes(N, rank, world, backend.relief)
    mask = (1 if hand_branch else 0) | (2 if obj_branch else 0)
    vs1 = 2.0 / (N - 1)
    _, _, box = backend.eval(z0 * N * N, z1 * N * N, vs1, [-1.0, -1.0, -1.0], mask)
    if box is None:                                   # empty slab (more ranks than planes)
        box = torch.tensor([INT_MAX] * 3 + [-1] * 3 + [INT_MAX] * 3 + [-1] * 3, dtype=torch.int32,
                           device=backend.device)
    box = reduce_bbox(box, group)
    mn, mx = _bbox_to_minmax(box, hand_branch, obj_branch)
    voxel, origin = _regrid(mn, mx, N, vs1)
    h, o, _ = backend.eval(z0 * N * N, z1 * N * N, float(voxel), origin.tolist(), 0)
    return dict(hand=h.view(z1 - z0, N, N), obj=o.view(z1 - z0, N, N), voxel=voxel, origin=origin, z0=z0, z1=z1)


def mesh_slab(backend: Backend, fields: dict, N: int, rank: int, world: int, which=("hand", "obj"), group=None):
    """Halo exchange + marching cubes + gather + stitch.  Returns {tag: (verts, points, faces)} as
    tensors on rank 0's device (None elsewhere); points = f32(origin) + verts like utils/mesh.py:360-363."""
    z0, z1 = fields["z0"], fields["z1"]
    nz = z1 - z0
    dev = backend.device
    first = torch.stack([fields["hand"][0] if nz else torch.zeros(N, N, device=dev),
                         fields["obj"][0] if nz else torch.zeros(N, N, device=dev)])
    halo = exchange_halo(first, rank, world, group)
    vs = float(fields["voxel"])
    org = fields["origin"].tolist()
    pieces, tags = [], []
    for ti, tag in enumerate(("hand", "obj")):
        if tag not in which:
            continue
        vol = fields[tag]
        if halo is not None and nz:
            vol = torch.cat([vol, halo[ti:ti + 1]], 0)
        if vol.shape[0] >= 2:
            v, _, f, k = backend.mc(vol.contiguous(), vs, org, z0)
        else:
            v = torch.zeros((0, 3), dtype=torch.float32, device=dev)
            f = torch.zeros((0, 3), dtype=torch.int32, device=dev)
            k = torch.zeros((0,), dtype=torch.int64, device=dev)
        pieces.append((v, f, k)); tags.append(tag)
    gathered = gather_pieces(pieces, rank, world, group)
    if rank != 0:
        return None
    out = {}
    org32 = torch.tensor(org, dtype=torch.float32, device=dev)
    for s_i, tag in enumerate(tags):
        verts, faces = stitch([gathered[r][s_i] for r in range(world)])
        out[tag] = (verts, org32[None] + verts, faces)
    return out


# ----------------------------------------------------------------------------
# product backend + public entry point
# ----------------------------------------------------------------------------
def gpu_backend(bound, N, grid_mode="reference", path=None) -> Backend:
    from . import engine

    def ev(begin, end, voxel, origin, bbox_mask):
        if end <= begin:
            e = torch.zeros(0, dtype=torch.float32, device=bound.device)
            return e, e.clone(), None
        h, o, _, box = bound.eval_grid(N, voxel, origin, grid_mode, begin, end, bbox_mask, path=path)
        return h, o, box

    def mc(vol, voxel, origin, index0_offset):
        r = engine.marching_cubes(vol, 0.0, [voxel] * 3, origin, index0_offset, want_keys=True,
                                  check_range=False)
        return r["verts"], r["points"], r["faces"], r["keys"]

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return Backend(ev, mc, bound.device, default_relief(N, world))


def create_mesh_combined_decoder_slab(hand_branch, obj_branch, cls_branch, decoder, latent_vec, mano_results,
                                      obj_results, cam_intr, specs, filename, N=256, group=None,
                                      grid_mode="reference", write=True):
    """z-slab sharded equivalent of ``mesh.create_mesh_combined_decoder`` (call on every rank of an
    initialised process group; rank 0 writes the files and returns the meshes)."""
    import logging
    from . import engine
    from .trimesh_lite import largest_watertight_component_mc
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = engine._device_of(latent_vec)
    bound = engine.get_engine(decoder, dev).bind(latent_vec, specs, mano_results, obj_results)
    be = gpu_backend(bound, N, grid_mode)
    fields = two_pass_slab(be, N, rank, world, hand_branch, obj_branch, group)
    which = tuple(t for t, use in (("hand", hand_branch), ("obj", obj_branch)) if use)
    meshes = mesh_slab(be, fields, N, rank, world, which, group)
    if rank != 0:
        return None
    result = {"hand": None, "obj": None}
    from .trimesh_lite import Mesh, export_ply_records
    vs = float(fields["voxel"])
    for tag in which:
        verts_d, points_d, faces_d = meshes[tag]
        if faces_d.shape[0] == 0:
            logging.warning("Cannot reconstruct mesh from '{}'".format(f"{filename}_{tag}.ply"))
            continue
        if points_d.is_cuda:
            # component filter + PLY face records on the GPU (csrc/cc.cu), like mesh.convert_sdf_samples_to_ply
            sel_p, sel_f, _ = engine.select_component(points_d, faces_d, verts_d, (N, N, N), [vs] * 3)
            rec = engine.ply_face_records(sel_f).cpu().numpy()
            points = sel_p.cpu().numpy()
            if tag == "obj" and hand_branch:
                points = points * np.array([1]) + np.array([0, 0, 0])  # utils/mesh.py:366-369, hand's values
            m = Mesh(points, sel_f.cpu().numpy())
            if write:
                export_ply_records(f"{filename}_{tag}.ply", points, rec)
        else:                                   # host tensors (gloo tests with the oracle as backend)
            verts, points, faces = (x.cpu().numpy() for x in meshes[tag])
            if tag == "obj" and hand_branch:
                points = points * np.array([1]) + np.array([0, 0, 0])
            m = largest_watertight_component_mc(points, faces, verts, (N, N, N), [vs] * 3)
            if write:
                m.export(f"{filename}_{tag}.ply")
        result[tag] = m
    return result
