"""z-slab sharding of one reconstruction across the GPUs of a box.

The reference only has sample-level parallelism (dist_reconstruct.py:63-84: one subprocess per
GPU, no communication).  This module adds the north star's second mode: grid axis 0 is cut into
``world`` contiguous slabs, one process per GPU (torchrun / torch.distributed):

  pass 1   each rank evaluates its slab and reduces a local bbox   -> ONE all_reduce(MIN) of 17 ints
                                                                      (bbox + the kernel's range / calibration flags)
  re-grid  identical arithmetic on every rank, on the device        (utils/mesh.py:249-254, asdf_regrid)
  pass 2   each rank evaluates its slab of the refit grid
  halo     every rank sends its first plane (both fields) to the rank below it    -> ONE neighbour send / recv
  MC       each rank counts the surface of [z0, z1] (its slab + the neighbour's first plane)
  sizes    vertex / face counts and the pass-2 flags of every rank -> ONE all_gather of 11 ints, read on the host:
           the only point where a rank waits for its GPU.  If any rank's flags reject the kernel kind in use,
           EVERY rank repeats the sample with the next safer kind (the stitched field never mixes two kinds).
  gather   vertices / faces / global keys go to rank 0 un-padded   -> point-to-point sends of the exact sizes
  stitch   rank 0 drops each slab's copies of the next slab's first-plane vertices and re-indexes their faces
           by key look-up; the result is bit-identical to the single-GPU mesh (same vertex order = key order,
           same face order = cell order).

The numerical kernels are injected (``Backend``) so the collective / stitching logic can be tested
on CPU with gloo (tests/test_slab_gloo.py uses the oracle as backend); ``gpu_backend`` is the product one.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

INT_MAX = 2 ** 31 - 1
INT64_MAX = 2 ** 63 - 1
N_FLAGS = 5                    # engine.BoundSample.pending_flags()


def slab_planes(N: int, rank: int, world: int, relief: int = 0, relief_ranks: int = 1):
    """Planes [z0, z1) of axis 0 owned by ``rank`` (as even as possible, contiguous).  ``relief`` planes are
    taken off each of the first ``relief_ranks`` ranks and spread over the others: those ranks also stitch, filter
    and write a surface, and with back-to-back samples that tail would otherwise make every other rank wait at
    the next collective."""
    relief_ranks = min(relief_ranks, world - 1)
    if relief <= 0 or world < 2 or relief_ranks < 1:
        base, rem = divmod(N, world)
        z0 = rank * base + min(rank, rem)
        return z0, z0 + base + (1 if rank < rem else 0)
    first = max(N // world - relief, 1)
    if rank < relief_ranks:
        return rank * first, (rank + 1) * first
    base, rem = divmod(N - relief_ranks * first, world - relief_ranks)
    r = rank - relief_ranks
    z0 = relief_ranks * first + r * base + min(r, rem)
    return z0, z0 + base + (1 if r < rem else 0)


def default_relief(N: int, world: int, spread: bool = False) -> int:
    """Planes to take off a surface-owning rank (see slab_planes): the gather + stitch tail of BOTH surfaces is worth
    ~2 planes of two passes at any N (both scale with N^2) -- one plane per owner when the tail is spread over two
    ranks --, if slabs are thick enough for it not to matter otherwise."""
    if world < 2 or N // world < 16:
        return 0
    return int(round((RELIEF_PLANES / 2 if spread else RELIEF_PLANES) * (world - 1) / world))


RELIEF_PLANES = 2.0


@dataclass
class Backend:
    """The numerical kernels behind the slab logic.  All calls are asynchronous where the device allows it.

    pass1(begin, end, mask)          -> (box int32[12], flags int32[N_FLAGS])      bbox-only evaluation of grid indices [begin, end)
    regrid(box, mask)                -> grid f32[4] = (voxel, origin x3)           utils/mesh.py:198-256 on the reduced box
    pass2(begin, end, grid)          -> (hand [n], obj [n], flags int32[N_FLAGS])
    mc_count(vol, grid, z0)          -> (totals int64[5]: n_verts, n_tris, -, -, n_segments; handle)
    mc_emit(handle, nv, nt, nseg)    -> (verts [V,3] f32, faces [F,3] i32, keys [V] i64)
    decide(flags: list[int])         -> True when the sample must be repeated (the backend has switched to a safer
                                        kernel kind); None: never"""
    pass1: callable
    regrid: callable
    pass2: callable
    mc_count: callable
    mc_emit: callable
    device: torch.device
    relief: int = 0            # planes taken off each surface-owning rank (slab_planes)
    decide: callable = None
    relief_ranks: int = 1      # 2 when the tail is spread (hand on rank 0, object on rank 1)


def _peer(group, r):
    return r if group is None else dist.get_global_rank(group, r)


_SIGN = {}


def reduce_bbox(box: torch.Tensor, flags: torch.Tensor = None, group=None):
    """Global bbox from per-rank {min x3, max x3} x2, and the elementwise MAX of the flag words, with a single MIN
    all-reduce (maxima travel negated).  -> (box int32[12], flags int32[N_FLAGS])"""
    b = box.reshape(-1).to(torch.int32)
    if flags is None:
        flags = torch.zeros(N_FLAGS, dtype=torch.int32, device=b.device)
    sign = _SIGN.get(b.device)                 # cached on the device: a pageable H2D copy would wait for the stream
    if sign is None:
        sign = _SIGN[b.device] = torch.tensor([1, 1, 1, -1, -1, -1] * 2 + [-1] * N_FLAGS, dtype=torch.int32, device=b.device)
    packed = torch.cat([b, flags.to(torch.int32)]) * sign
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    packed = packed * sign
    return packed[:12], packed[12:]


def exchange_halo(first_planes: torch.Tensor, rank: int, world: int, group=None):
    """Send this rank's first planes to the rank below, receive the next rank's (None on the last rank):
    neighbour-only point-to-point transfers."""
    ops, recv = [], None
    send = first_planes.contiguous()
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, send, _peer(group, rank - 1), group))
    if rank + 1 < world:
        recv = torch.empty_like(send)
        ops.append(dist.P2POp(dist.irecv, recv, _peer(group, rank + 1), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv


def _piece_bytes(nv, nf):
    return (nv * (8 + 12) + nf * 12 + 7) // 8 * 8


def gather_pieces(pieces, counts, rank: int, world: int, group=None, dst: int = 0):
    """Bring every rank's mesh pieces to rank ``dst``, un-padded.

    pieces: list (one per surface) of (verts [V,3] f32, faces [F,3] i32, keys [V] i64) on this rank;
    counts: host int array [world, n_surfaces, 2] = (V, F) of every rank (already exchanged).
    Returns, on rank ``dst``, a list over ranks of such lists; None elsewhere."""
    dev = pieces[0][0].device
    nbytes = [sum(_piece_bytes(int(counts[r, s_i, 0]), int(counts[r, s_i, 1])) for s_i in range(len(pieces)))
              for r in range(world)]

    def pack():
        buf = torch.empty(max(nbytes[rank], 8), dtype=torch.uint8, device=dev)
        off = 0
        for v, f, k in pieces:                                               # keys first (8-byte aligned)
            for t in (k, v, f):
                n = t.numel() * t.element_size()
                if n:
                    buf[off:off + n].view(t.dtype).copy_(t.reshape(-1))
                off += n
            off = (off + 7) // 8 * 8
        return buf

    def unpack(buf, r):
        off, lst = 0, []
        for s_i in range(len(pieces)):
            nv, nf = int(counts[r, s_i, 0]), int(counts[r, s_i, 1])
            k = buf[off:off + nv * 8].view(torch.int64); off += nv * 8
            v = buf[off:off + nv * 12].view(torch.float32).view(nv, 3); off += nv * 12
            f = buf[off:off + nf * 12].view(torch.int32).view(nf, 3); off += nf * 12
            off = (off + 7) // 8 * 8
            lst.append((v, f, k))
        return lst

    if rank != dst:
        if nbytes[rank]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, pack(), _peer(group, dst), group)]):
                w.wait()
        return None
    recv = {r: torch.empty(nbytes[r], dtype=torch.uint8, device=dev) for r in range(world) if r != dst and nbytes[r]}
    ops = [dist.P2POp(dist.irecv, recv[r], _peer(group, r), group) for r in recv]
    works = dist.batch_isend_irecv(ops) if ops else []
    for w in works:
        w.wait()
    empty = torch.empty(8, dtype=torch.uint8, device=dev)
    return [list(pieces) if r == dst else unpack(recv.get(r, empty), r) for r in range(world)]


def stitch(parts, bounds):
    """parts: per rank (verts, faces, keys) tensors of ONE surface (any device), keys ascending within a rank;
    bounds: per rank the first key of the NEXT rank's slab (its vertices at or beyond that key are this rank's copies
    of vertices the next rank owns; INT64_MAX on the last rank).  Drops the copies, re-indexes the faces that use
    them by key look-up.  Vertex order = key order, face order = rank (== cell) order.
    Returns (verts [V,3] f32, faces [F,3] i32)."""
    dev = parts[0][0].device
    nv = [int(p[2].shape[0]) for p in parts]
    if sum(nv) == 0:
        return torch.zeros((0, 3), dtype=torch.float32, device=dev), torch.zeros((0, 3), dtype=torch.int32, device=dev)
    keys = torch.cat([p[2] for p in parts])
    verts = torch.cat([p[0] for p in parts])
    bound = torch.cat([torch.full((n,), int(b), dtype=torch.int64, device=dev) for n, b in zip(nv, bounds)])
    own = keys < bound
    own_idx = torch.nonzero(own).flatten()
    own_keys = keys[own_idx]                                    # ascending: slabs are ordered, each is sorted
    remap = torch.searchsorted(own_keys, keys)                  # owned: its own rank; copy: the owner's position
    faces, off = [], 0
    for p, n in zip(parts, nv):
        faces.append(remap[p[1].long() + off])
        off += n
    return verts[own_idx], torch.cat(faces).to(torch.int32)


def _grid_host(grid):
    g = grid.detach().cpu()
    return g[0].clone(), g[1:4].clone()


def surface_owner(tag_index: int, world: int, spread: bool) -> int:
    """Rank that stitches, filters and writes surface ``tag_index`` (0 = first requested surface): rank 0, or with
    ``spread`` one surface per rank (hand on 0, object on 1) so that the serial tail is shared."""
    return tag_index % world if spread else 0


def reconstruct_slab(backend: Backend, N: int, rank: int, world: int, hand_branch=True, obj_branch=True,
                     which=("hand", "obj"), group=None, keep_fields=False, spread=False):
    """Both evaluation passes on this rank's slab, halo exchange, marching cubes, gather and stitch.
    Returns dict(grid f32[4] device tensor (voxel, origin; identical on every rank), z0, z1,
    meshes = {tag: (verts, points, faces)} of the surfaces this rank owns (all on rank 0 unless ``spread``; None on
    ranks that own none), and with ``keep_fields`` hand / obj [nz,N,N])."""
    dev = backend.device
    z0, z1 = slab_planes(N, rank, world, backend.relief, backend.relief_ranks)
    nz = z1 - z0
    mask = (1 if hand_branch else 0) | (2 if obj_branch else 0)
    plane = N * N
    while True:
        box, fl1 = backend.pass1(z0 * plane, z1 * plane, mask)
        box, fl1 = reduce_bbox(box, fl1, group)
        grid = backend.regrid(box, mask)
        h, o, fl2 = backend.pass2(z0 * plane, z1 * plane, grid)
        fields = dict(hand=h.view(nz, N, N), obj=o.view(nz, N, N))
        first = torch.stack([fields["hand"][0] if nz else torch.zeros(N, N, device=dev),
                             fields["obj"][0] if nz else torch.zeros(N, N, device=dev)])
        halo = exchange_halo(first, rank, world, group)
        handles, msg = [], []
        tags = [t for t in ("hand", "obj") if t in which]
        for tag in tags:
            vol = fields[tag]
            if halo is not None and nz:
                vol = torch.cat([vol, halo[0 if tag == "hand" else 1][None]], 0)
            if vol.shape[0] >= 2:
                totals, hd = backend.mc_count(vol.contiguous(), grid, z0)
                msg.append(totals.to(torch.int64)[[0, 1, 4]])
            else:
                hd = None
                msg.append(torch.zeros(3, dtype=torch.int64, device=dev))
            handles.append(hd)
        msg.append(torch.maximum(fl1, fl2.to(fl1.device)).to(torch.int64))
        msg = torch.cat(msg)
        allm = [torch.empty_like(msg) for _ in range(world)]
        dist.all_gather(allm, msg, group=group)
        allm = torch.stack(allm).cpu().numpy()                 # the one host synchronisation of the sample
        flags = allm[:, 3 * len(tags):].max(0).tolist()
        if backend.decide is not None and backend.decide(flags):
            continue                                           # every rank saw the same flags: all repeat
        break
    counts = allm[:, :3 * len(tags)].reshape(world, len(tags), 3)
    pieces = []
    for s_i, hd in enumerate(handles):
        nv, nt, nseg = (int(x) for x in counts[rank, s_i])
        if hd is not None and (nv or nt):
            pieces.append(backend.mc_emit(hd, nv, nt, nseg))
        else:
            pieces.append((torch.zeros((0, 3), dtype=torch.float32, device=dev),
                           torch.zeros((0, 3), dtype=torch.int32, device=dev),
                           torch.zeros((0,), dtype=torch.int64, device=dev)))
    out = dict(grid=grid, z0=z0, z1=z1, meshes=None)
    if keep_fields:
        out.update(fields)
    bounds = [slab_planes(N, r, world, backend.relief, backend.relief_ranks)[1] * plane * 4 if r + 1 < world else INT64_MAX
              for r in range(world)]
    owners = [surface_owner(s_i, world, spread) for s_i in range(len(tags))]
    meshes = {}
    for dst in sorted(set(owners)):                             # one un-padded gather per owning rank
        mine = [s_i for s_i, o in enumerate(owners) if o == dst]
        gathered = gather_pieces([pieces[s_i] for s_i in mine], counts[:, mine, :2], rank, world, group, dst)
        if rank == dst:
            for k, s_i in enumerate(mine):
                verts, faces = stitch([gathered[r][k] for r in range(world)], bounds)
                meshes[tags[s_i]] = (verts, grid[1:4].to(verts.device)[None] + verts, faces)
    if meshes:
        out["meshes"] = meshes
    return out


# ----------------------------------------------------------------------------
# product backend + public entry point
# ----------------------------------------------------------------------------
def gpu_backend(bound, N, grid_mode="reference", path=None, spread=False) -> Backend:
    """Kernels of libalignsdf_b200.so on ``bound`` (an engine.BoundSample of ONE sample).  Nothing here waits for
    the GPU; the kernel kind is the decoder's current level, checked through the flags (``decide``).  ``spread``:
    the slab sizes anticipate reconstruct_slab(..., spread=True)."""
    from . import _lib, engine
    dev = bound.device
    mode = engine._GRID_MODES[grid_mode]
    vs1 = 2.0 / (N - 1)
    state = dict(level=None)

    def zero_flags():
        return torch.zeros(N_FLAGS, dtype=torch.int32, device=dev)

    forced = engine._PATH_ALIASES.get(path, path) in engine._PATH_LEVEL or bound.engine.path != "auto"

    def pass1(begin, end, mask):
        state["level"] = lvl = bound.auto_level(path, calibrate=False)
        box = engine.new_bbox(dev)
        auto = not forced and bound.tc_ok and lvl < engine.LEVEL_SIMT
        tau = bound.engine.fast_tau(N) if (auto and not bound.redo_fast) else None
        state["fast"] = tau is not None and end > begin
        if end > begin:
            q = engine.make_query(mode, N, begin, end, vs1, (-1.0, -1.0, -1.0), bbox_mask=mask)
            if lvl >= engine.LEVEL_SIMT:
                bound._launch_simt(q, end - begin, False, box[0])
            elif tau is not None:           # single-product kind + exact re-evaluation of the shell around the surface
                bound.fast_bbox_pass(engine.LEVEL_KIND[lvl], q, end - begin, box, tau, calibrate=True)
            else:
                bound.launch_tc(engine.LEVEL_KIND[lvl], q, end - begin, False, box)
        if auto:
            bound._calibrate()              # (once per sample) queued behind pass 1: its host work overlaps the pass
        return box[0], bound.pending_flags()

    def regrid(box, mask):
        grid = torch.empty((1, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().asdf_regrid(_lib.ptr(box.contiguous()), 1, int(mask), N, float(np.float32(vs1)),
                                              _lib.ptr(grid), None, _lib.stream_ptr(dev)), "asdf_regrid")
        engine.LAUNCHES["count"] += 1
        return grid[0]

    def pass2(begin, end, grid):
        lvl = state["level"]
        n = end - begin
        if n <= 0:
            e = torch.zeros(0, dtype=torch.float32, device=dev)
            return e, e.clone(), zero_flags()
        if lvl >= engine.LEVEL_SIMT:
            g = grid.cpu().tolist()                            # the generic kernel takes its lattice by value
            q = engine.make_query(mode, N, begin, end, g[0], g[1:4])
            h, o, _, _ = bound._launch_simt(q, n, False, None)
            return h, o, bound.pending_flags()
        q = engine.make_query(mode, N, begin, end, 0.0, (0.0, 0.0, 0.0))
        pmax = 2.0                                             # the refit cube lies inside [-1 - 2 voxels, 1 + 2 voxels]
        h, o, _ = bound.launch_tc(engine.LEVEL_KIND[lvl], q, n, True, None, grid.view(1, 4), pmax)
        return h[0], o[0], bound.pending_flags()

    def mc_count(vol, grid, z0):
        return engine.mc_count(vol, 0.0, index0_offset=z0, grid_dev=grid)

    def mc_emit(handle, nv, nt, nseg):
        r = engine.mc_emit(handle, nv, nt, nseg, want_keys=True)
        return r["verts"], r["faces"], r["keys"]

    def decide(flags):
        need = bound.decide(flags)
        if forced:                                             # a forced kind cannot be replaced: keep its results
            return False
        return need > state["level"] or (bound.redo_fast and state.get("fast", False))

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return Backend(pass1, regrid, pass2, mc_count, mc_emit, dev, default_relief(N, world, spread), decide,
                   2 if spread else 1)


def create_mesh_combined_decoder_slab(hand_branch, obj_branch, cls_branch, decoder, latent_vec, mano_results,
                                      obj_results, cam_intr, specs, filename, N=256, group=None,
                                      grid_mode="reference", write=True, spread=False):
    """z-slab sharded equivalent of ``mesh.create_mesh_combined_decoder`` (call on every rank of an
    initialised process group; rank 0 writes the files and returns the meshes).  ``spread=True``: the hand surface
    is stitched, filtered and written by rank 0 and the object surface by rank 1 (same files; each of the two ranks
    returns the mesh it wrote, the other entry stays None) -- halves the serial tail of a sample."""
    import logging
    from . import engine
    from .trimesh_lite import Mesh, largest_watertight_component_mc
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = engine._device_of(latent_vec)
    bound = engine.get_engine(decoder, dev).bind(latent_vec, specs, mano_results, obj_results, cam_intr=cam_intr)
    be = gpu_backend(bound, N, grid_mode, spread=spread)
    which = tuple(t for t, use in (("hand", hand_branch), ("obj", obj_branch)) if use)
    res = reconstruct_slab(be, N, rank, world, hand_branch, obj_branch, which, group, spread=spread)
    if res["meshes"] is None:
        return None
    result = {"hand": None, "obj": None}
    vs = float(res["grid"][0])
    for tag in which:
        if tag not in res["meshes"]:
            continue
        verts_d, points_d, faces_d = res["meshes"][tag]
        if faces_d.shape[0] == 0:
            logging.warning("Cannot reconstruct mesh from '{}'".format(f"{filename}_{tag}.ply"))
            continue
        if points_d.is_cuda:
            # component filter + PLY image on the GPU (csrc/cc.cu), like mesh.convert_sdf_samples_to_ply; the object
            # mesh's "* scale + offset" with the hand's values (utils/mesh.py:366-369) is x * 1 + 0 outside eval_mode
            sel_p, sel_f, _ = engine.select_component(points_d, faces_d, verts_d, (N, N, N), [vs] * 3)
            if write:
                points, faces = engine.export_ply_from_device(f"{filename}_{tag}.ply", sel_p, sel_f)
            else:
                points, faces = sel_p.cpu().numpy(), sel_f.cpu().numpy()
            m = Mesh(points, faces)
        else:                                   # host tensors (gloo tests with the oracle as backend)
            verts, points, faces = (x.cpu().numpy() for x in res["meshes"][tag])
            m = largest_watertight_component_mc(points, faces, verts, (N, N, N), [vs] * 3)
            if write:
                m.export(f"{filename}_{tag}.ply")
        result[tag] = m
    return result
