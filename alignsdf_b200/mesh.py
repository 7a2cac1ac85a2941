"""Drop-in for the reference's ``utils/mesh.py`` (same function names, argument order
and side effects), backed by the CUDA library.

    create_mesh_combined_decoder   <-> utils/mesh.py:17-195
    get_higher_res_cube            <-> utils/mesh.py:198-256
    convert_sdf_samples_to_ply     <-> utils/mesh.py:331-399
    write_verts_label_to_npz       <-> utils/mesh.py:281-297
    write_verts_label_to_obj       <-> utils/mesh.py:258-278       (viz)
    write_color_labeled_ply        <-> utils/mesh.py:300-329       (viz)

Differences, all additive: the functions accept CUDA tensors and keep volumes on the
device; ``create_mesh_combined_decoder`` additionally *returns* the meshes (the reference
returns None and only writes files); ``grid_mode="regular"`` selects the intended
(unsheared) lattice, the default reproduces the reference's lattice bit for bit.
``eval_mode=True`` (what dist_reconstruct.py always passes) aligns every mesh to the
ground-truth mesh on disk with the scale / translation ICP of deep_sdf/metrics/icp_trans_scale.py,
its nearest-neighbour searches running on the GPU (alignsdf_b200/deep_sdf/metrics).
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from . import engine as _engine
from .trimesh_lite import Mesh, export_ply_records, largest_watertight_component_mc, split as _split  # noqa: F401

INT_MAX = 2 ** 31 - 1
DATA_ROOT = "data"          # eval_mode reads <DATA_ROOT>/<task>/test/mesh_{hand,obj}/<id>.obj like utils/mesh.py:386-390
ICP_RNG = None              # numpy Generator for the ICP surface samples (None: fresh entropy, like the reference)


def _bbox_to_minmax(box, hand_branch, obj_branch):
    """int32[12] device bbox -> (min_index, max_index) f32 tensors with the reference's
    conventions: an empty branch contributes (0,0,0)/(0,0,0) (utils/mesh.py:209-211,225-227)
    and the two branches are merged with elementwise min / max (:239-247)."""
    b = box.cpu().tolist()
    mins, maxs = [], []
    for use, o in ((hand_branch, 0), (obj_branch, 6)):
        if not use:
            continue
        if b[o + 3] < 0:
            mins.append(torch.zeros(3)); maxs.append(torch.zeros(3))
        else:
            mins.append(torch.tensor(b[o:o + 3], dtype=torch.float32))
            maxs.append(torch.tensor(b[o + 3:o + 6], dtype=torch.float32))
    if len(mins) == 1:
        return mins[0], maxs[0]
    return torch.min(mins[0], mins[1]), torch.max(maxs[0], maxs[1])


def _regrid(min_index, max_index, N, voxel_size):
    """utils/mesh.py:249-254, same f32 arithmetic on 4 scalars."""
    new_cube_size = (torch.max(max_index - min_index) + 4) * voxel_size
    new_voxel_size = new_cube_size / (N - 1)
    new_origin = (min_index - 2) * voxel_size - 1.0
    return new_voxel_size, new_origin


def get_higher_res_cube(hand_branch, obj_branch, sdf_values_hand, sdf_values_obj, N, voxel_origin,
                        voxel_size):
    """Same contract as the reference; volumes may live on the GPU (the reduction then runs there
    as part of the evaluation kernels; for externally supplied volumes a torch reduction is used)."""
    mins, maxs = [], []
    for use, vol in ((hand_branch, sdf_values_hand), (obj_branch, sdf_values_obj)):
        if not use:
            continue
        idx = torch.nonzero(vol < 0)
        if idx.shape[0] == 0:
            mins.append(torch.zeros(3)); maxs.append(torch.zeros(3))
        else:
            mins.append(idx.min(0).values.float().cpu()); maxs.append(idx.max(0).values.float().cpu())
    mn = mins[0] if len(mins) == 1 else torch.min(mins[0], mins[1])
    mx = maxs[0] if len(maxs) == 1 else torch.max(maxs[0], maxs[1])
    return _regrid(mn, mx, N, voxel_size)


def _as_float_list(x):
    if isinstance(x, torch.Tensor):
        return [float(v) for v in x.reshape(-1).tolist()]
    return [float(v.item() if isinstance(v, torch.Tensor) else v) for v in x]


def convert_sdf_samples_to_ply(pytorch_3d_sdf_tensor, voxel_grid_origin, voxel_size, ply_filename_out,
                               offset=None, scale=None, eval_mode=False, task='obman',
                               return_mesh=False, raw_on_device=False):
    """Marching cubes (GPU) -> origin shift -> optional scale/offset -> largest watertight
    component if the mesh splits -> PLY.  Returns (verts, faces, trans, scale) like the reference
    (raw marching-cubes vertices, i.e. before the origin shift and the component filter); with
    ``raw_on_device`` those two stay CUDA tensors (create_mesh_combined_decoder only needs them on the
    host for the label pass)."""
    vol = pytorch_3d_sdf_tensor
    if not isinstance(vol, torch.Tensor):
        vol = torch.as_tensor(np.asarray(vol))
    if not vol.is_cuda:
        vol = vol.to(torch.device("cuda", torch.cuda.current_device()))
    vs = float(voxel_size.item() if isinstance(voxel_size, torch.Tensor) else voxel_size)
    origin = _as_float_list(voxel_grid_origin)
    try:
        out = _engine.marching_cubes(vol, level=0.0, spacing=[vs] * 3, origin=origin)
        if out["faces"].shape[0] == 0:
            raise RuntimeError("No surface found at the given iso value.")
    except (ValueError, RuntimeError) as e:
        logging.warning("Cannot reconstruct mesh from '{}'".format(ply_filename_out))
        print(e)
        res = (None, None, np.array([0, 0, 0]), np.array([1]))
        return res + (None,) if return_mesh else res
    # trimesh.graph.split + "largest area piece if more than one" (:371-381) on the GPU (csrc/cc.cu); only the
    # kept component and the raw marching-cubes arrays the reference returns travel to the host
    sel_points, sel_faces, _ = _engine.select_component(out["points"], out["faces"], out["verts"], vol.shape, [vs] * 3)
    whole = sel_faces is out["faces"]
    # x * 1 + 0 (what the reference applies to the object mesh outside eval_mode, utils/mesh.py:186-194) is the identity
    identity = ((scale is None or bool(np.all(np.asarray(scale) == 1)))
                and (offset is None or not bool(np.any(np.asarray(offset)))))
    if identity:
        # vertex block + face records -> one pinned buffer -> one write; origin + verts (f32), :360-363
        mesh_points, sel_faces_np = _engine.export_ply_from_device(ply_filename_out, sel_points, sel_faces,
                                                                   write=not eval_mode)
    else:
        ply_faces = _engine.ply_face_records(sel_faces).cpu().numpy()
        mesh_points = sel_points.cpu().numpy()
        sel_faces_np = sel_faces.cpu().numpy()
        if scale is not None:
            mesh_points = mesh_points * scale
        if offset is not None:
            mesh_points = mesh_points + offset
        if not eval_mode:
            export_ply_records(ply_filename_out, mesh_points, ply_faces)
    if raw_on_device:
        verts, faces = out["verts"], out["faces"]
    else:
        verts = out["verts"].cpu().numpy()
        faces = sel_faces_np if whole else out["faces"].cpu().numpy()
    source_mesh = Mesh(mesh_points, sel_faces_np)
    trans, scale = np.array([0, 0, 0]), np.array([1])
    if eval_mode:
        # utils/mesh.py:385-395: scale / translation ICP of the mesh onto the ground-truth mesh of the sample, the
        # ALIGNED mesh is what gets written; both nearest-neighbour searches run on the GPU (csrc/nn.cu)
        from .deep_sdf.metrics.icp_trans_scale import ICP_T_S
        from .trimesh_lite import load as _load
        mesh_dir = 'mesh_' + ply_filename_out.split('_')[-1].split('.')[0]
        gt_mesh_name = ply_filename_out.split('/')[-1].split('_')[0] + '.obj'
        gt_mesh_path = os.path.join(f'{DATA_ROOT}/{task}/test', mesh_dir, gt_mesh_name)
        target_mesh = _load(gt_mesh_path, process=False)
        icp_solver = ICP_T_S(source_mesh, target_mesh, device=vol.device)
        icp_solver.sample_mesh(30000, 'both', ICP_RNG)
        icp_solver.run_icp_f(max_iter=100)
        icp_solver.export_source_mesh(ply_filename_out)
        trans, scale = icp_solver.get_trans_scale()
    res = (verts, faces, trans, scale)
    return res + (source_mesh,) if return_mesh else res


def write_verts_label_to_npz(pytorch_3d_xyz_tensor, pytorch_label_tensor, npz_filename_out,
                             offset=None, scale=None):
    pts = pytorch_3d_xyz_tensor.data.cpu().numpy()
    labels = pytorch_label_tensor.cpu().numpy()
    if scale is not None:
        pts = pts * scale
    if offset is not None:
        pts = pts + offset
    np.savez(npz_filename_out, points=pts, labels=labels)


# colours of the six hand-part labels in the label visualisation (utils/mesh.py:315-320)
_PART_COLOR = np.array([[13, 212, 128], [250, 70, 42], [131, 66, 37], [78, 137, 54], [187, 246, 163], [67, 220, 74]],
                       dtype=np.uint8)


def _label_points(pytorch_3d_xyz_tensor, offset, scale):
    pts = pytorch_3d_xyz_tensor.data.cpu().numpy()
    if scale is not None:
        pts = pts * scale
    if offset is not None:
        pts = pts + offset
    return pts


def write_verts_label_to_obj(pytorch_3d_xyz_tensor, pytorch_label_tensor, obj_filename_out, offset=None, scale=None):
    """``viz`` output of the label pass (utils/mesh.py:258-278): one ``v x y z g g g`` line per marching-cubes vertex,
    grey level = 45 x label."""
    pts = _label_points(pytorch_3d_xyz_tensor, offset, scale)
    grey = pytorch_label_tensor.cpu().numpy() * 45.0
    with open(obj_filename_out, "w") as fp:
        fp.write("".join("v %.4f %.4f %.4f %.2f %.2f %.2f\n" % (p[0], p[1], p[2], c, c, c) for p, c in zip(pts, grey)))


def write_color_labeled_ply(pytorch_3d_xyz_tensor, numpy_faces, pytorch_label_tensor, ply_filename_out, offset=None,
                            scale=None):
    """``viz`` output of the label pass (utils/mesh.py:300-329 + utils/customized_export_ply.py): ASCII PLY of the raw
    marching-cubes mesh with one RGBA colour per vertex (the part colour of its label, alpha 255)."""
    pts = _label_points(pytorch_3d_xyz_tensor, offset, scale)
    rgb = _PART_COLOR[pytorch_label_tensor.cpu().numpy().astype(np.int32)]
    faces = np.asarray(numpy_faces.cpu().numpy() if isinstance(numpy_faces, torch.Tensor) else numpy_faces).reshape(-1, 3)
    head = ("ply\nformat ascii 1.0\n"
            f"element vertex {len(pts)}\nproperty float x\nproperty float y\nproperty float z\n"
            "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
            f"element face {len(faces)}\nproperty list uchar int vertex_indices\nend_header\n")
    with open(ply_filename_out, "w") as fp:
        fp.write(head)
        fp.write("".join("%f %f %f %d %d %d 255\n" % (p[0], p[1], p[2], c[0], c[1], c[2]) for p, c in zip(pts, rgb)))
        fp.write("".join("3 %d %d %d\n" % (f[0], f[1], f[2]) for f in faces))


def _two_pass_verified(bound, N, mask, grid_mode, keep_pass1, path=None):
    """bound.two_pass + the host check of its speculative parts (calibration, operand-range flags, threshold of
    the fast bounding-box pass); re-runs through the next safer kernel / the exact pass when they say so."""
    lvl, r = bound.auto_level(path, calibrate=False), None
    auto = (_engine._PATH_ALIASES.get(path, path) if path else bound.engine.path) == "auto"
    while True:
        if r is None:
            r = bound.two_pass(N, mask, grid_mode, lvl, keep_pass1, calibrate=auto)
        need = bound.verify()
        if need <= lvl and not (bound.redo_fast and r.get("fast_bbox")):
            bound.last_kind = _engine.LEVEL_NAMES[lvl]
            return r
        lvl, r = max(need, lvl), None          # rejected kind, or a fast bounding-box pass whose threshold was too small


def _volumes_from_two_pass(r, bound, N):
    shp = (N, N, N)
    g = r["grid"][0].cpu()
    mm = r["minmax"][0].cpu()
    view = lambda t: None if t is None else t[0].view(shp)
    return dict(pass1_hand=view(r["pass1_hand"]), pass1_obj=view(r["pass1_obj"]), hand=view(r["hand"]),
                obj=view(r["obj"]), cls=None, voxel=g[0].clone(), origin=g[1:4].clone(),
                min_index=mm[:3].clone(), max_index=mm[3:].clone(), bound=bound)


def sdf_volumes(decoder, latent_vec, mano_results, obj_results, specs, N, hand_branch=True,
                obj_branch=True, cls_branch=False, device=None, grid_mode="reference", path=None, bound=None,
                keep_pass1=True, cam_intr=None):
    """The two evaluation passes of utils/mesh.py:24-120 on the GPU.

    Returns dict(pass1_hand, pass1_obj, hand, obj, cls, voxel (0-dim f32 tensor),
    origin (f32[3] tensor), bound) with [N,N,N] CUDA volumes (pass-1 volumes only with ``keep_pass1``:
    the reference uses them for nothing but the bounding box)."""
    dev = _engine._device_of(latent_vec, device)
    eng = _engine.get_engine(decoder, dev)
    if bound is None:                    # ``bound``: a sample already bound by the caller (pipelined batches)
        bound = eng.bind(latent_vec, specs, mano_results, obj_results, cam_intr=cam_intr)
    voxel_size = 2.0 / (N - 1)
    mask = (1 if hand_branch else 0) | (2 if obj_branch else 0)
    if mask == 0:
        raise ValueError("at least one of hand_branch / obj_branch must be set")
    shp = (N, N, N)
    want_cls = cls_branch and eng.topo.classifier is not None
    if bound.tc_ok and not want_cls:
        # tensor-core kernel: pass 1 -> asdf_regrid -> pass 2 without a host round trip in between
        return _volumes_from_two_pass(_two_pass_verified(bound, N, mask, grid_mode, keep_pass1, path), bound, N)
    h1, o1, _, box = bound.eval_grid(N, voxel_size, [-1.0, -1.0, -1.0], grid_mode, bbox_mask=mask,
                                     path=path)
    mn, mx = _bbox_to_minmax(box, hand_branch, obj_branch)
    new_voxel_size, new_origin = _regrid(mn, mx, N, voxel_size)
    h2, o2, c2, _ = bound.eval_grid(N, float(new_voxel_size), new_origin.tolist(), grid_mode,
                                    want_cls=want_cls, path=path)
    return dict(pass1_hand=h1.view(shp), pass1_obj=None if o1 is None else o1.view(shp),
                hand=h2.view(shp), obj=None if o2 is None else o2.view(shp),
                cls=None if c2 is None else c2.view(shp), voxel=new_voxel_size, origin=new_origin,
                min_index=mn, max_index=mx, bound=bound)


def create_mesh_combined_decoder(hand_branch, obj_branch, cls_branch, decoder, latent_vec, mano_results,
                                 obj_results, cam_intr, specs, filename, N=256, max_batch=32 ** 3,
                                 offset=None, scale=None, device="cpu", label_out=False, viz=False,
                                 eval_mode=False, task='obman', grid_mode="reference"):
    """Same call as the reference (``max_batch`` is accepted and ignored: the whole grid is one
    launch).  Writes ``<filename>_hand.ply`` / ``_obj.ply`` (and ``_hand_label.npz`` with
    ``label_out``) and returns ``{"hand": mesh or None, "obj": mesh or None}``."""
    ply_filename_hand = filename + "_hand"
    ply_filename_obj = filename + "_obj"
    decoder.eval()
    vols = sdf_volumes(decoder, latent_vec, mano_results, obj_results, specs, N, hand_branch,
                       obj_branch, cls_branch, None if device == "cpu" else device, grid_mode, keep_pass1=False,
                       cam_intr=cam_intr)
    voxel_size = vols["voxel"]
    voxel_origin = vols["origin"].tolist()
    result = {"hand": None, "obj": None}
    if hand_branch:
        vertices, mesh_faces, offset, scale, mesh = convert_sdf_samples_to_ply(
            vols["hand"], voxel_origin, voxel_size, ply_filename_hand + ".ply", None, None, eval_mode,
            task, return_mesh=True, raw_on_device=True)
        result["hand"] = mesh
        if label_out and (vertices is not None):
            # utils/mesh.py:137-184: re-query the decoder at the marching-cubes vertices
            v = vertices.cpu().clone()
            for k in range(3):
                v[:, k] = voxel_origin[k] + v[:, k]
            bound = vols["bound"]
            if bound.engine.topo.classifier is None:
                raise IndexError("label_out needs a decoder with a classifier head "
                                 "(the reference fails the same way, utils/mesh.py:157)")
            _, _, cls = bound.eval_points(v.to(bound.device), want_cls=True)
            out_labels = cls.float().cpu()
            if viz:
                write_verts_label_to_obj(v, out_labels, ply_filename_hand + "_label.obj", offset, scale)
                write_color_labeled_ply(v, mesh_faces, out_labels, ply_filename_hand + "_color.ply", offset, scale)
            write_verts_label_to_npz(v, out_labels, ply_filename_hand + "_label.npz", offset, scale)
    if obj_branch:
        # the object mesh reuses the HAND call's offset/scale (utils/mesh.py:186-194)
        *_, mesh = convert_sdf_samples_to_ply(vols["obj"], voxel_origin, voxel_size,
                                              ply_filename_obj + ".ply", offset, scale, False,
                                              return_mesh=True, raw_on_device=True)
        result["obj"] = mesh
    return result


def create_meshes_pipelined(decoder, samples, filenames, N=256, hand_branch=True, obj_branch=True,
                            grid_mode="reference", device=None):
    """Batch form of ``create_mesh_combined_decoder`` for loops like reconstruct.py:70-93 (one call per test
    image): the two grid passes of sample i+1 run while a worker thread extracts, filters, reads back and
    writes the meshes of sample i on a second CUDA stream, so the host-side stages no longer leave the GPU
    idle between samples.  ``samples``: iterable of objects with ``latent``, ``mano_results``,
    ``obj_results``, ``specs`` and, for PixelAlign samples, ``cam_intr`` (host or CUDA tensors); ``filenames``:
    output prefixes.  Same files and
    meshes as the one-call-per-sample loop; returns the list of ``{"hand": mesh, "obj": mesh}`` dicts."""
    import queue
    import threading
    decoder.eval()
    results, errors = {}, []
    work = queue.Queue(maxsize=2)                      # at most two samples' volumes alive

    def finish():
        stream = None
        while True:
            item = work.get()
            if item is None:
                return
            idx, vols, done, prefix = item
            try:
                dev = vols["hand"].device
                if stream is None:
                    stream = torch.cuda.Stream(dev)
                with torch.cuda.device(dev), torch.cuda.stream(stream):
                    stream.wait_event(done)
                    out = {"hand": None, "obj": None}
                    voxel_origin = vols["origin"].tolist()
                    offset = scale = None
                    if hand_branch:
                        _, _, offset, scale, out["hand"] = convert_sdf_samples_to_ply(
                            vols["hand"], voxel_origin, vols["voxel"], prefix + "_hand.ply", None, None, False,
                            return_mesh=True, raw_on_device=True)
                    if obj_branch:
                        *_, out["obj"] = convert_sdf_samples_to_ply(
                            vols["obj"], voxel_origin, vols["voxel"], prefix + "_obj.ply", offset, scale, False,
                            return_mesh=True, raw_on_device=True)
                results[idx] = out
            except Exception as e:                     # surfaced on the caller's thread
                errors.append(e)
                results[idx] = None

    worker = threading.Thread(target=finish, daemon=True)
    worker.start()
    from concurrent.futures import ThreadPoolExecutor
    samples, filenames = list(samples), list(filenames)
    bind_stream = {}

    def bind(i):
        """Fold + pack + upload sample i (host work ~3 ms) on its own stream, ahead of its turn."""
        smp = samples[i]
        dev = _engine._device_of(smp.latent, device)
        if dev not in bind_stream:
            bind_stream[dev] = torch.cuda.Stream(dev)
        with torch.cuda.device(dev), torch.cuda.stream(bind_stream[dev]):
            to = lambda t: t.to(dev, non_blocking=True)
            latent = to(smp.latent)
            mano = None if smp.mano_results is None else {k: to(v) for k, v in smp.mano_results.items()}
            obj = None if smp.obj_results is None else {k: to(v) for k, v in smp.obj_results.items()}
            cam = getattr(smp, "cam_intr", None)         # PixelAlign samples
            bound = _engine.get_engine(decoder, dev).bind(latent, smp.specs, mano, obj,
                                                          cam_intr=None if cam is None else to(cam))
            lvl = bound.auto_level()                     # starts the calibration run as well
            if lvl < _engine.LEVEL_SIMT:
                bound.tc_blocks(_engine.LEVEL_KIND[lvl], 2.0)    # P tiles built on the device, ahead of the sample's turn
            ready = torch.cuda.Event()
            ready.record(bind_stream[dev])
        return dev, latent, mano, obj, bound, ready

    n = 0
    mask = (1 if hand_branch else 0) | (2 if obj_branch else 0)

    side = {}

    def finish_passes(p):
        """Host check of a sample whose two passes were queued one sample ago (the GPU is already busy with the
        next one): its flags and lattice are read on a side stream that only waits for THAT sample's kernels.
        Then its volumes go to the mesh worker."""
        idx, dev, bound, lvl, r, prefix, done, small = p
        with torch.cuda.device(dev):
            if r is not None:
                if dev not in side:
                    side[dev] = torch.cuda.Stream(dev)
                with torch.cuda.stream(side[dev]):
                    side[dev].wait_event(done)
                    host = small.cpu().tolist()
                need = bound.decide([int(x) for x in host[:5]])
                if need <= lvl and not (bound.redo_fast and r.get("fast_bbox")):
                    bound.last_kind = _engine.LEVEL_NAMES[lvl]
                    g, mm = torch.tensor(host[5:9], dtype=torch.float32), torch.tensor(host[9:15], dtype=torch.float32)
                    view = lambda t: t[0].view(N, N, N)
                    vols = dict(hand=view(r["hand"]), obj=view(r["obj"]), voxel=g[0].clone(), origin=g[1:4].clone(),
                                min_index=mm[:3].clone(), max_index=mm[3:].clone(), bound=bound)
                else:                                    # rejected: the same sample through the safer kernel, in order
                    r = None
            if r is None:                                # also: decoders / queries outside the tensor-core path
                smp = samples[idx]
                vols = sdf_volumes(decoder, bound.inputs[0][0], bound.inputs[0][2], bound.inputs[0][3], smp.specs, N,
                                   hand_branch, obj_branch, False, dev, grid_mode, bound=bound, keep_pass1=False)
                vols = {k: vols[k] for k in ("hand", "obj", "voxel", "origin", "bound")}
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(dev))
        work.put((idx, vols, done, prefix))

    try:
        with ThreadPoolExecutor(max_workers=1) as binder:
            nxt = binder.submit(bind, 0) if samples else None
            pending = None
            for idx, prefix in enumerate(filenames[:len(samples)]):
                if errors:
                    break
                dev, latent, mano, obj, bound, ready = nxt.result()
                nxt = binder.submit(bind, idx + 1) if idx + 1 < len(samples) else None
                torch.cuda.current_stream(dev).wait_event(ready)
                lvl, r, done, small = bound.auto_level(), None, None, None
                tc = bound.tc_ok and lvl < _engine.LEVEL_SIMT
                if tc:
                    with torch.cuda.device(dev):
                        ctx = bound.two_pass_begin(N, mask, grid_mode, lvl, False)   # pass 1, queued behind the previous sample's passes
                if pending is not None:
                    finish_passes(pending)           # the previous sample's volumes go to the mesh worker meanwhile
                if tc:
                    with torch.cuda.device(dev):
                        r = bound.two_pass_end(ctx)  # (reads the fast pass's list sizes,) re-grid, pass 2
                        # flags (exact in f64: int32 words / f32 bit patterns), lattice and bbox in one small tensor
                        small = torch.cat([bound.pending_flags().double(), r["grid"][0].double(), r["minmax"][0].double()])
                        done = torch.cuda.Event()
                        done.record(torch.cuda.current_stream(dev))
                pending = (idx, dev, bound, lvl, r, prefix, done, small)
                n += 1
            if pending is not None and not errors:
                finish_passes(pending)
    finally:
        work.put(None)
        worker.join()
    if errors:
        raise errors[0]
    return [results[i] for i in range(n)]
