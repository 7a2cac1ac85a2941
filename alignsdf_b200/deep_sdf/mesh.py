"""Drop-in for ``deep_sdf/mesh.py`` (the legacy DeepSDF routine named by the north star).

    create_mesh(decoder, latent_vec, filename, N=256, max_batch=32**3)   <-> deep_sdf/mesh.py:14-61
    convert_sdf_samples_to_ply(t, origin, voxel, out)                    <-> deep_sdf/mesh.py:64-116

The decoder is a single-output DeepSDF MLP (state-dict keys ``lin{i}.*``, one output); a
two-output AlignSDF ``CombinedDecoder`` is accepted too and its first (hand) output is used.
Single pass over [-1,1]^3, no component filtering, binary PLY like the plyfile writer.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .. import engine as _engine
from ..trimesh_lite import Mesh, export_ply


def create_mesh(decoder, latent_vec, filename, N=256, max_batch=32 ** 3, grid_mode="reference"):
    start = time.time()
    decoder.eval()
    inner = _engine.unwrap_decoder(decoder)
    dev = _engine._device_of(latent_vec)
    eng = _engine.get_engine(inner, dev)
    specs = dict(PointFeatSize=3, EncodeStyle="nerf", SdfScaleFactor=1.0, PixelAlign=False)
    bound = eng.bind(latent_vec, specs, None, None)
    voxel_origin = [-1, -1, -1]
    voxel_size = 2.0 / (N - 1)
    sdf, _, _, _ = bound.eval_grid(N, voxel_size, [-1.0, -1.0, -1.0], grid_mode)
    torch.cuda.synchronize(dev)
    print("sampling takes: %f" % (time.time() - start))
    return convert_sdf_samples_to_ply(sdf.view(N, N, N), voxel_origin, voxel_size, filename + ".ply")


def convert_sdf_samples_to_ply(pytorch_3d_sdf_tensor, voxel_grid_origin, voxel_size, ply_filename_out):
    """Unlike utils/mesh.py the legacy routine lets the marching-cubes ValueError propagate."""
    vol = pytorch_3d_sdf_tensor
    if not isinstance(vol, torch.Tensor):
        vol = torch.as_tensor(np.asarray(vol))
    if not vol.is_cuda:
        vol = vol.to(torch.device("cuda", torch.cuda.current_device()))
    out = _engine.marching_cubes(vol, 0.0, [float(voxel_size)] * 3,
                                 [float(v) for v in voxel_grid_origin])
    pts = out["points"].cpu().numpy()
    faces = out["faces"].cpu().numpy()
    export_ply(ply_filename_out, pts, faces)
    return Mesh(pts, faces)
