"""Drop-in for deep_sdf/metrics/chamfer.py:183-231: symmetric Chamfer distance between a predicted and a
ground-truth mesh (sum of the two mean squared nearest-neighbour distances, in cm^2), optionally after the
scale / translation ICP.  The two cKDTree queries are exact float64 brute-force searches on the GPU."""
from __future__ import annotations

import warnings

import numpy as np
import torch

from ... import trimesh_lite
from .icp_trans_scale import ICP_T_S, nn_search


def chamfer_points(points_source, points_target, device=None):
    """chamfer.py:212-231 on two point arrays [*, 3] (metres): -> gt_to_gen + gen_to_gt in cm^2."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    a = torch.as_tensor(np.asarray(points_source, np.float64) * 100.0).to(device)
    b = torch.as_tensor(np.asarray(points_target, np.float64) * 100.0).to(device)
    _, d_one = nn_search(b, a, want_dist=True)        # target -> generated
    _, d_two = nn_search(a, b, want_dist=True)        # generated -> target
    return float(d_one.mean() + d_two.mean())


def compute_trimesh_chamfer(gt_mesh_filename, pred_mesh_filename, optim=False, rot=False, rng=None):
    warnings.filterwarnings("ignore")
    source_mesh = trimesh_lite.load(pred_mesh_filename, process=False)
    target_mesh = trimesh_lite.load(gt_mesh_filename, process=False)
    if optim:
        if rot:
            raise NotImplementedError("rot=True uses trimesh.registration.icp (rigid Procrustes ICP), not built")
        icp_solver = ICP_T_S(source_mesh, target_mesh)
        icp_solver.sample_mesh(30000, 'both', rng)
        icp_solver.run_icp_f(max_iter=100)
        points_source = icp_solver.points_source * icp_solver.scale + icp_solver.trans
        points_target = icp_solver.points_target
    else:
        points_source, _ = trimesh_lite.sample_surface(source_mesh, 30000, rng)
        points_target, _ = trimesh_lite.sample_surface(target_mesh, 30000, rng)
    return chamfer_points(points_source, points_target)
