"""Drop-in for deep_sdf/metrics/icp_trans_scale.py: ICP restricted to a scale and a translation between points
sampled on a source (predicted) and a target (ground-truth) mesh -- what ``--eval_mode`` runs after every mesh
(utils/mesh.py:385-395; dist_reconstruct.py:9-13 always passes --eval_mode).

Same class, methods, attributes and stopping rules as the reference.  Differences in mechanism only: both
nearest-neighbour queries of an iteration are exact float64 brute-force searches on the GPU (asdf_nn_search; the
reference queries two sklearn KD-trees on the CPU), the point clouds stay on the device for the whole loop, and the
4-unknown least-squares problem (icp_trans_scale.py:71-102) is solved from its normal equations -- ten sums reduced on
the device -- instead of an SVD of the [6 n, 4] matrix.  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _lib
from ...trimesh_lite import sample_surface


def nn_search(query: torch.Tensor, ref: torch.Tensor, want_dist=False):
    """Index (int64 CUDA tensor) of the nearest ``ref`` point for every ``query`` point, and the squared distances
    when asked.  Both CUDA float64 [*, 3]."""
    _lib.require_cuda(query, "query")
    q, r = query.to(torch.float64).contiguous(), ref.to(torch.float64).contiguous()
    idx = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
    d2 = torch.empty(q.shape[0], dtype=torch.float64, device=q.device) if want_dist else None
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().asdf_nn_search(_lib.ptr(q), q.shape[0], _lib.ptr(r), r.shape[0], _lib.ptr(idx),
                                             _lib.ptr(d2), _lib.stream_ptr(q.device)), "asdf_nn_search")
    return (idx.long(), d2) if want_dist else idx.long()


class ICP_T_S():
    def __init__(self, mesh_source, mesh_target, device=None):
        self.mesh_source = mesh_source
        self.mesh_target = mesh_target
        self.points_source = np.array(self.mesh_source.vertices, dtype=np.float64)
        self.points_target = np.array(self.mesh_target.vertices, dtype=np.float64)
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.AsdfError("ICP_T_S needs a CUDA device (B200); there is no CPU path")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)

    def sample_mesh(self, n=30000, mesh_id='both', rng=None):
        """icp_trans_scale.py:19-30: sample both surfaces, then move the source cloud onto the target's centroid
        and RMS radius."""
        if mesh_id == 'source' or mesh_id == 'both':
            self.points_source, _ = sample_surface(self.mesh_source, n, rng)
        if mesh_id == 'target' or mesh_id == 'both':
            self.points_target, _ = sample_surface(self.mesh_target, n, rng)
        self.normalize_points()

    def normalize_points(self):
        self.offset_source = self.points_source.mean(0)
        self.scale_source = np.sqrt(((self.points_source - self.offset_source) ** 2).sum() / len(self.points_source))
        self.offset_target = self.points_target.mean(0)
        self.scale_target = np.sqrt(((self.points_target - self.offset_target) ** 2).sum() / len(self.points_target))
        self.points_source = (self.points_source - self.offset_source) / self.scale_source * self.scale_target + self.offset_target

    # ------------------------------------------------------------------
    def _solve(self, ps, pt, idx_s, closest_target):
        """icp_trans_scale.py:71-102: least squares for (scale, t) over the 3 (n_s + n_t) equations
        scale * a_k + t_{c(k)} = b_k with a = [source points; matched source points], b = [matched target points;
        target points].  Normal equations from ten sums."""
        a = torch.cat([ps, ps[idx_s]])
        b = torch.cat([closest_target, pt])
        sums = torch.cat([(a * a).sum().reshape(1), (a * b).sum().reshape(1), a.sum(0), b.sum(0)]).cpu().numpy()
        n = float(a.shape[0])
        saa, sab, sa, sb = sums[0], sums[1], sums[2:5], sums[5:8]
        M = np.zeros((4, 4))
        M[0, 0], M[0, 1:], M[1:, 0] = saa, sa, sa
        M[1:, 1:] = np.eye(3) * n
        x = np.linalg.solve(M, np.concatenate([[sab], sb]))
        return x[0:1].copy(), x[1:].reshape(1, 3).copy()

    def _run(self, max_iter, stop_error, stop_improvement, verbose, rebuild):
        dev = self.device
        ps = torch.as_tensor(self.points_source, dtype=torch.float64).to(dev)
        pt = torch.as_tensor(self.points_target, dtype=torch.float64).to(dev)
        self.trans = np.zeros((1, 3), dtype=np.float64)
        self.scale = 1.0
        ntot = ps.shape[0] + pt.shape[0]
        error = 1e8
        previous_error = error
        self.errors = []
        for i in range(0, max_iter):
            scale = float(np.asarray(self.scale).reshape(-1)[0])
            trans = torch.as_tensor(self.trans, dtype=torch.float64).to(dev)
            # closest target point for each source point
            query_source = ps * scale + trans
            idx_t = nn_search(query_source, pt)
            closest_target = pt[idx_t]
            # closest source point for each target point
            if rebuild:                               # run_icp: tree over the transformed source cloud
                idx_s = nn_search(pt, query_source)
            else:                                     # run_icp_f: target points pulled back into the source frame
                idx_s = nn_search((pt - trans) / scale, ps)
            closest_source = ps[idx_s] * scale + trans
            error = float(torch.sqrt((((query_source - closest_target) ** 2).sum() + ((pt - closest_source) ** 2).sum()) / ntot))
            self.errors.append(error)
            if verbose >= 1:
                print(i, "th iter, error: ", error)
            if not rebuild:
                if previous_error - error < stop_improvement:
                    break
                else:
                    previous_error = error
            if error < stop_error:
                break
            self.scale, self.trans = self._solve(ps, pt, idx_s, closest_target)

    def run_icp_f(self, max_iter=10, stop_error=1e-3, stop_improvement=1e-5, verbose=0):
        """icp_trans_scale.py:32-113 (both KD-trees built once)."""
        self._run(max_iter, stop_error, stop_improvement, verbose, rebuild=False)

    def run_icp(self, max_iter=10, stop_error=1e-3):
        """icp_trans_scale.py:115-186 (source tree rebuilt from the transformed cloud in every iteration)."""
        self._run(max_iter, stop_error, 0.0, 0, rebuild=True)

    def get_trans_scale(self):
        all_scale = self.scale_target * self.scale / self.scale_source
        all_trans = self.trans + self.offset_target * self.scale - self.offset_source * self.scale_target * self.scale / self.scale_source
        return all_trans, all_scale

    def export_source_mesh(self, output_name):
        self.mesh_source.vertices = (self.mesh_source.vertices - self.offset_source) / self.scale_source * self.scale_target + self.offset_target
        self.mesh_source.vertices = self.mesh_source.vertices * self.scale + self.trans
        self.mesh_source.export(output_name)
