"""CPU restatement of the step after the hot path: scale / translation ICP and symmetric Chamfer distance.
TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Follows deep_sdf/metrics/icp_trans_scale.py:19-30 (normalisation), :32-113 (run_icp_f), :188-191 (get_trans_scale)
and deep_sdf/metrics/chamfer.py:212-231, with the same libraries the reference uses (sklearn.neighbors.KDTree,
numpy.linalg.lstsq).  Pinned: oracle/make_golden_icp.py runs the reference's own, unmodified ICP_T_S class on seeded
point clouds and asserts this restatement reproduces its scale / translation / error trace (tests/golden/icp_*.npz).
"""
from __future__ import annotations

import numpy as np
from sklearn.neighbors import KDTree


def normalize(points_source, points_target):
    """icp_trans_scale.py:25-30 -> (source moved onto the target's centroid / RMS radius, offsets, scales)"""
    offset_source = points_source.mean(0)
    scale_source = np.sqrt(((points_source - offset_source) ** 2).sum() / len(points_source))
    offset_target = points_target.mean(0)
    scale_target = np.sqrt(((points_target - offset_target) ** 2).sum() / len(points_target))
    moved = (points_source - offset_source) / scale_source * scale_target + offset_target
    return moved, (offset_source, scale_source, offset_target, scale_target)


def run_icp_f(points_source, points_target, max_iter=10, stop_error=1e-3, stop_improvement=1e-5):
    """icp_trans_scale.py:32-113 -> (scale (1,), trans (1,3), [error per iteration])"""
    target_tree = KDTree(points_target)
    source_tree = KDTree(points_source)
    trans = np.zeros((1, 3), dtype=np.float64)
    scale = 1.0
    ns, nt = len(points_source), len(points_target)
    ones = np.tile(np.eye(3), (ns + nt, 1))                     # the A_c1 | A_c2 | A_c3 indicator columns (:83-92)
    errors = []
    previous_error = 1e8
    for _ in range(max_iter):
        query_source = points_source * scale + trans
        _, it = target_tree.query(query_source)
        closest_target = points_target[it[:, 0], :]
        query_target = (points_target - trans) / scale
        _, isrc = source_tree.query(query_target)
        closest_source = points_source[isrc[:, 0], :] * scale + trans
        error = ((((query_source - closest_target) ** 2).sum() + ((points_target - closest_source) ** 2).sum())
                 / (ns + nt)) ** 0.5
        errors.append(float(error))
        if previous_error - error < stop_improvement:
            break
        previous_error = error
        if error < stop_error:
            break
        A = np.hstack([np.vstack([points_source.reshape(-1, 1), points_source[isrc[:, 0], :].reshape(-1, 1)]), ones])
        b = np.vstack([closest_target.reshape(-1, 1), points_target.reshape(-1, 1)])
        x = np.linalg.lstsq(A, b, rcond=-1)
        scale = x[0][0]
        trans = (x[0][1:]).transpose()
    return np.asarray(scale, np.float64).reshape(-1), np.asarray(trans, np.float64).reshape(1, 3), errors


def get_trans_scale(scale, trans, norm):
    """icp_trans_scale.py:188-191"""
    offset_source, scale_source, offset_target, scale_target = norm
    all_scale = scale_target * scale / scale_source
    all_trans = trans + offset_target * scale - offset_source * scale_target * scale / scale_source
    return all_trans, all_scale


def chamfer(points_source, points_target):
    """chamfer.py:212-231 (metres in, cm^2 out)"""
    a, b = points_source * 100.0, points_target * 100.0
    one, _ = KDTree(a).query(b)
    two, _ = KDTree(b).query(a)
    return float(np.mean(np.square(one)) + np.mean(np.square(two)))
