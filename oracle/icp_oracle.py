"""CPU restatement of the step after the hot path: scale / translation ICP and symmetric Chamfer distance.
TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Follows deep_sdf/metrics/icp_trans_scale.py:19-30 (normalisation), :32-113 (run_icp_f), :188-191 (get_trans_scale)
and deep_sdf/metrics/chamfer.py:13-180 (alignment helpers), :212-231 (distance), with the same libraries the reference
uses (sklearn.neighbors.KDTree, scipy cKDTree, numpy.linalg).  Pinned: oracle/make_golden_icp.py runs the reference's
own, unmodified ICP_T_S class and chamfer.py functions on seeded point clouds and asserts this restatement reproduces
them (tests/golden/icp_*.npz, tests/golden/align_*.npz).

``registration_icp`` restates ``trimesh.registration.icp`` (third party, unpinned in requirements.txt:3, absent here;
called at chamfer.py:203 for ``rot=True``) from its published algorithm on the Procrustes step the reference file
vendors (:61-104): that one function is PARITY UNPINNED against trimesh itself.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree
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


# ---------------------------------------------------------------------------------------------------------------
# chamfer.py:13-180 -- the alignment helpers
def apply_matrix(points, matrix):
    """chamfer.py:13-58 (translate=True): rows of ``points`` through a homogeneous matrix; near-identity is a no-op."""
    points = np.asarray(points, np.float64)
    matrix = np.asarray(matrix, np.float64)
    if len(points) == 0 or np.abs(matrix - np.eye(len(matrix))).max() < 1e-8:
        return points.copy()
    d = points.shape[1]
    homog = np.concatenate([points, np.ones((len(points), 1))], axis=1)
    return np.ascontiguousarray((matrix @ homog.T).T[:, :d])


def procrustes(a, b, reflection=True, translation=True, scale=True):
    """chamfer.py:61-104 -> (matrix, transformed a, mean squared residual)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = a.shape[1]
    ca, cb = (a.mean(0), b.mean(0)) if translation else (np.zeros(d), np.zeros(d))
    ra = np.sqrt(((a - ca) ** 2).sum() / len(a)) if scale else 1
    rb = np.sqrt(((b - cb) ** 2).sum() / len(b)) if scale else 1
    u, _, vh = np.linalg.svd(((b - cb) / rb).T @ ((a - ca) / ra))
    rot = u @ vh if reflection else u @ np.diag([1, 1, np.linalg.det(u @ vh)]) @ vh
    k = rb / ra
    matrix = np.eye(d + 1)
    matrix[:d, :d] = k * rot
    matrix[:d, d] = cb - k * (rot @ ca)
    moved = apply_matrix(a, matrix)
    return matrix, moved, ((b - moved) ** 2).mean()


def procrustes_without_rot(a, b):
    """chamfer.py:107-130: one scale + one translation by the pseudo-inverse of the [3 n, 4] system"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    A = np.concatenate([a.reshape(-1, 1), np.tile(np.eye(3), (len(a), 1))], axis=1)
    x = np.linalg.inv(A.T @ A) @ A.T @ b.reshape(-1)
    matrix = np.diag([x[0], x[0], x[0], 1.0])
    matrix[:3, 3] = x[1:]
    moved = apply_matrix(a, matrix)
    return matrix, moved, ((b - moved) ** 2).mean()


def icp_two_sided(a, b, initial=np.identity(4), threshold=1e-5, max_iterations=20, rot=False):
    """chamfer.py:133-180.  The two trees index the clouds AS PASSED IN (:136-137); the matches are read from the
    moving clouds."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    tree_a, tree_b = cKDTree(a), cKDTree(b)
    a, b = apply_matrix(a, initial), apply_matrix(b, initial)
    step = procrustes if rot else procrustes_without_rot
    last = np.inf
    n_iter = 0
    for _ in range(max_iterations):
        n_iter += 1
        _, ia = tree_b.query(a, 1)
        _, new_a, cost_a = step(a, b[ia])
        _, ib = tree_a.query(b, 1)
        _, new_b, cost_b = step(b, a[ib])
        cost = cost_a + cost_b
        a, b = new_a, new_b
        if last - cost < threshold:
            break
        last = cost
    return a, b, cost, n_iter


def registration_icp(a, b, initial=None, threshold=1e-5, max_iterations=20):
    """trimesh.registration.icp (published algorithm; see the header) -> (matrix, transformed a, cost, iterations)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    tree = cKDTree(b)
    total = np.eye(4) if initial is None else np.asarray(initial, np.float64)
    a = apply_matrix(a, total)
    last = np.inf
    n_iter = 0
    for _ in range(max_iterations):
        n_iter += 1
        _, ix = tree.query(a, 1)
        matrix, a, cost = procrustes(a, b[ix])
        total = matrix @ total
        if last - cost < threshold:
            break
        last = cost
    return total, a, cost, n_iter
