"""Drop-in for the reference's ``deep_sdf`` package as far as the hot path and its evaluation reach
(deep_sdf/__init__.py:4-9 star-imports its sub-modules, so callers write ``deep_sdf.create_mesh``,
``deep_sdf.decode_sdf``, ``deep_sdf.ICP_T_S``, ``deep_sdf.metrics.chamfer.compute_trimesh_chamfer`` -- evaluate.py:64).
``deep_sdf.data`` / ``deep_sdf.workspace`` (datasets, experiment directories) are outside the path."""
from . import mesh, metrics, utils  # noqa: F401
from .mesh import convert_sdf_samples_to_ply, create_mesh  # noqa: F401
from .metrics.chamfer import (compute_trimesh_chamfer, icp, procrustes, procrustes_without_rot,  # noqa: F401
                              transform_points)
from .metrics.icp_trans_scale import ICP_T_S  # noqa: F401
from .utils import decode_sdf  # noqa: F401
