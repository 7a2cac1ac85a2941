"""GPU versions of the two evaluation helpers that follow the hot path:
``icp_trans_scale.ICP_T_S`` <-> deep_sdf/metrics/icp_trans_scale.py, ``chamfer.compute_trimesh_chamfer`` <->
deep_sdf/metrics/chamfer.py:183-231.  Same names and call surface; the KD-tree queries run as an exact fp64
brute-force search on the GPU (csrc/nn.cu)."""
from . import chamfer, icp_trans_scale  # noqa: F401
