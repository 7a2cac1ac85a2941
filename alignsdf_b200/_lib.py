"""ctypes binding of libalignsdf_b200.so (include/alignsdf_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised.  Nothing in this package computes SDF values or
meshes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

ASDF_MAX_LAYERS = 8
ASDF_MAX_POINT_DIM = 64
QUERY_GRID_REFERENCE, QUERY_GRID_REGULAR, QUERY_POINTS = 0, 1, 2
TC_F16X3, TC_F16_F8, TC_F16X1 = 0, 1, 2          # asdf_tc_launch.kind
ABI_VERSION = 4

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libalignsdf_b200.so")


class AsdfError(RuntimeError):
    pass


class Query(C.Structure):
    _fields_ = [("mode", C.c_int32), ("N", C.c_int32), ("begin", C.c_int64), ("end", C.c_int64),
                ("voxel", C.c_float), ("origin", C.c_float * 3), ("points_dev", C.c_void_p),
                ("point_stride", C.c_int32), ("bbox_mask", C.c_int32)]


class PixelAlign(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("fh", C.c_int32), ("fw", C.c_int32), ("layer", C.c_int32 * 2),
                ("point_affine", C.c_float * 12), ("cam", C.c_float * 12), ("image_size", C.c_float),
                ("reserved", C.c_int32), ("slot_stride", C.c_int64), ("maps_dev", C.c_void_p)]


class SimtDesc(C.Structure):
    _fields_ = [("n_branches", C.c_int32), ("n_layers", C.c_int32), ("n_outputs", C.c_int32),
                ("pre_tanh", C.c_int32), ("n_class", C.c_int32), ("nerf_freqs", C.c_int32),
                ("point_dim", C.c_int32 * 2),
                ("point_index", (C.c_int32 * ASDF_MAX_POINT_DIM) * 2),
                ("table", ((C.c_int32 * 8) * ASDF_MAX_LAYERS) * 2), ("pa", PixelAlign)]


class TcLaunch(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_decoders", C.c_int32), ("n_samples", C.c_int32), ("reserved", C.c_int32),
                ("static_dev", C.c_void_p), ("samples_dev", C.c_void_p), ("sample_stride", C.c_int64),
                ("grid_dev", C.c_void_p), ("out_hand_dev", C.c_void_p), ("out_obj_dev", C.c_void_p),
                ("out_stride", C.c_int64), ("bbox_dev", C.c_void_p), ("status_dev", C.c_void_p),
                ("bbox_tau", C.c_float), ("amb_capacity", C.c_int32), ("amb_dev", C.c_void_p),
                ("amb_count_dev", C.c_void_p)]


class TcBindDesc(C.Structure):
    _fields_ = [("n_decoders", C.c_int32), ("latent_size", C.c_int32), ("n_features", C.c_int32 * 2),
                ("feature_index", (C.c_int32 * ASDF_MAX_POINT_DIM) * 2), ("decoder_stride", C.c_int64),
                ("act_scale", C.c_float), ("p_absmax", C.c_float), ("w_scale", (C.c_double * 3) * 2)]


class McParams(C.Structure):
    _fields_ = [("n0", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32), ("full1", C.c_int32),
                ("full2", C.c_int32), ("index0_offset", C.c_int64), ("iso", C.c_float),
                ("spacing", C.c_double * 3), ("origin", C.c_float * 3), ("grid_dev", C.c_void_p)]


_lib = None


def lib():
    """Load (once) and return the shared library; raise if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise AsdfError(f"{_LIB_PATH} is missing: build it with `python -m alignsdf_b200.build` "
                        "(there is no CPU / PyTorch fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i32p = C.c_void_p, C.c_void_p
    L.asdf_abi_version.restype = C.c_int
    L.asdf_last_error.restype = C.c_char_p
    L.asdf_device_ok.restype = C.c_int
    L.asdf_simt_eval.restype = C.c_int
    L.asdf_simt_eval.argtypes = [C.POINTER(SimtDesc), vp, vp, vp, C.POINTER(Query), vp, vp, i32p, vp, i32p, vp]
    L.asdf_tc_eval.restype = C.c_int
    L.asdf_tc_eval.argtypes = [C.POINTER(TcLaunch), C.POINTER(Query), vp]
    L.asdf_tc_static_bytes.restype = C.c_int64
    L.asdf_tc_static_bytes.argtypes = [C.c_int32, C.c_int32]
    L.asdf_tc_sample_bytes.restype = C.c_int64
    L.asdf_tc_bind.restype = C.c_int
    L.asdf_tc_bind.argtypes = [C.POINTER(TcBindDesc), vp, vp, vp, C.c_int32, vp, vp, C.c_int64, vp, vp]
    L.asdf_tc_bind_static_doubles.restype = C.c_int64
    L.asdf_tc_bind_static_doubles.argtypes = [C.c_int32, C.c_int32]
    L.asdf_regrid.restype = C.c_int
    L.asdf_regrid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp, vp, vp]
    L.asdf_grid_points.restype = C.c_int
    L.asdf_grid_points.argtypes = [C.POINTER(Query), vp, vp]
    L.asdf_nerf_embed.restype = C.c_int
    L.asdf_nerf_embed.argtypes = [vp, C.c_int64, C.c_int32, vp, vp]
    L.asdf_embed_points.restype = C.c_int
    L.asdf_embed_points.argtypes = [vp, C.c_int64, vp, C.c_int32, vp, vp]
    L.asdf_mc_scratch_bytes.restype = C.c_size_t
    L.asdf_mc_scratch_bytes.argtypes = [C.POINTER(McParams)]
    L.asdf_mc_count.restype = C.c_int
    L.asdf_mc_count.argtypes = [vp, C.POINTER(McParams), vp, vp, vp]
    L.asdf_mc_emit.restype = C.c_int
    L.asdf_mc_emit.argtypes = [vp, C.POINTER(McParams), vp, C.c_int64, vp, vp, vp, vp, vp]
    L.asdf_cc_label.restype = C.c_int
    L.asdf_cc_label.argtypes = [vp, C.c_int64, C.c_int64, vp, vp]
    L.asdf_cc_stats.restype = C.c_int
    L.asdf_cc_stats.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, C.POINTER(C.c_float * 3), vp, vp, vp, vp, vp]
    L.asdf_cc_mark.restype = C.c_int
    L.asdf_cc_mark.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int32, vp, vp, vp]
    L.asdf_cc_gather.restype = C.c_int
    L.asdf_cc_gather.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, vp]
    L.asdf_nn_search.restype = C.c_int
    L.asdf_nn_search.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, vp, vp]
    if L.asdf_abi_version() != ABI_VERSION:
        raise AsdfError("ABI version mismatch between alignsdf_b200 and its shared library")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        raise AsdfError(f"{what} failed ({rc}): {lib().asdf_last_error().decode()}")


def require_cuda(t: torch.Tensor, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise AsdfError(f"{name} must be a CUDA tensor: alignsdf_b200 has no CPU path")
    return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
