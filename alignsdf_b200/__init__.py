"""B200-native hot path of AlignSDF (dense-grid SDF query + marching cubes) behind the reference's Python surface.

The reference's ``utils`` package star-imports its sub-modules (utils/__init__.py:3-4), so callers write either
``utils.mesh.create_mesh_combined_decoder`` (reconstruct.py:93) or ``utils.create_mesh_combined_decoder``.  Both
spellings resolve here -- ``alignsdf_b200.mesh.<name>`` and ``alignsdf_b200.<name>`` -- for the functions of the path;
the second is looked up lazily so that importing the package stays as light as importing one of its modules."""
import importlib

_PATH_FUNCTIONS = {
    # utils/mesh.py
    "create_mesh_combined_decoder": "mesh", "get_higher_res_cube": "mesh", "convert_sdf_samples_to_ply": "mesh",
    "write_verts_label_to_npz": "mesh", "write_verts_label_to_obj": "mesh", "write_color_labeled_ply": "mesh",
    # utils/utils.py
    "kinematic_embedding": "utils", "get_nerf_embedder": "utils", "decode_sdf_multi_output": "utils",
    # batch / multi-GPU entry points that have no counterpart in the reference
    "create_meshes_pipelined": "mesh", "sdf_volumes": "mesh", "decode_sdf_points": "utils",
}


def __getattr__(name):
    mod = _PATH_FUNCTIONS.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    return getattr(importlib.import_module("." + mod, __name__), name)


def __dir__():
    return sorted(list(globals()) + list(_PATH_FUNCTIONS))
