"""Drop-in for the hot-path part of the reference's ``utils/utils.py``.

    kinematic_embedding      <-> utils/utils.py:376-430
    get_nerf_embedder        <-> utils/utils.py:521-533
    decode_sdf_multi_output  <-> utils/utils.py:561-572
    decode_sdf_points        (new) the fast arbitrary-point path: raw xyz in, pose-align folded

Both keep the reference's argument order.  Inputs must be CUDA tensors.
"""
from __future__ import annotations

import torch

from . import engine as _engine


def kinematic_embedding(xyz, mano_results, num_points_per_scene, point_feat_size, scale_factor,
                        obj_results, encode_style):
    """[P,3] -> [P,point_feat_size] pose-aligned features (batch of one sample, as on the
    reconstruction path: utils/mesh.py:51-52 calls it with num_points_per_scene = chunk size)."""
    xyz = xyz.reshape(-1, 3)
    if xyz.shape[0] != num_points_per_scene:
        raise ValueError("alignsdf_b200.kinematic_embedding handles one sample per call "
                         "(batched training-time use is outside the reconstruction path)")
    specs = dict(PointFeatSize=point_feat_size, EncodeStyle=encode_style, SdfScaleFactor=scale_factor)
    if point_feat_size <= 3:
        raise ValueError("kinematic_embedding needs PointFeatSize > 3")
    return _engine.embed_points(xyz, specs, mano_results, obj_results)


def get_nerf_embedder(multires):
    """-> (embed, out_dim) like the reference: embed(x[..., 3]) = [x, sin(2^f x), cos(2^f x) ...]."""
    multires = int(multires)
    return (lambda x, m=multires: _engine.nerf_embed(x, m)), 3 + 6 * multires


def decode_sdf_multi_output(decoder, latent_vector, queries, mano_results, cam_intr, specs):
    """queries are already-embedded features [P, PointFeatSize] (what the reference passes);
    returns (sdf_hand [P,1], sdf_obj [P,1], predicted_class) where predicted_class is what the reference's
    decoder returns as its third output: the raw classifier logits [P, num_class] (networks/model.py:161-162,188)
    or ``Tensor([0])`` when the decoder has no classifier head (:188,350)."""
    eng = _engine.get_engine(decoder, queries.device)
    # specs['PixelAlign']: latent_vector is the image feature map, the per-point latent is sampled in the kernel
    # from queries[:, :3] like utils/utils.py:563-566 (alignsdf_b200/pixel_align.py)
    bound = eng.bind(latent_vector, specs, mano_results, None, feature_mode=True, cam_intr=cam_intr)
    want_cls = eng.topo.classifier is not None
    hand, obj, cls = bound.eval_points(queries, want_cls=want_cls, want_logits=want_cls)
    if obj is None:
        obj = torch.zeros_like(hand)
    predicted = cls[1] if want_cls else torch.zeros(1, device=queries.device)
    return hand.unsqueeze(1), obj.unsqueeze(1), predicted


def decode_sdf_points(decoder, latent_vector, xyz, mano_results, obj_results, specs, want_cls=False):
    """Raw xyz [P,3] -> (sdf_hand [P], sdf_obj [P], class argmax or None); pose-align folded."""
    eng = _engine.get_engine(decoder, xyz.device)
    bound = eng.bind(latent_vector, specs, mano_results, obj_results)
    return bound.eval_points(xyz, want_cls=want_cls)
