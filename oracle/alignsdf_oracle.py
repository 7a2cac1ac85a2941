"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not a fallback.

CPU restatement (torch-CPU fp32 + numpy) of AlignSDF's dense-grid SDF query
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package ``alignsdf_b200`` never does.

Parity status
-------------
* SDF field (grid -> pose-align embedding -> decoder -> bbox re-grid): PINNED.
  ``oracle/make_golden.py`` runs the reference's own, unmodified
  ``utils.mesh.create_mesh_combined_decoder`` (imported from /root/reference in
  the authoring container) on the same synthetic inputs and asserts this
  restatement reproduces its grid coordinates bit-exactly and its fields to
  <= 1e-6 (<= 2e-6 for the LayerNorm decoders) over 15 configurations (pose-align
  styles, NeRF positional encoding, LayerNorm, CombinedDecoder variants); the
  captured reference outputs are committed under ``tests/golden``.
* Marching cubes / component filter: PARITY UNPINNED.  The arithmetic lives in
  scikit-image (``marching_cubes_lewiner``, requirements.txt:5, unpinned, not
  installed, no network) and trimesh; see ``oracle/mc_oracle.py``.

Each function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import torch


# ----------------------------------------------------------------------------
# grid generation  (utils/mesh.py:24-40, 82-94 ; deep_sdf/mesh.py:21-37)
# ----------------------------------------------------------------------------
def grid_points(N: int, voxel_size, origin, mode: str = "reference",
                start: int = 0, stop: int | None = None) -> torch.Tensor:
    """[stop-start, 3] f32 query coordinates of linear indices start..stop.

    mode="reference": the sheared grid the reference really evaluates: integer
    index tensor ``/ N`` is *true* division under torch >= 1.7, so axis 1 gets
    ``fmod(float(i)/N, N)`` and axis 0 ``fmod((float(i)/N)/N, N)``.
    mode="regular": floor division (the intended lattice).
    ``voxel_size`` may be a python float (pass 1: rounded to f32 by the multiply)
    or a 0-dim f32 tensor; ``origin`` 3 floats / f32 tensor, column k uses
    origin[k] (pass 2 pairing; pass 1 pairs k with origin[2-k] but all are -1).
    """
    stop = N ** 3 if stop is None else stop
    idx = torch.arange(start, stop, 1, dtype=torch.int64)
    out = torch.zeros(stop - start, 3)
    if mode == "reference":
        q = idx / N                      # int64 -> f32, true division
        out[:, 2] = idx % N
        out[:, 1] = q % N
        out[:, 0] = (q / N) % N
    elif mode == "regular":
        out[:, 2] = idx % N
        out[:, 1] = torch.div(idx, N, rounding_mode="floor") % N
        out[:, 0] = torch.div(idx, N * N, rounding_mode="floor") % N
    else:
        raise ValueError(mode)
    org = [origin[k] for k in range(3)]
    for k in range(3):
        out[:, k] = (out[:, k] * voxel_size) + org[k]
    return out


# ----------------------------------------------------------------------------
# pose-align embedding  (utils/utils.py:376-430)
# ----------------------------------------------------------------------------
def kinematic_embedding(xyz, mano_results, point_feat_size, scale_factor, obj_results, encode_style):
    """Explicit (unfolded) fp32 embedding for one sample, xyz [P,3] -> [P,pf]."""
    P = xyz.shape[0]
    wrist = xyz * 2 / scale_factor                                   # :384
    ones = torch.ones(P, 1, dtype=xyz.dtype)
    blocks = []
    if encode_style in ("hand", "both"):
        mano = wrist + mano_results["rot_center"].reshape(1, 3)      # :387
        Ginv = torch.linalg.inv(mano_results["global_trans"][0])    # :393  [16,4,4]
        homo = torch.cat([mano, ones], 1)                            # :390
        inv = torch.einsum("jab,pb->pja", Ginv, homo)                # :394
        inv = inv[..., :3] / inv[..., 3:4]                           # :396
        single = (point_feat_size == 6 and encode_style == "hand") or \
                 (point_feat_size == 9 and encode_style == "both")   # :399
        if single:
            inv = inv[:, :1]
        hand = torch.cat([mano[:, None], inv], 1).reshape(P, -1)     # :403
        blocks.append(hand * scale_factor / 2)                       # :408
    if encode_style in ("obj", "both"):
        Tinv = torch.linalg.inv(obj_results["obj_trans"])[0]        # :414
        o = torch.cat([wrist, ones], 1) @ Tinv.T                     # :415
        o = o[:, :3] / o[:, 3:4]                                     # :416
        o = o * scale_factor / 2                                     # :417
        if encode_style == "obj":
            blocks.append(xyz)                                       # :418
        blocks.append(o)
    out = torch.cat(blocks, 1)
    assert out.shape[1] == point_feat_size, (out.shape, point_feat_size)
    return out


def nerf_embedding(xyz, multires):
    """utils/utils.py:433-463,521-533 (Embedder / get_nerf_embedder): [x, sin(x f), cos(x f) for
    f = 2^0 .. 2^(multires-1)] along the last axis, fp32; the frequencies come from
    2.**torch.linspace(0, multires-1, multires) like the reference's."""
    freq_bands = 2. ** torch.linspace(0., multires - 1, steps=multires)
    out = [xyz]
    for freq in freq_bands:
        out.append(torch.sin(xyz * freq))
        out.append(torch.cos(xyz * freq))
    return torch.cat(out, -1)


def embed(xyz, specs, mano_results, obj_results):
    """Feature selection logic of utils/mesh.py:49-55."""
    if specs["PointFeatSize"] > 3:
        if mano_results is not None and specs["EncodeStyle"] != "nerf":
            return kinematic_embedding(xyz, mano_results, specs["PointFeatSize"],
                                       specs["SdfScaleFactor"], obj_results, specs["EncodeStyle"])
        return nerf_embedding(xyz, (specs["PointFeatSize"] - 3) // 6)      # utils/mesh.py:54-55
    return xyz


# ----------------------------------------------------------------------------
# decoders as pure functions of a state dict (networks/model.py:79-350)
# ----------------------------------------------------------------------------
def _eff_weight(sd, name):
    if f"{name}.weight_g" in sd:                                     # :249-250 weight_norm
        v, g = sd[f"{name}.weight_v"], sd[f"{name}.weight_g"]
        return g * v / v.norm(dim=1, keepdim=True)
    return sd[f"{name}.weight"]


def _mlp(sd, prefix, x, latent_in, pre_tanh, xyz_all=None):
    inp = x
    n = 0
    while f"{prefix}{n}.bias" in sd:
        n += 1
    cls_in = None
    for l in range(n):
        if l == n - 1:
            cls_in = x                                               # :134-137 classifier input
        if l in latent_in:
            x = torch.cat([x, inp], 1)                               # :141-142 / :310-311
        elif l != 0 and xyz_all is not None:
            x = torch.cat([x, xyz_all], 1)                           # :143-144
        x = torch.nn.functional.linear(x, _eff_weight(sd, f"{prefix}{l}"), sd[f"{prefix}{l}.bias"])
        if l == n - 1 and pre_tanh:
            x = torch.tanh(x)
        if l < n - 1:
            bn = prefix.replace("lin", "bn") + str(l)                # :254-255,317-319 LayerNorm (weight_norm off)
            if f"{bn}.weight" in sd:
                x = torch.nn.functional.layer_norm(x, (x.shape[1],), sd[f"{bn}.weight"], sd[f"{bn}.bias"], 1e-5)
            x = torch.relu(x)
    return torch.tanh(x), cls_in                                     # :324-325 final tanh


def decoder_forward(sd, cfg, inputs):
    """(sdf_hand [P,1], sdf_obj [P,1], class logits or None).

    ``cfg``: dict(kind, latent_size, point_feat_size, encode_style, latent_in,
    xyz_in_all, use_tanh).
    """
    L, pf, style = cfg["latent_size"], cfg["point_feat_size"], cfg["encode_style"]
    lat_in = tuple(cfg.get("latent_in", ()))
    if cfg["kind"] == "separate":
        if style == "nerf":
            xh, xo = inputs, inputs                                  # :288-299
        elif style == "hand":
            xh, xo = inputs, inputs[:, :L + 3]
        elif style == "obj":
            xh, xo = inputs[:, :L + 3], inputs
        else:
            xh, xo = inputs[:, :-3], torch.cat([inputs[:, :L + 3], inputs[:, -3:]], 1)
        h, _ = _mlp(sd, "linh", xh, lat_in, cfg.get("use_tanh", False))
        o, _ = _mlp(sd, "lino", xo, lat_in, cfg.get("use_tanh", False))
        return h[:, 0:1], o[:, 0:1], None
    xyz_all = inputs[:, -pf:] if cfg.get("xyz_in_all") else None
    y, cls_in = _mlp(sd, "lin", inputs, lat_in, cfg.get("use_tanh", False), xyz_all)
    logits = None
    if "classifier_head.weight" in sd:
        logits = torch.nn.functional.linear(cls_in, sd["classifier_head.weight"],
                                            sd["classifier_head.bias"])
    return y[:, 0:1], y[:, 1:2], logits


def pixel_alignment(img_feat, xyz, cam_intr, mano_results, image_size, scale_factor):
    """utils/utils.py:536-558: per-point latent = bicubic sample (align_corners=True) of the image feature map at the
    point's projection; points projecting outside the image take the mean feature."""
    pred_root = mano_results["joints"][:, [0]]
    xyz = xyz.reshape((img_feat.shape[0], -1, 3))
    xyz_cam = (xyz * 2 / scale_factor) + pred_root                       # :539
    batch_size, n = img_feat.shape[0], xyz.shape[1]
    homo = torch.cat([xyz_cam, torch.ones([batch_size, n, 1])], 2)
    xy_img = torch.bmm(cam_intr, homo.transpose(1, 2)).transpose(1, 2)   # :545
    xy_img = (xy_img[:, :, :2] / xy_img[:, :, [2]]).unsqueeze(2)
    uv = xy_img / image_size * 2 - 1                                     # :549
    feat = torch.nn.functional.grid_sample(img_feat, uv, align_corners=True, mode="bicubic")[:, :, :, 0].transpose(1, 2)
    uv = uv.squeeze().reshape((-1, 2))
    inside = (uv[:, 0] >= -1.0) & (uv[:, 0] <= 1.0) & (uv[:, 1] >= -1.0) & (uv[:, 1] <= 1.0)
    out_mask = (~inside).reshape((batch_size, n, -1))
    feat[torch.where(out_mask)[:2]] = img_feat.mean(3).mean(2)[torch.where(out_mask)[:1]]   # :555
    return feat.reshape((batch_size * n, -1))


def decode_points(sd, cfg, latent, xyz, specs, mano_results, obj_results, cam_intr=None):
    """utils/utils.py:561-572 on raw xyz; the PixelAlign branch (:563-566) samples the latent per point from the
    FIRST THREE embedded features (queries[:, :3]), like the reference."""
    feats = embed(xyz, specs, mano_results, obj_results)
    if specs.get("PixelAlign", False):
        lat = pixel_alignment(latent, feats[:, :3], cam_intr, mano_results, specs["ImageSize"][0], specs["SdfScaleFactor"])
        return decoder_forward(sd, cfg, torch.cat([lat, feats], 1))
    inputs = torch.cat([latent.expand(xyz.shape[0], -1), feats], 1)
    return decoder_forward(sd, cfg, inputs)


# ----------------------------------------------------------------------------
# bbox re-grid  (utils/mesh.py:198-256)
# ----------------------------------------------------------------------------
def higher_res_cube(vol_hand, vol_obj, N, voxel_size):
    """-> (new_voxel_size f32 0-dim tensor, new_origin f32[3], min_idx, max_idx)."""
    mins, maxs = [], []
    for vol in (vol_hand, vol_obj):
        if vol is None:
            continue
        idx = torch.nonzero(vol < 0).float()
        if idx.shape[0] == 0:
            mins.append(torch.zeros(3)); maxs.append(torch.zeros(3))   # :209-211
        else:
            mins.append(idx.min(0).values); maxs.append(idx.max(0).values)
    mn = mins[0] if len(mins) == 1 else torch.min(mins[0], mins[1])    # :239-247
    mx = maxs[0] if len(maxs) == 1 else torch.max(maxs[0], maxs[1])
    cube = (torch.max(mx - mn) + 4) * voxel_size                       # :250
    return cube / (N - 1), (mn - 2) * voxel_size - 1.0, mn, mx         # :252-254


# ----------------------------------------------------------------------------
# the whole field path  (utils/mesh.py:17-120 ; deep_sdf/mesh.py:14-55)
# ----------------------------------------------------------------------------
def eval_volume(sd, cfg, latent, specs, mano_results, obj_results, N, voxel_size, origin,
                mode="reference", max_batch=2 ** 18, start=0, stop=None, cam_intr=None):
    stop = N ** 3 if stop is None else stop
    hand = torch.empty(stop - start)
    obj = torch.empty(stop - start)
    cls = torch.zeros(stop - start)
    head = start
    while head < stop:
        end = min(head + max_batch, stop)
        xyz = grid_points(N, voxel_size, origin, mode, head, end)
        h, o, logits = decode_points(sd, cfg, latent, xyz, specs, mano_results, obj_results, cam_intr)
        hand[head - start:end - start] = h[:, 0]
        obj[head - start:end - start] = o[:, 0]
        if logits is not None:
            cls[head - start:end - start] = logits.argmax(1).float()
        head = end
    return hand, obj, cls


def two_pass_field(sd, cfg, latent, specs, mano_results, obj_results, N,
                   hand_branch=True, obj_branch=True, mode="reference", max_batch=2 ** 18, cam_intr=None):
    """Pass 1 on [-1,1]^3, bbox of sdf<0, pass 2 on the refit cube.

    Returns dict(pass1_hand, pass1_obj, voxel, origin, hand, obj, cls) with the
    volumes shaped [N,N,N]."""
    with torch.no_grad():
        vs1 = 2.0 / (N - 1)
        h1, o1, _ = eval_volume(sd, cfg, latent, specs, mano_results, obj_results, N, vs1,
                                [-1, -1, -1], mode, max_batch, cam_intr=cam_intr)
        h1, o1 = h1.reshape(N, N, N), o1.reshape(N, N, N)
        nv, no, mn, mx = higher_res_cube(h1 if hand_branch else None,
                                         o1 if obj_branch else None, N, vs1)
        h2, o2, c2 = eval_volume(sd, cfg, latent, specs, mano_results, obj_results, N, nv, no,
                                 mode, max_batch, cam_intr=cam_intr)
    return dict(pass1_hand=h1, pass1_obj=o1, voxel=nv, origin=no, min_idx=mn, max_idx=mx,
                hand=h2.reshape(N, N, N), obj=o2.reshape(N, N, N), cls=c2.reshape(N, N, N))


def decoder_cfg(dec) -> dict:
    """Config dict from a decoder module (reference's or alignsdf_b200's)."""
    sd = dec.state_dict()
    kind = "separate" if any(k.startswith("linh") for k in sd) else "combined"
    pf, style = dec.point_feat_size, dec.encode_style
    if kind == "separate":
        sub = {"nerf": pf, "hand": pf, "obj": 3, "both": pf - 3}[style]
        L = sd["linh0.bias"].shape[0] and (
            (sd["linh0.weight_v"] if "linh0.weight_v" in sd else sd["linh0.weight"]).shape[1] - sub)
    else:
        L = (sd["lin0.weight_v"] if "lin0.weight_v" in sd else sd["lin0.weight"]).shape[1] - pf
    return dict(kind=kind, latent_size=int(L), point_feat_size=pf, encode_style=style,
                latent_in=tuple(dec.latent_in), xyz_in_all=bool(getattr(dec, "xyz_in_all", False)),
                use_tanh=bool(getattr(dec, "use_tanh", False)))
