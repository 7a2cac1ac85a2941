"""Host-side weight folding and packing (pure host logic, fp64 -> fp32/fp16).

The reference evaluates, for every query point, ``decoder(cat([latent,
kinematic_embedding(xyz)]))`` (utils/utils.py:376-430,561-572,
networks/model.py:285-350).  Two facts make most of that work per-sample
constants:

* the latent is the same for every point of a sample, so the latent columns of
  layer 0 and of every ``latent_in`` layer fold into that layer's bias;
* for rigid ``global_trans`` / ``obj_trans`` the homogeneous divide is by
  exactly 1, so every pose-aligned feature block is affine in xyz and the
  feature columns fold into a [out,3] matrix applied to xyz directly.

``fold_decoder`` produces, per branch and layer, ``y = Wx.x_prev + M.u + B``
with ``u`` = xyz (grid / point mode, D=3) or = the already-embedded features
(feature mode used by ``decode_sdf_multi_output``, D=pf_branch).  All folding
is done in float64 and rounded once to float32.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

ASDF_MAX_LAYERS = 8
ASDF_MAX_POINT_DIM = 64


# ----------------------------------------------------------------------------
# topology extraction
# ----------------------------------------------------------------------------
def _clean_state_dict(decoder) -> dict:
    sd = decoder.state_dict() if hasattr(decoder, "state_dict") else dict(decoder)
    out = {}
    for k, v in sd.items():
        for pre in ("module.decoder.", "module.", "decoder."):
            if k.startswith(pre):
                k = k[len(pre):]
        out[k] = v.detach().cpu().double().numpy()
    return out


def _layer_weight(sd, name):
    """Effective [out,in] weight: g*v/||v||_row (networks/model.py:249-250) or plain."""
    if f"{name}.weight_g" in sd:
        v = sd[f"{name}.weight_v"]
        g = sd[f"{name}.weight_g"].reshape(-1, 1)
        return g * v / np.sqrt((v * v).sum(1, keepdims=True))
    if f"{name}.parametrizations.weight.original0" in sd:  # new-style weight_norm
        v = sd[f"{name}.parametrizations.weight.original1"]
        g = sd[f"{name}.parametrizations.weight.original0"].reshape(-1, 1)
        return g * v / np.sqrt((v * v).sum(1, keepdims=True))
    return sd[f"{name}.weight"]


@dataclass
class Topology:
    kind: str                 # "separate" | "combined"
    latent_size: int
    point_feat_size: int
    encode_style: str
    latent_in: tuple
    xyz_in_all: bool
    pre_tanh: bool
    branches: list            # [(tag, prefix)]
    n_layers: int             # linear layers per branch
    layers: dict = field(default_factory=dict)      # prefix -> [(W, b)]
    layer_norms: dict = field(default_factory=dict) # prefix -> [None | (gamma, beta)]  (LayerNorm before the ReLU)
    classifier: tuple | None = None                 # (Wc, bc)
    f32_cache: dict = field(default_factory=dict, repr=False)   # (prefix, layer, shape) -> f32 copy of the static weights


def decoder_topology(decoder) -> Topology:
    sd = _clean_state_dict(decoder)
    separate = any(k.startswith("linh") for k in sd)
    prefixes = [("hand", "linh"), ("obj", "lino")] if separate else [("both", "lin")]
    layers, layer_norms = {}, {}
    for _, prefix in prefixes:
        ls, lns, i = [], [], 0
        bn = prefix.replace("lin", "bn")                  # bnh / bno / bn: nn.LayerNorm when weight_norm is off
        while any(k.startswith(f"{prefix}{i}.") for k in sd):
            ls.append((_layer_weight(sd, f"{prefix}{i}"), sd[f"{prefix}{i}.bias"]))
            lns.append((sd[f"{bn}{i}.weight"], sd[f"{bn}{i}.bias"]) if f"{bn}{i}.weight" in sd else None)
            i += 1
        if lns and lns[-1] is not None:
            raise ValueError("LayerNorm on the output layer is not something the reference can build")
        layers[prefix], layer_norms[prefix] = ls, lns
    n_layers = len(next(iter(layers.values())))
    if n_layers < 2 or n_layers > ASDF_MAX_LAYERS:
        raise ValueError(f"unsupported number of linear layers: {n_layers}")
    pf = int(getattr(decoder, "point_feat_size", 3))          # plain DeepSDF decoders: xyz only
    style = str(getattr(decoder, "encode_style", "nerf"))
    if separate:
        d0_hand = layers["linh"][0][0].shape[1]
        sub = {"nerf": pf, "hand": pf, "obj": 3, "both": pf - 3}[style]
        latent_size = d0_hand - sub
    else:
        latent_size = layers["lin"][0][0].shape[1] - pf
    cls = None
    if "classifier_head.weight" in sd:
        cls = (sd["classifier_head.weight"], sd["classifier_head.bias"])
    return Topology(
        kind="separate" if separate else "combined", latent_size=int(latent_size),
        point_feat_size=pf, encode_style=style,
        latent_in=tuple(int(x) for x in getattr(decoder, "latent_in", ())),
        xyz_in_all=bool(getattr(decoder, "xyz_in_all", False)),
        pre_tanh=bool(getattr(decoder, "use_tanh", False)),
        branches=prefixes, n_layers=n_layers, layers=layers, layer_norms=layer_norms, classifier=cls)


# ----------------------------------------------------------------------------
# pose-align embedding as an affine map (utils/utils.py:376-430, Appendix A)
# ----------------------------------------------------------------------------
def uses_kinematic_embedding(specs, mano_results) -> bool:
    """Condition at utils/mesh.py:49-50."""
    return (specs["PointFeatSize"] > 3 and mano_results is not None
            and specs["EncodeStyle"] != "nerf")


def nerf_freqs(specs, mano_results) -> int:
    """Number of NeRF frequencies when the reference takes the positional-encoding branch
    (utils/mesh.py:49-55: PointFeatSize > 3 and (no MANO results or EncodeStyle 'nerf')), else 0."""
    pf = int(specs["PointFeatSize"])
    if pf <= 3 or uses_kinematic_embedding(specs, mano_results):
        return 0
    if (pf - 3) % 6 != 0:
        raise ValueError(f"PointFeatSize={pf} is not 3 + 6 * multires (NeRF positional encoding)")
    return (pf - 3) // 6


def _rigid_inverse(T: np.ndarray, what: str):
    inv = np.linalg.inv(T)
    if not np.allclose(inv[3], [0, 0, 0, 1], atol=1e-6):
        raise ValueError(f"{what} is not affine (last row of its inverse is {inv[3]}); "
                         "the homogeneous divide cannot be folded")
    return inv[:3, :3], inv[:3, 3]


def embedding_affine(specs, mano_results, obj_results):
    """Return (A [pf,3], c [pf]) with features = A.xyz + c, in float64."""
    pf = int(specs["PointFeatSize"])
    if not uses_kinematic_embedding(specs, mano_results):
        if pf != 3:
            raise ValueError("NeRF positional encoding is not an affine map of xyz: fold with "
                             "feature_mode=True and let the kernel encode (packer.nerf_freqs)")
        return np.eye(3), np.zeros(3)
    style = specs["EncodeStyle"]
    s = float(specs["SdfScaleFactor"])
    blocks_A, blocks_c = [], []
    if style in ("hand", "both"):
        G = mano_results["global_trans"].detach().cpu().double().numpy()
        r = mano_results["rot_center"].detach().cpu().double().numpy().reshape(-1, 3)
        if G.shape[0] != 1:
            raise ValueError("reconstruction path is batch-1 (utils/mesh.py:51-52)")
        G, r = G[0], r[0]
        blocks_A.append(np.eye(3))
        blocks_c.append(r * s / 2)
        single = (pf == 6 and style == "hand") or (pf == 9 and style == "both")
        for j in range(1 if single else G.shape[0]):
            Ri, ti = _rigid_inverse(G[j], f"global_trans[{j}]")
            blocks_A.append(Ri)
            blocks_c.append((Ri @ r + ti) * s / 2)
    if style in ("obj", "both"):
        T = obj_results["obj_trans"].detach().cpu().double().numpy()
        Ri, ti = _rigid_inverse(T[0], "obj_trans")
        if style == "obj":
            blocks_A.append(np.eye(3))
            blocks_c.append(np.zeros(3))
        blocks_A.append(Ri)
        blocks_c.append(ti * s / 2)
    A, c = np.concatenate(blocks_A, 0), np.concatenate(blocks_c, 0)
    if A.shape[0] != pf:
        raise ValueError(f"embedding produces {A.shape[0]} features but PointFeatSize={pf}")
    return A, c


def embed_points_torch(xyz: torch.Tensor, sample, dtype=torch.float32) -> torch.Tensor:
    """Host helper (fit / tests): features = A.xyz + c evaluated with torch."""
    A, c = embedding_affine(sample.specs, sample.mano_results, sample.obj_results)
    A = torch.as_tensor(A, dtype=dtype, device=xyz.device)
    c = torch.as_tensor(c, dtype=dtype, device=xyz.device)
    return xyz.to(dtype) @ A.T + c


def branch_feature_index(topo: Topology, tag: str) -> np.ndarray:
    """Which of the pf embedded features each branch sees (networks/model.py:288-299)."""
    pf, s = topo.point_feat_size, topo.encode_style
    allf = np.arange(pf)
    if topo.kind == "combined" or s == "nerf":
        return allf
    if s == "hand":
        return allf if tag == "hand" else allf[:3]
    if s == "obj":
        return allf[:3] if tag == "hand" else allf
    return allf[:pf - 3] if tag == "hand" else np.concatenate([allf[:3], allf[pf - 3:]])


# ----------------------------------------------------------------------------
# folding
# ----------------------------------------------------------------------------
@dataclass
class FoldedLayer:
    Wx: np.ndarray | None     # [out, h] f32, weights on the previous activations
    M: np.ndarray | None      # [out, D] f32, weights on the per-point vector u
    B: np.ndarray             # [out]    f32
    ln: tuple | None = None   # (gamma [out], beta [out]) f32: LayerNorm(eps 1e-5) between the linear map and the ReLU


@dataclass
class FoldedBranch:
    tag: str
    layers: list
    point_dim: int            # D


def fold_decoder(topo: Topology, latent, specs, mano_results, obj_results,
                 feature_mode: bool = False, affine=None):
    """Fold latent (+ embedding affine unless ``feature_mode``) into the layers.  ``affine``: the result of
    ``embedding_affine`` for this sample when the caller already has it."""
    L = topo.latent_size
    if specs.get("PixelAlign", False):
        z = np.zeros(L)             # per-point latents: applied by the kernel from projected feature maps (pixel_align.py)
    else:
        z = latent.detach().cpu().double().numpy().reshape(-1)
    if z.shape[0] != L:
        raise ValueError(f"latent has {z.shape[0]} entries, decoder expects {L}")
    if feature_mode:
        A_full = np.eye(topo.point_feat_size)
        c_full = np.zeros(topo.point_feat_size)
    else:
        A_full, c_full = affine if affine is not None else embedding_affine(specs, mano_results, obj_results)
    out = []
    for tag, prefix in topo.branches:
        idx = branch_feature_index(topo, tag)
        nf = len(idx)
        if feature_mode:        # u = this branch's own slice of the embedded features
            A, c = np.eye(nf), np.zeros(nf)
        else:
            A, c = A_full[idx], c_full[idx]
        folded = []
        for l, (W, b) in enumerate(topo.layers[prefix]):
            d0 = L + nf
            if l == 0:
                if W.shape[1] != d0:
                    raise ValueError(f"{prefix}0 expects {W.shape[1]} inputs, got {d0}")
                Wz, Wf = W[:, :L], W[:, L:]
                folded.append(FoldedLayer(None, Wf @ A, b + Wz @ z + Wf @ c))
            elif l in topo.latent_in:
                h = W.shape[1] - d0
                Wz, Wf = W[:, h:h + L], W[:, h + L:]
                folded.append(FoldedLayer(W[:, :h], Wf @ A, b + Wz @ z + Wf @ c))
            elif topo.xyz_in_all and topo.kind == "combined":
                h = W.shape[1] - topo.point_feat_size
                Wf = W[:, h:]
                folded.append(FoldedLayer(W[:, :h], Wf @ A_full, b + Wf @ c_full))
            else:
                folded.append(FoldedLayer(W, None, b.copy()))
        D = A.shape[1]
        for l, fl in enumerate(folded):
            ln = topo.layer_norms.get(prefix, [None] * len(folded))[l]
            fl.ln = None if ln is None else (np.ascontiguousarray(ln[0], dtype=np.float32),
                                             np.ascontiguousarray(ln[1], dtype=np.float32))
        for l, fl in enumerate(folded):
            if fl.Wx is not None:                  # static per decoder: convert once, not once per sample
                key = (prefix, l, fl.Wx.shape)
                cached = topo.f32_cache.get(key)
                if cached is None:
                    cached = topo.f32_cache[key] = np.ascontiguousarray(fl.Wx, dtype=np.float32)
                fl.Wx = cached
            fl.M = None if fl.M is None else np.ascontiguousarray(fl.M, dtype=np.float32)
            fl.B = np.ascontiguousarray(fl.B, dtype=np.float32)
        out.append(FoldedBranch(tag, folded, D))
    return out


def folded_forward_numpy(branches, u: np.ndarray, pre_tanh=False, dtype=np.float32):
    """Evaluate the folded network on host (tests only; validates folding)."""
    res = []
    for br in branches:
        x = None
        n = len(br.layers)
        for l, fl in enumerate(br.layers):
            y = np.broadcast_to(fl.B.astype(dtype), (u.shape[0], fl.B.shape[0])).copy()
            if fl.Wx is not None:
                y += x @ fl.Wx.astype(dtype).T
            if fl.M is not None:
                y += u.astype(dtype) @ fl.M.astype(dtype).T
            if l == n - 1:
                if pre_tanh:
                    y = np.tanh(y)
                y = np.tanh(y)
            else:
                if fl.ln is not None:
                    mu = y.mean(1, keepdims=True)
                    var = ((y - mu) ** 2).mean(1, keepdims=True)
                    y = (y - mu) / np.sqrt(var + dtype(1e-5)) * fl.ln[0].astype(dtype) + fl.ln[1].astype(dtype)
                y = np.maximum(y, 0)
            x = y
        res.append(x)
    return res


# ----------------------------------------------------------------------------
# generic fp32 pack (csrc/k1_simt.cu)
# ----------------------------------------------------------------------------
def _pad8(n):
    return (n + 7) // 8 * 8


@dataclass
class SimtPack:
    """Flat float buffers + an int32 table in the layout k1_simt.cu reads.

    static  : per branch, per layer  WxT [h][npad]   (transposed, zero padded), then gamma [n] | beta [n]
              when the layer is followed by a LayerNorm
    sample  : per branch, per layer  MB  [npad][D+1] (M row then B)
    table   : per branch, per layer  (h, n, npad, has_M, off_static, off_sample, off_layernorm or -1, 0)
    """
    static: np.ndarray
    sample: np.ndarray
    table: np.ndarray
    n_branches: int
    n_layers: int
    point_dim: np.ndarray     # per branch D
    n_outputs: int
    max_width: int


def pack_simt(branches, want_static=True) -> SimtPack:
    """``want_static=False``: only the per-sample part and the table (the weight block does not depend on the
    sample and is uploaded once per decoder)."""
    n_layers = len(branches[0].layers)
    static, sample, table = [], [], []
    off_s = off_p = 0
    max_w = 0
    for br in branches:
        D = br.point_dim
        if D > ASDF_MAX_POINT_DIM:
            raise ValueError(f"point dim {D} > {ASDF_MAX_POINT_DIM}")
        for fl in br.layers:
            n = fl.B.shape[0]
            npad = _pad8(n)
            h = 0 if fl.Wx is None else fl.Wx.shape[1]
            max_w = max(max_w, npad, h)
            if h and want_static:
                wt = np.zeros((h, npad), np.float32)
                wt[:, :n] = fl.Wx.T
                static.append(wt.reshape(-1))
            mb = np.zeros((npad, D + 1), np.float32)
            if fl.M is not None:
                mb[:n, :D] = fl.M
            mb[:n, D] = fl.B
            sample.append(mb.reshape(-1))
            off_ln = -1
            off_w = off_s
            off_s += h * npad
            if fl.ln is not None:
                if want_static:
                    static.append(np.concatenate([fl.ln[0], fl.ln[1]]).astype(np.float32))
                off_ln = off_s
                off_s += 2 * n
                if off_s % 4:                           # keep the next weight block 16-byte aligned
                    if want_static:
                        static.append(np.zeros(4 - off_s % 4, np.float32))
                    off_s += 4 - off_s % 4
            table.append((h, n, npad, int(fl.M is not None), off_w, off_p, off_ln, 0))
            off_p += npad * (D + 1)
    return SimtPack(
        static=np.concatenate(static) if static else np.zeros(1, np.float32),
        sample=np.concatenate(sample),
        table=np.asarray(table, np.int32).reshape(len(branches), n_layers, 8),
        n_branches=len(branches), n_layers=n_layers,
        point_dim=np.asarray([b.point_dim for b in branches], np.int32),
        n_outputs=int(branches[0].layers[-1].B.shape[0]), max_width=int(max_w))
