"""Weight containers for the two AlignSDF decoder families.

These classes mirror the *constructor signature, attribute names and
state-dict keys* of the reference decoders so that checkpoints written by the
reference (``module.decoder.linh0.weight_g`` ...) load unchanged and so that
``create_mesh_combined_decoder`` accepts either the reference's own module or
one of these:

* ``SeparateDecoder``  <->  /root/reference networks/model.py:191-350
  (ModelType "1encoder2decoder": two independent MLPs ``linh*`` / ``lino*``)
* ``CombinedDecoder``  <->  networks/model.py:79-188
  (ModelType "1encoder1decoder": one MLP ``lin*`` with 2 outputs, optional
  ``classifier_head`` and ``xyz_in_all``)

Their ``forward`` is a plain eager-PyTorch fp32 statement of the network.  It is
NOT on the product path: the product pulls the weights out of ``state_dict()``
(alignsdf_b200/packer.py) and runs hand-written CUDA.  ``forward`` exists so the
classes are usable as ordinary modules (e.g. the fp32 torch cross-check in
tests) and never dispatches to the CUDA library.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class WNLinear(nn.Module):
    """Linear layer parametrised as W = g * v / ||v||_row.

    Same parameter names / shapes as old-style ``torch.nn.utils.weight_norm``
    applied to ``nn.Linear`` (reference networks/model.py:127,250,268):
    ``weight_g [out,1]``, ``weight_v [out,in]``, ``bias [out]``.
    """

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        lin = nn.Linear(in_features, out_features)
        v = lin.weight.detach().clone()
        self.weight_g = nn.Parameter(v.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(v)
        self.bias = nn.Parameter(lin.bias.detach().clone())

    def effective_weight(self) -> torch.Tensor:
        return self.weight_g * self.weight_v / self.weight_v.norm(dim=1, keepdim=True)

    def forward(self, x):
        return F.linear(x, self.effective_weight(), self.bias)


def _make_linear(in_f, out_f, normed):
    return WNLinear(in_f, out_f) if normed else nn.Linear(in_f, out_f)


def _branch_dims(latent_size, point_feat_size, encode_style, dims):
    """Input widths of the hand / object MLPs (networks/model.py:213-225)."""
    if encode_style == "nerf":
        h = o = latent_size + point_feat_size
    elif encode_style == "hand":
        h, o = latent_size + point_feat_size, latent_size + 3
    elif encode_style == "obj":
        h, o = latent_size + 3, latent_size + point_feat_size
    elif encode_style == "both":
        h, o = latent_size + point_feat_size - 3, latent_size + 6
    else:
        raise ValueError(f"unknown encode_style {encode_style!r}")
    return [h] + list(dims) + [1], [o] + list(dims) + [1]


class SeparateDecoder(nn.Module):
    def __init__(self, latent_size, point_feat_size, encode_style, dims, num_class=6,
                 dropout=None, dropout_prob=0.0, norm_layers=(), latent_in=(),
                 weight_norm=False, xyz_in_all=None, use_tanh=False,
                 latent_dropout=False, use_classifier=False):
        super().__init__()
        if use_classifier:
            # the reference raises AttributeError here (networks/model.py:258)
            raise AttributeError("'SeparateDecoder' object has no attribute 'num_layers'")
        self.latent_size = latent_size
        self.point_feat_size = point_feat_size
        self.encode_style = encode_style
        dims_hand, dims_obj = _branch_dims(latent_size, point_feat_size, encode_style, dims)
        self.num_hand_layers = len(dims_hand)
        self.num_obj_layers = len(dims_hand)
        self.num_class = num_class
        self.norm_layers = tuple(norm_layers)
        self.latent_in = tuple(latent_in)
        self.latent_dropout = latent_dropout
        self.xyz_in_all = xyz_in_all
        self.weight_norm = weight_norm
        self.use_classifier = False
        self.use_tanh = use_tanh
        self.dropout = dropout
        self.dropout_prob = dropout_prob
        for prefix, bn, d in (("linh", "bnh", dims_hand), ("lino", "bno", dims_obj)):
            for layer in range(len(d) - 1):
                out_dim = d[layer + 1] - d[0] if (layer + 1) in self.latent_in else d[layer + 1]
                normed = weight_norm and layer in self.norm_layers
                setattr(self, f"{prefix}{layer}", _make_linear(d[layer], out_dim, normed))
                if (not weight_norm) and layer in self.norm_layers:
                    setattr(self, f"{bn}{layer}", nn.LayerNorm(out_dim))

    def _run(self, x, prefix, bn):
        inp = x
        last = self.num_hand_layers - 2
        for layer in range(self.num_hand_layers - 1):
            if layer in self.latent_in:
                x = torch.cat([x, inp], 1)
            x = getattr(self, f"{prefix}{layer}")(x)
            if layer == last and self.use_tanh:
                x = torch.tanh(x)
            if layer < last:
                if layer in self.norm_layers and not self.weight_norm:
                    x = getattr(self, f"{bn}{layer}")(x)
                x = F.relu(x)
        return torch.tanh(x)

    def split_inputs(self, inp):
        L, s = self.latent_size, self.encode_style
        if s == "nerf":
            return inp, inp
        if s == "hand":
            return inp, inp[:, :L + 3]
        if s == "obj":
            return inp[:, :L + 3], inp
        return inp[:, :-3], torch.cat([inp[:, :L + 3], inp[:, -3:]], 1)

    def forward(self, inp):
        xh, xo = self.split_inputs(inp)
        return (self._run(xh, "linh", "bnh")[:, 0:1], self._run(xo, "lino", "bno")[:, 0:1],
                torch.zeros(1, device=inp.device))


class CombinedDecoder(nn.Module):
    def __init__(self, latent_size, point_feat_size, encode_style, dims, num_class=6,
                 dropout=None, dropout_prob=0.0, norm_layers=(), latent_in=(),
                 weight_norm=False, xyz_in_all=None, use_tanh=False,
                 latent_dropout=False, use_classifier=False):
        super().__init__()
        d = [latent_size + point_feat_size] + list(dims) + [2]
        self.latent_size = latent_size
        self.point_feat_size = point_feat_size
        self.encode_style = encode_style
        self.num_layers = len(d)
        self.num_class = num_class
        self.norm_layers = tuple(norm_layers)
        self.latent_in = tuple(latent_in)
        self.latent_dropout = latent_dropout
        self.xyz_in_all = xyz_in_all
        self.weight_norm = weight_norm
        self.use_classifier = use_classifier
        self.use_tanh = use_tanh
        self.dropout = dropout
        self.dropout_prob = dropout_prob
        for layer in range(self.num_layers - 1):
            if (layer + 1) in self.latent_in:
                out_dim = d[layer + 1] - d[0]
            else:
                out_dim = d[layer + 1]
                if xyz_in_all and layer != self.num_layers - 2:
                    out_dim -= point_feat_size
            normed = weight_norm and layer in self.norm_layers
            setattr(self, f"lin{layer}", _make_linear(d[layer], out_dim, normed))
            if (not weight_norm) and layer in self.norm_layers:
                setattr(self, f"bn{layer}", nn.LayerNorm(out_dim))
            if use_classifier and layer == self.num_layers - 2:
                self.classifier_head = nn.Linear(d[layer], num_class)

    def forward(self, inp):
        xyz = inp[:, -self.point_feat_size:]
        x = inp
        cls = None
        last = self.num_layers - 2
        for layer in range(self.num_layers - 1):
            if self.use_classifier and layer == last:
                cls = self.classifier_head(x)
            if layer in self.latent_in:
                x = torch.cat([x, inp], 1)
            elif layer != 0 and self.xyz_in_all:
                x = torch.cat([x, xyz], 1)
            x = getattr(self, f"lin{layer}")(x)
            if layer == last and self.use_tanh:
                x = torch.tanh(x)
            if layer < last:
                if layer in self.norm_layers and not self.weight_norm:
                    x = getattr(self, f"bn{layer}")(x)
                x = F.relu(x)
        x = torch.tanh(x)
        if cls is None:
            cls = torch.zeros(1, device=inp.device)
        return x[:, 0:1], x[:, 1:2], cls
