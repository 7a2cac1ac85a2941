"""PixelAlign (specs['PixelAlign']): per-point latents sampled from an image feature map.

Reference: utils/utils.py:536-558 (pixel_alignment) and :563-566 (decode_sdf_multi_output): the latent of a query
point is the bicubic ``grid_sample`` (align_corners=True) of the encoder's feature map [1, C, fh, fw] at the point's
projection into the image -- the mean feature for points that project outside it -- concatenated with the point's
embedded features.  That breaks the per-sample latent fold of the other configurations, but sampling is LINEAR in
the feature map: the latent columns W_z of the two layers that see the latent (layer 0 and the skip layer) are
applied to the map once per sample,

    G[pixel] = W_z F[:, pixel]   (pixel < fh fw),      G[fh fw] = W_z mean(F),

and the generic kernel (csrc/k1_simt.cu) adds the 16-tap bicubic combination of G's rows to the pre-activations of
those layers: 16 x 512 gathered multiply-adds per point and layer instead of a 256 x 512 product, and no [P, 256]
latent array is ever materialised (4 GB at 256^3).  The remaining (pose-align feature) columns fold as usual.
"""
from __future__ import annotations

import numpy as np
import torch

from . import packer


def latent_layers(topo):
    """The (at most two) layers whose input contains the latent: layer 0 and the skip layer."""
    skips = [int(l) for l in topo.latent_in if 1 <= int(l) < topo.n_layers]
    if len(skips) > 1:
        raise ValueError("PixelAlign supports one skip connection (latent_in) besides the input layer")
    return [0] + skips


def setup(topo, feat_map, specs, mano_results, cam_intr, affine, feature_mode, device):
    """Per-sample PixelAlign state for asdf_simt_eval.

    feat_map: [1, C, fh, fw] (C == decoder latent size); mano_results['joints'] [1, J, 3] (root = joint 0);
    cam_intr [1, 3, 4]; affine = packer.embedding_affine(...) of the sample (grid / xyz queries) -- ignored in
    feature mode, where the query rows already ARE the embedded features.
    -> dict(maps f32 CUDA [branches][2][fh*fw + 1][npad], npad, fh, fw, layers, point_affine[12], cam[12], image_size)"""
    F = torch.as_tensor(feat_map).detach().to(device=device, dtype=torch.float32)
    if F.dim() != 4 or F.shape[0] != 1 or F.shape[1] != topo.latent_size:
        raise ValueError(f"PixelAlign: the latent must be a feature map [1, {topo.latent_size}, fh, fw], got {tuple(F.shape)}")
    if cam_intr is None or mano_results is None or "joints" not in mano_results:
        raise ValueError("PixelAlign needs cam_intr [1,3,4] and mano_results['joints'] (utils/utils.py:537,545)")
    fh, fw = int(F.shape[2]), int(F.shape[3])
    L = topo.latent_size
    mean = F.mean(3).mean(2)[0]                                           # utils/utils.py:555, fp32 like the reference
    cols = torch.cat([F[0].reshape(L, fh * fw), mean[:, None]], 1).double()   # [C, fh fw + 1]
    layers = latent_layers(topo)
    npad = None
    maps = []
    for tag, prefix in topo.branches:
        nf = len(packer.branch_feature_index(topo, tag))
        d0 = L + nf
        per = []
        for slot in range(2):
            if slot >= len(layers):
                per.append(None)
                continue
            W = topo.layers[prefix][layers[slot]][0]
            h = 0 if layers[slot] == 0 else W.shape[1] - d0
            Wz = torch.as_tensor(np.ascontiguousarray(W[:, h:h + L]), dtype=torch.float64, device=device)
            n = Wz.shape[0]
            npad = (n + 7) // 8 * 8 if npad is None else npad
            if (n + 7) // 8 * 8 != npad:
                raise ValueError("PixelAlign: the layers that see the latent must have equal widths")
            G = torch.zeros((fh * fw + 1, npad), dtype=torch.float32, device=device)
            G[:, :n] = (Wz @ cols).T.float()
            per.append(G)
        maps.append(torch.stack([g if g is not None else torch.zeros_like(per[0]) for g in per]))
    maps = torch.stack(maps).contiguous()
    s = float(specs["SdfScaleFactor"])
    root = mano_results["joints"].detach().double().cpu().numpy()[0, 0]
    P = np.zeros((3, 4))
    if feature_mode:
        P[:, :3] = np.eye(3) * (2.0 / s)                                  # queries[:, :3] * 2 / s + root, utils/utils.py:539
        P[:, 3] = root
    else:
        A, c = affine
        P[:, :3] = A[:3] * (2.0 / s)                                      # first three embedded features of xyz
        P[:, 3] = c[:3] * (2.0 / s) + root
    cam = torch.as_tensor(cam_intr).detach().double().cpu().numpy().reshape(3, 4)
    return dict(maps=maps, npad=int(npad), fh=fh, fw=fw, layers=layers + [-1] * (2 - len(layers)),
                point_affine=P.reshape(-1), cam=cam.reshape(-1), image_size=float(specs["ImageSize"][0]))
