"""Packing for the v2 tcgen05 kernel (csrc/k1_tc2.cu).

Static stream (bytes): main weight tiles ``[decoder][cta rank][128 tiles]`` of 8 KiB, each the
shared-memory image (K-major, 128B swizzle) of 64 weight rows x 64 k of fp16, alternating hi / lo
halves of ``s_l * W_l``:
    L1: nb = 0..1, kc = 0..7 -> (hi, lo)   rows n = 128 nb + 64 c + r   k = 64 kc + kk
    L2: nb = 0..3, kc = 0..3               (W2[:, :h])
    L3: nb = 0..3, j  = 0..7
followed by 2 x 520 floats: w4[512] | b4, 1/s1, 1/s2, 1/(t s3), pad.

Per-sample block: "P tiles" ``[decoder][cta rank][14]`` of 8 KiB -- one per N block (4 of layer 0,
2 of layer 1, 4 + 4 of layers 2 and 3) holding, in k columns 0..15 of each feature row,
    [Mx_h My_h Mz_h B_h | Mx_h My_h Mz_h 0 | Mx_l My_l Mz_l B_l | 0 0 0 0]
with M' = (S_l / cp) M, B' = (S_l / c1) B split into fp16 hi + lo, so that one K=16 UMMA against the
point operand [cp p_h, c1, cp p_l, 0, cp p_h, c1, 0...] adds S_l (M.p + B) to the accumulator;
then 16 floats: inv0[2] = t / S_0, cp, c1.   (S_l = t s_l for l = 1..3, S_0 chosen per sample.)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .tc_pack import ACT_SCALE, _padded, _pow2_scale, split_f16, supported  # noqa: F401

ROWS, TK = 64, 64
TILE_ELEMS = ROWS * TK
TILE_BYTES = TILE_ELEMS * 2
MAIN_TILES = 128
P_TILES = 14
STATIC_PARAM_FLOATS = 520
F16_SAFE = 16384.0

_SWZ = None


def _swz():
    global _SWZ
    if _SWZ is None:
        r = np.arange(ROWS)[:, None]
        k = np.arange(TK)[None, :]
        _SWZ = (((r // 8) * 1024 + (r % 8) * 128 + (((k // 8) ^ (r % 8)) * 16) + (k % 8) * 2) // 2).astype(np.int64)
    return _SWZ


def swizzle_tile(mat):
    out = np.zeros(TILE_ELEMS, np.float16)
    out[_swz().reshape(-1)] = np.asarray(mat, np.float16).reshape(-1)
    return out


def unswizzle_tile(flat):
    return np.asarray(flat, np.float16)[_swz()]


def pack_static_numpy(topo):
    """-> (uint8 stream, scales [2][3])"""
    stream = np.zeros((2, 2, MAIN_TILES, TILE_ELEMS), np.float16)
    params = np.zeros((2, STATIC_PARAM_FLOATS), np.float32)
    scales = np.ones((2, 3))
    for d, (_, prefix) in enumerate(topo.branches):
        ls = topo.layers[prefix]
        h = ls[1][0].shape[0]
        W1 = _padded(ls[1][0], 256, 512)
        W2 = _padded(ls[2][0][:, :h], 512, 256)
        W3 = ls[3][0]
        s = (_pow2_scale(W1), _pow2_scale(W2), _pow2_scale(W3))
        scales[d] = s
        for c in range(2):
            i = 0

            def put(block):
                nonlocal i
                hi, lo = split_f16(block)
                stream[d, c, i] = swizzle_tile(hi)
                stream[d, c, i + 1] = swizzle_tile(lo)
                i += 2
            for W, sc, nbs, kcs in ((W1, s[0], 2, 8), (W2, s[1], 4, 4), (W3, s[2], 4, 8)):
                for nb in range(nbs):
                    r0 = 128 * nb + 64 * c
                    for kc in range(kcs):
                        put(sc * W[r0:r0 + 64, 64 * kc:64 * kc + 64])
            assert i == MAIN_TILES
        p = params[d]
        p[:512] = ls[4][0][0].astype(np.float32)
        p[512:516] = [ls[4][1][0], 1.0 / s[0], 1.0 / s[1], 1.0 / (s[2] * ACT_SCALE)]
    raw = np.concatenate([stream.reshape(-1).view(np.uint8), params.reshape(-1).view(np.uint8)])
    return raw, scales


def _pow2_floor(x):
    return float(2.0 ** np.floor(np.log2(x)))


def choose_point_scales(layer_terms, p_absmax):
    """layer_terms: list over decoders of [(S_l or None, M [n,3] or None, B [n])] for l = 0..3.
    Returns (cp, c1, S0 per decoder) as powers of two such that every fp16 operand stays in range;
    raises ValueError when no choice exists (caller falls back to the generic kernel)."""
    p_absmax = max(float(p_absmax), 1e-3)
    cp_max = _pow2_floor(60000.0 / p_absmax)
    cp_min, c1_min = 1.0, 1.0
    for terms in layer_terms:
        for S, M, B in terms[1:]:
            if M is not None and np.abs(M).max() > 0:
                cp_min = max(cp_min, S * float(np.abs(M).max()) / F16_SAFE)
            if np.abs(B).max() > 0:
                c1_min = max(c1_min, S * float(np.abs(B).max()) / F16_SAFE)
    cp = 2.0 ** np.ceil(np.log2(cp_min))
    c1 = 2.0 ** np.ceil(np.log2(c1_min))
    if cp > cp_max or c1 > 32768.0:
        raise ValueError(f"point/bias terms do not fit fp16 operands (cp in [{cp_min:.3g}, {cp_max:.3g}], c1 >= {c1_min:.3g})")
    cp = max(cp, min(cp_max, 1024.0))        # prefer a large cp: more headroom for the lo part of p
    c1 = max(c1, 1024.0)
    S0 = []
    for terms in layer_terms:
        _, M, B = terms[0]
        lim = min(F16_SAFE * cp / max(float(np.abs(M).max()), 1e-30), F16_SAFE * c1 / max(float(np.abs(B).max()), 1e-30))
        S0.append(_pow2_floor(lim))
    return float(cp), float(c1), S0


def pack_sample_numpy(branches, scales, p_absmax=1.25, act_scale=ACT_SCALE):
    """Per-sample block from the folded branches (packer.fold_decoder, xyz mode); ``act_scale`` is the
    power of two t the kernel keeps its activations multiplied by (16 for k1_tc2.cu, 1 for k1_tc3.cu)."""
    terms = []
    for d, br in enumerate(branches):
        L = br.layers
        terms.append([(None, L[0].M.astype(np.float64), L[0].B.astype(np.float64)),
                      (act_scale * scales[d][0], None, _pad1(L[1].B, 256)),
                      (act_scale * scales[d][1], L[2].M.astype(np.float64), L[2].B.astype(np.float64)),
                      (act_scale * scales[d][2], None, L[3].B.astype(np.float64))])
    cp, c1, S0 = choose_point_scales(terms, p_absmax)
    rows = np.zeros((2, 2, P_TILES, ROWS, TK), np.float16)      # [decoder][cta rank][N block][row][k], unswizzled
    for d in range(2):
        g = 0
        for l, (S, M, B) in enumerate(terms[d]):
            S = S0[d] if l == 0 else S
            n = B.shape[0]
            Ms = np.zeros((n, 3)) if M is None else (S / cp) * M
            Bs = (S / c1) * B
            if max(np.abs(Ms).max(), np.abs(Bs).max()) > 60000:
                raise ValueError("point/bias operand overflows fp16")
            mh, ml = split_f16(Ms)
            bh, bl = split_f16(Bs)
            full = np.zeros((n, TK), np.float16)
            full[:, 0:3], full[:, 3] = mh, bh
            full[:, 4:7] = mh
            full[:, 8:11], full[:, 11] = ml, bl
            nbs = n // 128
            rows[d, :, g:g + nbs] = full.reshape(nbs, 2, ROWS, TK).transpose(1, 0, 2, 3)
            g += nbs
        assert g == P_TILES
    tiles = np.zeros((2, 2, P_TILES, TILE_ELEMS), np.float16)   # all 56 tiles swizzled in one scatter
    tiles[..., _swz().reshape(-1)] = rows.reshape(2, 2, P_TILES, TILE_ELEMS)
    scal = np.zeros(16, np.float32)
    scal[0], scal[1] = act_scale / S0[0], act_scale / S0[1]
    scal[2], scal[3] = cp, c1
    return np.concatenate([tiles.reshape(-1).view(np.uint8), scal.view(np.uint8)]), dict(cp=cp, c1=c1, S0=S0)


def _pad1(b, n):
    out = np.zeros(n, np.float64)
    out[:b.shape[0]] = b
    return out


def pack_static(engine) -> torch.Tensor:
    raw, scales = pack_static_numpy(engine.topo)
    expect = _lib.lib().asdf_tc2_static_bytes()
    if raw.nbytes != expect:
        raise _lib.AsdfError(f"packed v2 weight stream is {raw.nbytes} B, library expects {expect} B")
    engine.tc2_scales = scales
    return torch.from_numpy(raw).to(engine.device)


@dataclass
class Tc2Bound:
    sample: torch.Tensor
    info: dict


def bind(engine, branches, p_absmax=1.25) -> Tc2Bound:
    raw, info = pack_sample_numpy(branches, engine.tc2_scales, p_absmax)
    assert raw.nbytes == _lib.lib().asdf_tc2_sample_bytes()
    info["p_absmax"] = p_absmax
    return Tc2Bound(torch.from_numpy(raw).to(engine.device, non_blocking=True), info)
