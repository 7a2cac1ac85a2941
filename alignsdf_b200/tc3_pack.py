"""Packing for the v3 tcgen05 kernel (csrc/k1_tc3.cu): fp16 main product + fp8 (e4m3) corrections.

Static stream (bytes): main weight tiles ``[decoder][cta rank][128 tiles]`` of 8 KiB in the v2 order
(tc2_pack.py: L1 nb 0..1 x kc 0..7, L2 nb 0..3 x kc 0..3, L3 nb 0..3 x j 0..7) -- except that the LAST N
block of layer 3 stores its K chunks in the order kc = (j + 4) % 8: x3's chunk c lives at K position
(c + 4) % 8, and that block walks the positions 0..7 in natural order so that positions 0..3 are released
early for the next decoder's layer-0 epilogues (k1_tc3.cu) -- alternating
    fp16 tile   shared-memory image (K-major, 128B swizzle) of 64 rows x 64 k of  hi16(s_l W_l)
    fp8 tile    shared-memory image of 64 rows x 128 B: bytes 0..63  = e4m3(2^-10 s_l W_l[k]),
                                                       bytes 64..127 = e4m3((s_l W_l - hi16(s_l W_l))[k])
followed by 2 x 520 floats: w4[512] | b4, 1/s1, 1/s2, 1/s3, pad.

The per-sample block (P tiles: biases and pose-align point terms as K=16 fp16 products) has the v2
layout with activation scale t = 1.  The kernel's activation operands are hi16(x) in tensor memory
and, per 64-k slot row, e4m3(2^10 (x - hi16(x))) | e4m3(hi16(x)) in shared memory, so every product
carries the scale s_l.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from . import tc2_pack
from .tc2_pack import (MAIN_TILES, P_TILES, ROWS, STATIC_PARAM_FLOATS, TILE_BYTES, TILE_ELEMS, TK,  # noqa: F401
                       swizzle_tile, unswizzle_tile)
from .tc_pack import _padded, _pow2_scale, supported  # noqa: F401

ACT_SCALE = 1.0          # t: activations are not pre-scaled (no fp16 lo half that could go subnormal)
LO_SCALE = 1024.0        # activations: lo8 = e4m3(LO_SCALE lo(x));   weights: W8  = e4m3(s W / LO_SCALE)
FP8_LIMIT = 448.0        # x8 = e4m3(hi16(x)) saturates beyond this -> the kernel raises its status flag

_SWZ8 = None


def _swz8():
    """byte offset of byte column b of row r inside a 128B-swizzled [64 rows][128 B] tile."""
    global _SWZ8
    if _SWZ8 is None:
        r = np.arange(ROWS)[:, None]
        b = np.arange(128)[None, :]
        _SWZ8 = ((r // 8) * 1024 + (r % 8) * 128 + (((b // 16) ^ (r % 8)) * 16) + (b % 16)).astype(np.int64)
    return _SWZ8


def e4m3_encode(x: np.ndarray) -> np.ndarray:
    """float -> e4m3 bytes, round to nearest even, saturating at +-448 (== cvt.rn.satfinite.e4m3x2.f32)."""
    t = torch.from_numpy(np.clip(np.asarray(x, np.float64), -448.0, 448.0).astype(np.float32))
    return t.to(torch.float8_e4m3fn).view(torch.uint8).numpy()


def e4m3_decode(b: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(b, np.uint8)).view(torch.float8_e4m3fn).to(torch.float32).numpy()


def swizzle_tile8(mat_u8: np.ndarray) -> np.ndarray:
    """[64, 128] uint8 -> flat [8192] uint8 shared-memory image."""
    out = np.zeros(ROWS * 128, np.uint8)
    out[_swz8().reshape(-1)] = np.asarray(mat_u8, np.uint8).reshape(-1)
    return out


def unswizzle_tile8(flat_u8: np.ndarray) -> np.ndarray:
    return np.asarray(flat_u8, np.uint8)[_swz8()]


def pack_static_numpy(topo):
    """-> (uint8 stream, scales [2][3])"""
    stream = np.zeros((2, 2, MAIN_TILES, TILE_BYTES), np.uint8)
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
            for W, sc, nbs, kcs in ((W1, s[0], 2, 8), (W2, s[1], 4, 4), (W3, s[2], 4, 8)):
                for nb in range(nbs):
                    r0 = 128 * nb + 64 * c
                    for j in range(kcs):
                        kc = (j + 4) % 8 if (W is W3 and nb == nbs - 1) else j
                        blk = (sc * W[r0:r0 + 64, 64 * kc:64 * kc + 64]).astype(np.float64)
                        hi = blk.astype(np.float16)
                        lo = blk - hi.astype(np.float64)
                        stream[d, c, i] = swizzle_tile(hi).view(np.uint8)
                        stream[d, c, i + 1] = swizzle_tile8(
                            np.concatenate([e4m3_encode(blk / LO_SCALE), e4m3_encode(lo)], 1))
                        i += 2
            assert i == MAIN_TILES
        p = params[d]
        p[:512] = ls[4][0][0].astype(np.float32)
        p[512:516] = [ls[4][1][0], 1.0 / s[0], 1.0 / s[1], 1.0 / (s[2] * ACT_SCALE)]
    raw = np.concatenate([stream.reshape(-1), params.reshape(-1).view(np.uint8)])
    return raw, scales


def pack_static(engine) -> torch.Tensor:
    raw, scales = pack_static_numpy(engine.topo)
    expect = _lib.lib().asdf_tc3_static_bytes()
    if raw.nbytes != expect:
        raise _lib.AsdfError(f"packed v3 weight stream is {raw.nbytes} B, library expects {expect} B")
    engine.tc3_scales = scales
    return torch.from_numpy(raw).to(engine.device)


def pack_sample_numpy(branches, scales, p_absmax=1.25):
    return tc2_pack.pack_sample_numpy(branches, scales, p_absmax, act_scale=ACT_SCALE)


@dataclass
class Tc3Bound:
    sample: torch.Tensor
    info: dict


def bind(engine, branches, p_absmax=1.25) -> Tc3Bound:
    raw, info = pack_sample_numpy(branches, engine.tc3_scales, p_absmax)
    assert raw.nbytes == _lib.lib().asdf_tc3_sample_bytes()
    info["p_absmax"] = p_absmax
    return Tc3Bound(torch.from_numpy(raw).to(engine.device, non_blocking=True), info)
