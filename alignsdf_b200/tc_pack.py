"""Packing for the tcgen05 kernel (csrc/k1_tc.cu): fp16 hi/lo split weights, pre-swizzled into the
exact shared-memory image of each 128x64 B tile, in the order the kernel's producer streams them.

Stream layout (bytes):  [decoder d][cta rank c][tile i]  with 64 tiles of 16 KiB per (d, c):
    L1: kc = 0..7            -> (hi, lo)      rows n = 128c + r           k = 64kc + kk   (W1, [h,512])
    L2: nb = 0..1, kc = 0..3 -> (hi, lo)      rows n = 256nb + 128c + r   k = 64kc + kk   (W2[:, :h])
    L3: nb = 0..1, j = 0..7  -> (hi, lo)      rows n = 256nb + 128c + r   k = 64j + kk    (W3)
followed by 2 x 1288 floats of static parameters: b1*t [256] | (b3, w4) [512][2] | b4, 1/s1, 1/s2,
1/(s3 t), 4 pad.  Per-sample block (floats): per decoder M0B0*t [512][4] | M2B2*t [512][4].

Scales are powers of two (exact): activations are multiplied by t = act_scale before the fp16
split, layer-l weights by s_l chosen so max|s_l W_l| lies in [8192, 16384).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

TILE_ROWS, TILE_K = 128, 64
TILE_BYTES = TILE_ROWS * TILE_K * 2
TILES_PER_DECODER = 64
STATIC_PARAM_FLOATS = 256 + 1024 + 8
SAMPLE_FLOATS_PER_DECODER = 2 * 512 * 4
ACT_SCALE = 16.0


def supported(topo) -> bool:
    """The shipped topology: two 5-layer MLPs, 512 wide, skip into layer 2, plain final tanh."""
    if topo.kind != "separate" or topo.n_layers != 5 or topo.pre_tanh or topo.xyz_in_all:
        return False
    if tuple(topo.latent_in) != (2,):
        return False
    if any(ln is not None for lns in topo.layer_norms.values() for ln in lns):
        return False                      # LayerNorm decoders run on the generic kernel
    for _, prefix in topo.branches:
        ls = topo.layers[prefix]
        d0 = ls[0][0].shape[1]
        h = 512 - d0
        shapes = [w.shape for w, _ in ls]
        if not (0 < h <= 256):
            return False
        if shapes != [(512, d0), (h, 512), (512, h + d0), (512, 512), (1, 512)]:
            return False
    return True


_SWZ = None


def _swizzle_index():
    """byte offset (in fp16 elements) of element (r, k) inside a 128B-swizzled K-major tile."""
    global _SWZ
    if _SWZ is None:
        r = np.arange(TILE_ROWS)[:, None]
        k = np.arange(TILE_K)[None, :]
        off = (r // 8) * 1024 + (r % 8) * 128 + (((k // 8) ^ (r % 8)) * 16) + (k % 8) * 2
        _SWZ = (off // 2).astype(np.int64)
    return _SWZ


def swizzle_tile(mat: np.ndarray) -> np.ndarray:
    """[128, 64] fp16 -> flat [8192] fp16 shared-memory image."""
    out = np.zeros(TILE_ROWS * TILE_K, np.float16)
    out[_swizzle_index().reshape(-1)] = np.asarray(mat, np.float16).reshape(-1)
    return out


def unswizzle_tile(flat: np.ndarray) -> np.ndarray:
    return np.asarray(flat, np.float16)[_swizzle_index()]


def split_f16(w64: np.ndarray):
    hi = w64.astype(np.float16)
    lo = (w64 - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def _pow2_scale(w: np.ndarray) -> float:
    m = float(np.abs(w).max())
    if m == 0.0:
        return 1.0
    return float(2.0 ** np.floor(np.log2(16384.0 / m)))


def _padded(w, rows, cols):
    out = np.zeros((rows, cols), np.float64)
    out[:w.shape[0], :w.shape[1]] = w
    return out


def pack_static_numpy(topo):
    """-> (uint8 array of asdf_tc_static_bytes() bytes, w_scale [2][3], h [2])."""
    stream = np.zeros((2, 2, TILES_PER_DECODER, TILE_ROWS * TILE_K), np.float16)
    params = np.zeros((2, STATIC_PARAM_FLOATS), np.float32)
    scales = np.ones((2, 3))
    hs = []
    for d, (_, prefix) in enumerate(topo.branches):
        ls = topo.layers[prefix]
        h = ls[1][0].shape[0]
        hs.append(h)
        W1 = _padded(ls[1][0], 256, 512)
        W2 = _padded(ls[2][0][:, :h], 512, 256)
        W3 = ls[3][0]
        s1, s2, s3 = _pow2_scale(W1), _pow2_scale(W2), _pow2_scale(W3)
        scales[d] = (s1, s2, s3)
        for c in range(2):
            i = 0

            def put(block):
                nonlocal i
                hi, lo = split_f16(block)
                stream[d, c, i] = swizzle_tile(hi)
                stream[d, c, i + 1] = swizzle_tile(lo)
                i += 2
            for kc in range(8):
                put(s1 * W1[128 * c:128 * c + 128, 64 * kc:64 * kc + 64])
            for nb in range(2):
                for kc in range(4):
                    put(s2 * W2[256 * nb + 128 * c:256 * nb + 128 * c + 128, 64 * kc:64 * kc + 64])
            for nb in range(2):
                for j in range(8):
                    put(s3 * W3[256 * nb + 128 * c:256 * nb + 128 * c + 128, 64 * j:64 * j + 64])
            assert i == TILES_PER_DECODER
        p = params[d]
        p[:h] = (ACT_SCALE * ls[1][1]).astype(np.float32)
        bw = np.stack([ls[3][1], ls[4][0][0]], 1).astype(np.float32)        # (b3[n], w4[n])
        p[256:256 + 1024] = bw.reshape(-1)
        p[1280:1284] = [ls[4][1][0], 1.0 / s1, 1.0 / s2, 1.0 / (s3 * ACT_SCALE)]
    raw = np.concatenate([stream.reshape(-1).view(np.uint8), params.reshape(-1).view(np.uint8)])
    return raw, scales, hs


def pack_sample_numpy(branches) -> np.ndarray:
    """Per-sample block from the folded branches (packer.fold_decoder, xyz mode)."""
    out = np.zeros((2, 2, 512, 4), np.float32)
    for d, br in enumerate(branches):
        for slot, layer in ((0, br.layers[0]), (1, br.layers[2])):
            out[d, slot, :, :3] = ACT_SCALE * layer.M
            out[d, slot, :, 3] = ACT_SCALE * layer.B
    return out.reshape(-1)


def pack_static(engine) -> torch.Tensor:
    raw, scales, hs = pack_static_numpy(engine.topo)
    expect = _lib.lib().asdf_tc_static_bytes()
    if raw.nbytes != expect:
        raise _lib.AsdfError(f"packed weight stream is {raw.nbytes} B, library expects {expect} B")
    engine.tc_scales, engine.tc_h = scales, hs
    return torch.from_numpy(raw).to(engine.device)


@dataclass
class TcBound:
    desc: _lib.TcDesc
    sample: torch.Tensor


def bind(engine, branches) -> TcBound:
    if any(br.point_dim != 3 for br in branches):
        raise _lib.AsdfError("tensor-core path needs xyz-folded weights")
    samp = pack_sample_numpy(branches)
    assert samp.size == _lib.lib().asdf_tc_sample_floats()
    d = _lib.TcDesc()
    for b in range(2):
        d.h[b] = int(engine.tc_h[b])
        for l in range(3):
            d.w_scale[b][l] = float(engine.tc_scales[b][l])
    d.act_scale = ACT_SCALE
    d.branch_stride = 2 * TILES_PER_DECODER * TILE_BYTES
    d.debug_dev = None
    return TcBound(d, torch.from_numpy(samp).to(engine.device, non_blocking=True))
