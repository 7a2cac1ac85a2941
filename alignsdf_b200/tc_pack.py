"""Packing for the tcgen05 kernel (csrc/k1_tc.cu): static weight streams of both precision kinds, the
float64 arrays the device-side bind (csrc/bind.cu) folds a sample from, and a numpy statement of the
per-sample block the bind kernel writes (tests compare the two; the product never packs on the host).

Static stream (bytes): weight tiles ``[decoder][cta rank][tiles]`` of 8 KiB, N block after N block
    L1: nb = 0..1, kc = 0..7      rows n = 128 nb + 64 c + r   k = 64 kc + kk      (W1 padded to [256,512])
    L2: nb = 0..3, kc = 0..3                                                       (W2[:, :h] padded to [512,256])
    L3: nb = 0..3, j  = 0..7      kc = j, except the LAST N block: kc = (j + 4) % 8 -- x3's chunk c lives at
                                  K position (c + 4) % 8 and that block walks the positions in natural order so
                                  that positions 0..3 are released early for the next instance's layer-0
                                  epilogues (k1_tc.cu)
and inside an N block
    F16X3   first the tiles of its CORRECTION phase, chunk after chunk, then the hi tiles of its MAIN phase (the
            kernel accumulates the small correction products first: the tensor core truncates its accumulator
            after every UMMA, which costs the least while the accumulator is still tiny);
    F16_F8  chunk after chunk the (hi tile, correction tile) pair: main and correction UMMAs alternate (TMEM-A /
            SMEM-A forms; back-to-back SMEM-A UMMAs would saturate the shared-memory operand path, and this kind's
            error is the e4m3 significand, not the accumulator's truncation):
    hi tile     shared-memory image (K-major, 128B swizzle) of 64 rows x 64 k of  hi16(s_l W_l)
    correction  F16X3:  the hi tile (for lo16(x).hi16(W)) and the same image of lo16(s_l W_l) = fp16(s_l W_l - hi16(s_l W_l))
                F16_F8: one tile of 64 rows x 128 B: bytes 0..63  = e4m3(2^-10 s_l W_l[k]),
                                                     bytes 64..127 = e4m3((s_l W_l - hi16(s_l W_l))[k])
(3 / 2 tiles per chunk: 192 / 128 tiles per decoder and rank), followed by 2 x 520 floats: w4[512] | b4, 1/s1, 1/s2, 1/(t s3), pad  -- per decoder (SeparateDecoder) or per
output of the one MLP (CombinedDecoder).  s_l are powers of two with max|s_l W_l| in [8192, 16384); t is the
power of two the kernel keeps its activations multiplied by (16 for F16X3, 1 for F16_F8).

Per-sample block: "P tiles" ``[decoder][cta rank][14]`` of 8 KiB -- one per N block (4 of layer 0,
2 of layer 1, 4 + 4 of layers 2 and 3) holding, in k columns 0..15 of each feature row,
    [Mx_h My_h Mz_h B_h | Mx_h My_h Mz_h 0 | Mx_l My_l Mz_l B_l | 0 0 0 0]
with M' = (S_l / cp) M, B' = (S_l / c1) B split into fp16 hi + lo, so that one K=16 UMMA against the
point operand [cp p_h, c1, cp p_l, 0, cp p_h, c1, 0...] adds S_l (M.p + B) to the accumulator;
then 16 floats: inv0[2] = t / S_0, cp, c1.   (S_l = t s_l for l = 1..3, S_0 chosen per sample.)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

F16X3, F16_F8 = _lib.TC_F16X3, _lib.TC_F16_F8
ACT_SCALE = {F16X3: 16.0, F16_F8: 1.0}
ROWS, TK = 64, 64
TILE_ELEMS = ROWS * TK
TILE_BYTES = TILE_ELEMS * 2
CHUNKS = 64              # 64-wide K chunks of all N blocks of a decoder: 16 (L1) + 16 (L2) + 32 (L3)
TILES_PER_CHUNK = {_lib.TC_F16X3: 3, _lib.TC_F16_F8: 2}
P_TILES = 14
STATIC_PARAM_FLOATS = 520
SAMPLE_TILE_BYTES = 2 * 2 * P_TILES * TILE_BYTES
SAMPLE_BYTES = SAMPLE_TILE_BYTES + 64
F16_SAFE = 16384.0
LO_SCALE = 1024.0        # F16_F8 activations: lo8 = e4m3(LO_SCALE lo(x));   weights: W8 = e4m3(s W / LO_SCALE)
FP8_LIMIT = 448.0        # F16_F8: x8 = e4m3(hi16(x)) saturates beyond this -> the kernel raises its status flag
MAX_POINT_DIM = 64


def supported(topo) -> bool:
    """The shipped topology: 5-layer MLPs, 512 wide, skip into layer 2, plain final tanh -- two of them
    (SeparateDecoder) or one with two outputs (CombinedDecoder without xyz_in_all)."""
    if topo.n_layers != 5 or topo.pre_tanh or topo.xyz_in_all or tuple(topo.latent_in) != (2,):
        return False
    if any(ln is not None for lns in topo.layer_norms.values() for ln in lns):
        return False                      # LayerNorm decoders run on the generic kernel
    n_out = 1 if topo.kind == "separate" else 2
    if len(topo.branches) != (2 if topo.kind == "separate" else 1):
        return False
    for _, prefix in topo.branches:
        ls = topo.layers[prefix]
        d0 = ls[0][0].shape[1]
        h = 512 - d0
        if not (0 < h <= 256) or d0 - topo.latent_size > MAX_POINT_DIM or topo.latent_size < 1:
            return False
        if [w.shape for w, _ in ls] != [(512, d0), (h, 512), (512, h + d0), (512, 512), (n_out, 512)]:
            return False
    return True


def _pow2_scale(w: np.ndarray) -> float:
    m = float(np.abs(w).max())
    if m == 0.0:
        return 1.0
    return float(2.0 ** np.floor(np.log2(16384.0 / m)))


def _padded(W, rows, cols):
    out = np.zeros((rows, cols), np.float64)
    out[:W.shape[0], :W.shape[1]] = W
    return out


def split_f16(w64: np.ndarray):
    hi = w64.astype(np.float16)
    lo = (w64 - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


_SWZ = _SWZ8 = None


def _swz():
    """element offset of (r, k) inside a 128B-swizzled K-major [64 rows][64 fp16] tile."""
    global _SWZ
    if _SWZ is None:
        r = np.arange(ROWS)[:, None]
        k = np.arange(TK)[None, :]
        _SWZ = (((r // 8) * 1024 + (r % 8) * 128 + (((k // 8) ^ (r % 8)) * 16) + (k % 8) * 2) // 2).astype(np.int64)
    return _SWZ


def _swz8():
    """byte offset of byte column b of row r inside a 128B-swizzled [64 rows][128 B] tile."""
    global _SWZ8
    if _SWZ8 is None:
        r = np.arange(ROWS)[:, None]
        b = np.arange(128)[None, :]
        _SWZ8 = ((r // 8) * 1024 + (r % 8) * 128 + (((b // 16) ^ (r % 8)) * 16) + (b % 16)).astype(np.int64)
    return _SWZ8


def swizzle_tile(mat):
    out = np.zeros(TILE_ELEMS, np.float16)
    out[_swz().reshape(-1)] = np.asarray(mat, np.float16).reshape(-1)
    return out


def unswizzle_tile(flat):
    return np.asarray(flat, np.float16)[_swz()]


def swizzle_tile8(mat_u8):
    out = np.zeros(ROWS * 128, np.uint8)
    out[_swz8().reshape(-1)] = np.asarray(mat_u8, np.uint8).reshape(-1)
    return out


def unswizzle_tile8(flat_u8):
    return np.asarray(flat_u8, np.uint8)[_swz8()]


def e4m3_encode(x: np.ndarray) -> np.ndarray:
    """float -> e4m3 bytes, round to nearest even, saturating at +-448 (== cvt.rn.satfinite.e4m3x2.f32)."""
    t = torch.from_numpy(np.clip(np.asarray(x, np.float64), -448.0, 448.0).astype(np.float32))
    return t.to(torch.float8_e4m3fn).view(torch.uint8).numpy()


def e4m3_decode(b: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(b, np.uint8)).view(torch.float8_e4m3fn).to(torch.float32).numpy()


def _layer_mats(topo, prefix):
    ls = topo.layers[prefix]
    h = ls[1][0].shape[0]
    return _padded(ls[1][0], 256, 512), _padded(ls[2][0][:, :h], 512, 256), ls[3][0], h


def weight_scales(topo) -> np.ndarray:
    """[2][3] power-of-two scales of layers 1..3 per decoder (row 1 unused for a CombinedDecoder)."""
    scales = np.ones((2, 3))
    for d, (_, prefix) in enumerate(topo.branches):
        W1, W2, W3, _ = _layer_mats(topo, prefix)
        scales[d] = (_pow2_scale(W1), _pow2_scale(W2), _pow2_scale(W3))
    return scales


def pack_static_numpy(topo, kind):
    """-> (uint8 stream, scales [2][3])"""
    nd = len(topo.branches)
    t = ACT_SCALE[kind]
    ntiles = CHUNKS * TILES_PER_CHUNK[kind]
    stream = np.zeros((nd, 2, ntiles, TILE_BYTES), np.uint8)
    params = np.zeros((2, STATIC_PARAM_FLOATS), np.float32)
    scales = weight_scales(topo)
    for d, (_, prefix) in enumerate(topo.branches):
        W1, W2, W3, _ = _layer_mats(topo, prefix)
        s = scales[d]
        for c in range(2):
            i = 0
            for W, sc, nbs, kcs in ((W1, s[0], 2, 8), (W2, s[1], 4, 4), (W3, s[2], 4, 8)):
                for nb in range(nbs):
                    r0 = 128 * nb + 64 * c
                    his = []
                    for j in range(kcs):                 # correction phase
                        kc = (j + 4) % 8 if (W is W3 and nb == nbs - 1) else j
                        blk = (sc * W[r0:r0 + 64, 64 * kc:64 * kc + 64]).astype(np.float64)
                        hi = blk.astype(np.float16)
                        lo = blk - hi.astype(np.float64)
                        his.append(swizzle_tile(hi).view(np.uint8))
                        if kind == F16_F8:               # interleaved order: (hi, correction) pair per chunk
                            stream[d, c, i] = his[-1]
                            stream[d, c, i + 1] = swizzle_tile8(
                                np.concatenate([e4m3_encode(blk / LO_SCALE), e4m3_encode(lo)], 1))
                            i += 2
                        else:
                            stream[d, c, i] = his[-1]
                            stream[d, c, i + 1] = swizzle_tile(lo.astype(np.float16)).view(np.uint8)
                            i += 2
                    for j in range(kcs if kind != F16_F8 else 0):    # main phase (F16X3 only)
                        stream[d, c, i] = his[j]
                        i += 1
            assert i == ntiles
        W4, b4 = topo.layers[prefix][4]
        for o in range(W4.shape[0]):                     # one row per decoder, or the two outputs of the one MLP
            p = params[d + o]
            p[:512] = W4[o].astype(np.float32)
            p[512:516] = [b4[o], 1.0 / s[0], 1.0 / s[1], 1.0 / (s[2] * t)]
    raw = np.concatenate([stream.reshape(-1), params.reshape(-1).view(np.uint8)])
    return raw, scales


def pack_static(topo, kind, device) -> torch.Tensor:
    raw, _ = pack_static_numpy(topo, kind)
    expect = _lib.lib().asdf_tc_static_bytes(kind, len(topo.branches))
    if raw.nbytes != expect:
        raise _lib.AsdfError(f"packed weight stream is {raw.nbytes} B, library expects {expect} B")
    return torch.from_numpy(raw).to(device)


def bind_static_numpy(topo) -> np.ndarray:
    """float64 arrays of csrc/bind.cu, per decoder: Wz[2][512][L] | Wf[2][512][64] | b[4][512]."""
    from . import packer
    L = topo.latent_size
    out = []
    for tag, prefix in topo.branches:
        ls = topo.layers[prefix]
        nf = len(packer.branch_feature_index(topo, tag))
        h = ls[1][0].shape[0]
        W0, W2 = ls[0][0], ls[2][0]
        wz = np.stack([W0[:, :L], W2[:, h:h + L]])
        wf = np.zeros((2, 512, MAX_POINT_DIM))
        wf[0, :, :nf] = W0[:, L:]
        wf[1, :, :nf] = W2[:, h + L:]
        b = np.zeros((4, 512))
        b[0], b[2], b[3] = ls[0][1], ls[2][1], ls[3][1]
        b[1, :h] = ls[1][1]
        out.append(np.concatenate([wz.reshape(-1), wf.reshape(-1), b.reshape(-1)]))
    return np.concatenate(out).astype(np.float64)


# ----------------------------------------------------------------------------
# numpy statement of the per-sample block (tests; the product path is asdf_tc_bind)
# ----------------------------------------------------------------------------
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


def _pad1(b, n):
    out = np.zeros(n, np.float64)
    out[:b.shape[0]] = b
    return out


def pack_sample_numpy(branches, scales, p_absmax=2.0, kind=F16X3):
    """Per-sample block from the folded branches (packer.fold_decoder, xyz mode)."""
    act_scale = ACT_SCALE[kind]
    terms = []
    for d, br in enumerate(branches):
        L = br.layers
        terms.append([(None, L[0].M.astype(np.float64), L[0].B.astype(np.float64)),
                      (act_scale * scales[d][0], None, _pad1(L[1].B, 256)),
                      (act_scale * scales[d][1], L[2].M.astype(np.float64), L[2].B.astype(np.float64)),
                      (act_scale * scales[d][2], None, L[3].B.astype(np.float64))])
    cp, c1, S0 = choose_point_scales(terms, p_absmax)
    rows = np.zeros((2, 2, P_TILES, ROWS, TK), np.float16)      # [decoder][cta rank][N block][row][k], unswizzled
    for d in range(len(branches)):
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
    for d in range(len(branches)):
        scal[d] = act_scale / S0[d]
    scal[2], scal[3] = cp, c1
    return np.concatenate([tiles.reshape(-1).view(np.uint8), scal.view(np.uint8)]), dict(cp=cp, c1=c1, S0=S0)
