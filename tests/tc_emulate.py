"""Numpy emulation of csrc/k1_tc.cu's arithmetic, driven by the *packed* buffers (test helper).

It un-swizzles the streamed tiles exactly as the kernel's UMMAs consume them (tile order, A-slot
mapping, scales) and models each UMMA triple as a_hi.b_hi + a_lo.b_hi + a_hi.b_lo with exact fp16
products accumulated in (at least) fp32.  Used on the CPU to validate the packing order and the
fp16x3 precision against the reference's golden fields before any GPU time is spent."""
import numpy as np

from alignsdf_b200 import tc_pack as T


def _split(v32):
    v = np.minimum(v32, np.float32(60000.0)).astype(np.float32)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def _tiles(raw, d):
    s = raw[:2 * 2 * T.TILES_PER_DECODER * T.TILE_BYTES].view(np.float16).reshape(2, 2, T.TILES_PER_DECODER, -1)
    return [[T.unswizzle_tile(s[d, c, i]).astype(np.float64) for i in range(T.TILES_PER_DECODER)] for c in range(2)]


def _layer(tiles, first, n_blocks, kchunks, slot_of, a_hi, a_lo):
    """accumulators [P, 256*n_blocks]; a_* are dicts slot -> [P,64]."""
    P = next(iter(a_hi.values())).shape[0]
    out = np.zeros((P, 256 * n_blocks))
    i = first
    for nb in range(n_blocks):
        for kc in range(kchunks):
            s = slot_of(kc)
            for c in range(2):
                bhi, blo = tiles[c][i], tiles[c][i + 1]
                cols = slice(256 * nb + 128 * c, 256 * nb + 128 * c + 128)
                out[:, cols] += a_hi[s] @ bhi.T + a_lo[s] @ bhi.T + a_hi[s] @ blo.T
            i += 2
    return out.astype(np.float32), i


def emulate(raw_static, sample, xyz):
    """-> (sdf_hand [P], sdf_obj [P]) float32."""
    raw_static = np.asarray(raw_static, np.uint8)
    params = raw_static[2 * 2 * T.TILES_PER_DECODER * T.TILE_BYTES:].view(np.float32).reshape(2, T.STATIC_PARAM_FLOATS)
    samp = np.asarray(sample, np.float32).reshape(2, 2, 512, 4)
    p = np.asarray(xyz, np.float32)
    outs = []
    for d in range(2):
        tiles = _tiles(raw_static, d)
        b1s, b3w4 = params[d, :256], params[d, 256:1280].reshape(512, 2)
        b4, inv1, inv2, inv3 = params[d, 1280:1284]
        m0, m2 = samp[d, 0], samp[d, 1]

        def affine(m):
            # fmaf(m.x, px, fmaf(m.y, py, fmaf(m.z, pz, m.w))) -- evaluated in float64 then rounded
            return (p.astype(np.float64) @ m[:, :3].astype(np.float64).T + m[:, 3].astype(np.float64)).astype(np.float32)
        x1 = np.maximum(affine(m0), 0)
        hi, lo = _split(x1)
        a_hi = {s: hi[:, 64 * s:64 * s + 64] for s in range(8)}
        a_lo = {s: lo[:, 64 * s:64 * s + 64] for s in range(8)}
        acc1, i = _layer(tiles, 0, 1, 8, lambda kc: kc, a_hi, a_lo)
        x2 = np.maximum(acc1 * inv1 + b1s[None], 0).astype(np.float32)
        hi, lo = _split(x2)
        for s in range(4):
            a_hi[s], a_lo[s] = hi[:, 64 * s:64 * s + 64], lo[:, 64 * s:64 * s + 64]
        acc2, i = _layer(tiles, i, 2, 4, lambda kc: kc, a_hi, a_lo)
        x3 = np.maximum(acc2 * inv2 + affine(m2), 0).astype(np.float32)
        hi, lo = _split(x3)
        for f in range(8):                       # feature chunk f -> slot (f+4)%8
            s = (f + 4) % 8
            a_hi[s], a_lo[s] = hi[:, 64 * f:64 * f + 64], lo[:, 64 * f:64 * f + 64]
        acc3, i = _layer(tiles, i, 2, 8, lambda j: (j + 4) % 8, a_hi, a_lo)
        assert i == T.TILES_PER_DECODER
        x4 = np.maximum(acc3 * inv3 + b3w4[None, :, 0], 0).astype(np.float32)
        s4 = (x4.astype(np.float64) @ b3w4[:, 1].astype(np.float64)).astype(np.float32)
        outs.append(np.tanh(s4 + b4).astype(np.float32))
    return outs
