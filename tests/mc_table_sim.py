"""CPU simulation of what csrc/mc.cu does with the generated tables (test helper).

Mirrors the kernel's arithmetic step by step (config bits, decider products, variant index,
table walk, vertex keys) so the tables and the index arithmetic can be checked against the
independent oracle tracing without a GPU."""
import numpy as np

from alignsdf_b200 import mc_tables as T


def table_mc(vol, iso=0.0):
    t = T.build_tables()
    vol = np.ascontiguousarray(vol, np.float32)
    n0, n1, n2 = vol.shape
    f = vol - np.float32(iso)
    ins = f < 0
    keys = []
    for a in range(3):
        lo = [slice(None)] * 3; hi = [slice(None)] * 3
        lo[a] = slice(0, vol.shape[a] - 1); hi[a] = slice(1, vol.shape[a])
        idx = np.argwhere(ins[tuple(lo)] != ins[tuple(hi)])
        keys.append(((idx[:, 0] * n1 + idx[:, 1]) * n2 + idx[:, 2]) * 4 + a)
    faces = []
    for i in range(n0 - 1):
        for j in range(n1 - 1):
            for k in range(n2 - 1):
                v = [f[i + ((c >> 2) & 1), j + ((c >> 1) & 1), k + (c & 1)] for c in range(8)]
                config = sum((1 << c) for c in range(8) if v[c] < 0)
                if config in (0, 255):
                    continue
                amb = int(t["amb_mask"][config])
                variant, bit = 0, 0
                for face in range(6):
                    if amb >> face & 1:
                        q = t["face_corners"][face]
                        p02 = np.float32(v[q[0]]) * np.float32(v[q[2]])
                        p13 = np.float32(v[q[1]]) * np.float32(v[q[3]])
                        joined = (p02 > p13) if v[q[0]] < 0 else (p13 > p02)
                        variant |= int(joined) << bit
                        bit += 1
                e = int(t["var_offset"][config]) + variant
                if t["has_center"][e]:
                    keys.append(np.array([((i * n1 + j) * n2 + k) * 4 + 3]))
                for tt in range(int(t["n_tris"][e])):
                    tri = t["tri_edges"][int(t["tri_start"][e]) + tt]
                    ks = []
                    for ed in tri:
                        if ed == T.CENTER:
                            ks.append(((i * n1 + j) * n2 + k) * 4 + 3)
                            continue
                        c0 = int(t["edge_corner"][ed][0])
                        a = int(ed) // 4
                        p = ((i + ((c0 >> 2) & 1)) * n1 + (j + ((c0 >> 1) & 1))) * n2 + (k + (c0 & 1))
                        ks.append(p * 4 + a)
                    faces.append(ks)
    keys = np.sort(np.concatenate(keys)).astype(np.uint64)
    faces = np.asarray(faces, np.uint64).reshape(-1, 3)
    return keys, np.searchsorted(keys, faces).astype(np.int32)
