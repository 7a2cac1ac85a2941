"""Minimal stand-in for the two trimesh features the hot path uses.

The reference builds ``trimesh.Trimesh(vertices, faces, process=False)``, calls
``trimesh.graph.split`` and ``mesh.export(path)`` (utils/mesh.py:371-397).  trimesh
is not installable here, so this module provides the same surface; when trimesh
*is* importable, ``to_trimesh`` hands back a real ``trimesh.Trimesh``.

Host-side post-processing only (the reference does this on the CPU as well);
moving it to the GPU is a "next" row.
"""
from __future__ import annotations

import numpy as np


class Mesh:
    def __init__(self, vertices, faces, process=False):
        self.vertices = np.asarray(vertices)
        self.faces = np.asarray(faces).reshape(-1, 3)

    @property
    def area_faces(self):
        v = self.vertices.astype(np.float64)
        a, b, c = (v[self.faces[:, i]] for i in range(3))
        return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)

    @property
    def area(self):
        return float(self.area_faces.sum())

    @property
    def is_watertight(self):
        if len(self.faces) < 4:
            return False
        e = np.sort(np.concatenate([self.faces[:, [0, 1]], self.faces[:, [1, 2]], self.faces[:, [2, 0]]]), 1)
        _, counts = np.unique(e, axis=0, return_counts=True)
        return bool(np.all(counts == 2))

    def export(self, path):
        export_ply(path, self.vertices, self.faces)

    def to_trimesh(self):
        try:
            import trimesh
        except ImportError:
            return self
        return trimesh.Trimesh(vertices=self.vertices, faces=self.faces, process=False)


def export_ply(path, vertices, faces):
    """Binary little-endian PLY with float32 vertices and ``list uchar int`` faces (the layout
    trimesh's exporter and the legacy plyfile writer of deep_sdf/mesh.py:95-112 both produce)."""
    v = np.ascontiguousarray(vertices, dtype="<f4").reshape(-1, 3)
    f = np.ascontiguousarray(faces, dtype="<i4").reshape(-1, 3)
    header = ("ply\nformat binary_little_endian 1.0\n"
              f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n")
    rec = np.empty(len(f), dtype=[("n", "u1"), ("idx", "<i4", (3,))])
    rec["n"] = 3
    rec["idx"] = f
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(v.tobytes())
        fh.write(rec.tobytes())


def export_ply_records(path, vertices, face_records):
    """Same file as export_ply, from face records already serialised on the GPU
    (engine.ply_face_records: uint8 [F,13] = count byte 3 + three little-endian int32)."""
    v = np.ascontiguousarray(vertices, dtype="<f4").reshape(-1, 3)
    rec = np.ascontiguousarray(face_records, dtype=np.uint8).reshape(-1, 13)
    header = ("ply\nformat binary_little_endian 1.0\n"
              f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {len(rec)}\nproperty list uchar int vertex_indices\nend_header\n")
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(v)                       # contiguous arrays go out through the buffer protocol, no copy
        fh.write(rec)


def load(path, process=False):
    """``trimesh.load(path, process=False)`` for the two formats the path meets: Wavefront OBJ (ground-truth meshes,
    utils/mesh.py:390) and the binary little-endian PLY written by ``export_ply``.  Polygons are fan-triangulated."""
    path = str(path)
    if path.lower().endswith(".obj"):
        verts, faces = [], []
        with open(path, "r") as fh:
            for line in fh:
                if line.startswith("v "):
                    verts.append([float(x) for x in line.split()[1:4]])
                elif line.startswith("f "):
                    idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                    idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                    for k in range(1, len(idx) - 1):
                        faces.append([idx[0], idx[k], idx[k + 1]])
        return Mesh(np.asarray(verts, np.float64).reshape(-1, 3), np.asarray(faces, np.int64).reshape(-1, 3))
    if path.lower().endswith(".ply"):
        with open(path, "rb") as fh:
            raw = fh.read()
        end = raw.index(b"end_header\n") + len(b"end_header\n")
        header = raw[:end].decode("ascii").split("\n")
        if "format binary_little_endian 1.0" not in header:
            raise ValueError(f"{path}: only binary little-endian PLY is supported")
        nv = next(int(l.split()[-1]) for l in header if l.startswith("element vertex"))
        nf = next(int(l.split()[-1]) for l in header if l.startswith("element face"))
        props = [l for l in header if l.startswith("property") and "list" not in l]
        if len(props) != 3:
            raise ValueError(f"{path}: expected x, y, z float vertex properties only")
        v = np.frombuffer(raw, "<f4", nv * 3, end).reshape(nv, 3)
        rec = np.frombuffer(raw, np.dtype([("n", "u1"), ("idx", "<i4", (3,))]), nf, end + nv * 12)
        return Mesh(v.astype(np.float64), rec["idx"].astype(np.int64))
    raise ValueError(f"unsupported mesh format: {path}")


def sample_surface(mesh: Mesh, count, rng=None):
    """``trimesh.sample.sample_surface``: ``count`` points uniformly distributed over the surface (faces picked with
    probability proportional to their area, uniform barycentric coordinates).  -> (points [count,3] f64, face index).
    ``rng``: a numpy Generator (the reference draws from numpy's global state, so its samples are not
    reproducible; parity of what follows is defined for given samples)."""
    rng = np.random.default_rng() if rng is None else rng
    v = np.asarray(mesh.vertices, np.float64)
    f = np.asarray(mesh.faces)
    area = mesh.area_faces
    cum = np.cumsum(area)
    pick = rng.random(count) * cum[-1]
    face_index = np.minimum(np.searchsorted(cum, pick), len(f) - 1)
    origin = v[f[face_index, 0]]
    e1, e2 = v[f[face_index, 1]] - origin, v[f[face_index, 2]] - origin
    r = rng.random((count, 2))
    flip = r.sum(1) > 1.0
    r[flip] = 1.0 - r[flip]
    return origin + e1 * r[:, 0:1] + e2 * r[:, 1:2], face_index


def split(mesh: Mesh, only_watertight=True):
    """Connected components over faces sharing an edge, as sub-meshes (trimesh.graph.split)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    faces = mesh.faces
    F = len(faces)
    if F == 0:
        return []
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), 1).astype(np.int64)
    fid = np.tile(np.arange(F), 3)
    key = e[:, 0] * (int(faces.max()) + 1) + e[:, 1]
    order = np.argsort(key, kind="stable")
    key, fid = key[order], fid[order]
    uniq, start, counts = np.unique(key, return_index=True, return_counts=True)
    two = start[counts == 2]                       # edges shared by exactly two faces
    adj = coo_matrix((np.ones(len(two)), (fid[two], fid[two + 1])), shape=(F, F))
    n, labels = connected_components(adj, directed=False)
    edge_count_of = counts[np.searchsorted(uniq, key)]
    out = []
    for c in range(n):
        sel = np.nonzero(labels == c)[0]
        if only_watertight:
            if len(sel) < 4 or not np.all(edge_count_of[labels[fid] == c] == 2):
                continue
        sub = faces[sel]
        used, inv = np.unique(sub, return_inverse=True)
        out.append(Mesh(mesh.vertices[used], inv.reshape(-1, 3).astype(faces.dtype)))
    return out


def largest_watertight_component_mc(points, faces, verts_local, dims, spacing):
    """Fast path of utils/mesh.py:371-381 for meshes produced by the marching-cubes kernel.

    Same outcome as ``split`` + "keep the max-area piece iff more than one piece" but O(V+F):
    a marching-cubes mesh is a closed manifold except where it leaves the volume, so a component
    is watertight iff none of its triangle edges lies in a boundary plane of the volume (both end
    points on the same plane; ``verts_local`` are the raw grid-local vertices, ``dims`` the volume
    shape, ``spacing`` the voxel size -- lattice planes are hit exactly in float32).
    Returns a Mesh (the input mesh when there are fewer than two watertight pieces)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    faces = np.asarray(faces).reshape(-1, 3)
    V, F = len(points), len(faces)
    whole = Mesh(points, faces)
    if F == 0:
        return whole
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    g = coo_matrix((np.ones(2 * F, np.int8), (np.concatenate([a, b]), np.concatenate([b, c]))), shape=(V, V))
    n, vlabel = connected_components(g, directed=False)
    if n <= 1:
        return whole
    flabel = vlabel[a]
    vl = np.asarray(verts_local, np.float32)
    on = np.zeros(V, np.uint8)                     # bit 2k: on plane index 0 of axis k, bit 2k+1: on the last plane
    for k in range(3):
        on |= (vl[:, k] == 0.0).astype(np.uint8) << (2 * k)
        on |= (vl[:, k] == np.float32(float(dims[k] - 1) * float(spacing[k]))).astype(np.uint8) << (2 * k + 1)
    fa, fb, fc = on[a], on[b], on[c]
    open_face = ((fa & fb) | (fb & fc) | (fc & fa)) != 0
    is_open = np.bincount(flabel, weights=open_face, minlength=n) > 0
    nfaces = np.bincount(flabel, minlength=n)
    cand = np.nonzero(~is_open & (nfaces >= 4))[0]
    if len(cand) <= 1:
        return whole
    v = np.asarray(points, np.float64)
    e1, e2 = v[b] - v[a], v[c] - v[a]
    cx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    cy = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    cz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    area = 0.5 * np.sqrt(cx * cx + cy * cy + cz * cz)
    comp_area = np.bincount(flabel, weights=area, minlength=n)
    # trimesh orders the pieces by their first face; the reference keeps the first maximum
    first_face = np.full(n, F, np.int64)
    ul, ui = np.unique(flabel, return_index=True)
    first_face[ul] = ui
    cand = cand[np.argsort(first_face[cand], kind="stable")]
    best = cand[int(np.argmax(comp_area[cand]))]
    sel = faces[flabel == best]
    used = np.zeros(V, bool)
    used[sel.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return Mesh(np.asarray(points)[used], remap[sel].astype(faces.dtype))
