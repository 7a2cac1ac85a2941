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
    utils/mesh.py:390) and PLY (``_load_ply``: our own files, trimesh's, plyfile's; ascii or binary).  Polygons are
    fan-triangulated."""
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
        return _load_ply(path)
    raise ValueError(f"unsupported mesh format: {path}")


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def _load_ply(path):
    """PLY reader for what the path meets: files written by ``export_ply`` / trimesh / plyfile (binary, float x y z +
    ``list uchar int vertex_indices``) and, more generally, ascii or binary (either byte order) files whose vertex
    element carries x, y, z among any scalar properties (normals, colours ... are skipped) and whose face element
    starts with the index list.  Polygons are fan-triangulated; elements after ``face`` are ignored."""
    with open(path, "rb") as fh:
        raw = fh.read()
    try:
        end = raw.index(b"end_header")
        end = raw.index(b"\n", end) + 1
    except ValueError:
        raise ValueError(f"{path}: not a PLY file (no end_header)") from None
    lines = [l.strip() for l in raw[:end].decode("ascii", "replace").splitlines()]
    if not lines or lines[0] != "ply":
        raise ValueError(f"{path}: not a PLY file")
    fmt = next((l.split()[1] for l in lines if l.startswith("format ")), None)
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"{path}: unknown PLY format {fmt!r}")
    elements = []                                  # [name, count, [(kind, ...)]]
    for l in lines:
        tok = l.split()
        if not tok:
            continue
        if tok[0] == "element":
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if not elements:
                raise ValueError(f"{path}: property before any element")
            if tok[1] == "list":
                elements[-1][2].append(("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]], tok[4]))
            else:
                elements[-1][2].append(("scalar", _PLY_TYPES[tok[1]], tok[2]))
    names = [e[0] for e in elements]
    if "vertex" not in names:
        raise ValueError(f"{path}: no vertex element")
    verts = np.zeros((0, 3), np.float64)
    faces = np.zeros((0, 3), np.int64)
    if fmt == "ascii":
        toks = raw[end:].split()
        pos = 0
        for name, count, props in elements:
            if any(p[0] == "list" for p in props):
                rows = []
                for _ in range(count):
                    row = []
                    for p in props:
                        if p[0] == "list":
                            k = int(toks[pos])
                            row.append([int(t) for t in toks[pos + 1:pos + 1 + k]])
                            pos += 1 + k
                        else:
                            row.append(None)
                            pos += 1
                    rows.append(row)
                if name == "face":
                    li = next(i for i, p in enumerate(props) if p[0] == "list")
                    faces = _fan([r[li] for r in rows])
            else:
                block = np.asarray(toks[pos:pos + count * len(props)], dtype=np.float64).reshape(count, len(props))
                pos += count * len(props)
                if name == "vertex":
                    verts = _xyz(path, props, lambda j: block[:, j])
            if name == "face":
                break
        return Mesh(verts, faces)
    bo = "<" if fmt == "binary_little_endian" else ">"
    pos = end
    for name, count, props in elements:
        if all(p[0] == "scalar" for p in props):
            dt = np.dtype([(p[2] + "_%d" % j, bo + p[1]) for j, p in enumerate(props)])
            block = np.frombuffer(raw, dt, count, pos)
            pos += dt.itemsize * count
            if name == "vertex":
                fields = list(dt.names)
                verts = _xyz(path, props, lambda j: block[fields[j]])
        else:
            if name != "face":
                if "face" in names[names.index(name):]:
                    raise ValueError(f"{path}: list properties before the face element are not supported")
                break
            if props[0][0] != "list":
                raise ValueError(f"{path}: the face element must start with its index list")
            _, ct, it, _ = props[0]
            rest = [p for p in props[1:]]
            if any(p[0] == "list" for p in rest):
                raise ValueError(f"{path}: more than one list property per face is not supported")
            tail = sum(np.dtype(p[1]).itemsize for p in rest)
            csz, isz = np.dtype(ct).itemsize, np.dtype(it).itemsize
            rec3 = np.dtype([("n", bo + ct), ("idx", bo + it, (3,)), ("tail", "u1", (tail,))])
            fast = np.frombuffer(raw, rec3, count, pos) if pos + rec3.itemsize * count <= len(raw) else None
            if fast is not None and (count == 0 or bool((fast["n"] == 3).all())):
                faces = fast["idx"].astype(np.int64).reshape(-1, 3)
            else:                                  # polygons of varying size: walk the records
                polys = []
                for _ in range(count):
                    k = int(np.frombuffer(raw, bo + ct, 1, pos)[0])
                    polys.append(np.frombuffer(raw, bo + it, k, pos + csz).astype(np.int64).tolist())
                    pos += csz + k * isz + tail
                faces = _fan(polys)
            break
    return Mesh(verts, faces)


def _xyz(path, props, column):
    cols = {p[2]: j for j, p in enumerate(props)}
    if not all(k in cols for k in "xyz"):
        raise ValueError(f"{path}: the vertex element has no x / y / z properties")
    return np.stack([np.asarray(column(cols[k]), np.float64) for k in "xyz"], 1)


def _fan(polys):
    out = [[p[0], p[k], p[k + 1]] for p in polys for k in range(1, len(p) - 1)]
    return np.asarray(out, np.int64).reshape(-1, 3)


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
