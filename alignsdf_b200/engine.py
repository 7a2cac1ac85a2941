"""Host-side driver of the CUDA library: packs a decoder once, binds a sample
(latent + poses) and launches grid / point evaluations and marching cubes.

Everything numeric happens in libalignsdf_b200.so; this module only folds
weights (packer.py), owns device buffers (torch tensors) and marshals arguments.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import _lib, packer
from ._lib import AsdfError

INT_MAX = 2 ** 31 - 1
# which tensor-core kernel "auto" prefers: "tc3" = k1_tc3.cu (fp16 + fp8 corrections, falls back to
# "tc2" = k1_tc2.cu (fp16 x3) when an activation leaves the fp8 operand range), "tc" = k1_tc.cu
DEFAULT_TC_PATH = "tc3"
FALLBACKS = {"tc3_to_tc2": 0}    # launches whose fp8 range flag fired and were re-run through k1_tc2.cu
LAUNCHES = {"count": 0}          # kernels of libalignsdf_b200.so launched so far (bench.py reports it)
_GRID_MODES = {"reference": _lib.QUERY_GRID_REFERENCE, "regular": _lib.QUERY_GRID_REGULAR}


def _device_of(latent, device=None):
    if device is not None and str(device) not in ("cpu",):
        d = torch.device(device)
        if d.type == "cuda":
            return d
    if isinstance(latent, torch.Tensor) and latent.is_cuda:
        return latent.device
    if not torch.cuda.is_available():
        raise AsdfError("alignsdf_b200 needs a CUDA device (B200); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class BoundSample:
    """A decoder with one sample's latent/pose folded in, ready to be queried."""

    def __init__(self, engine, branches, feature_mode, nerf_freqs=0):
        self.engine = engine
        self.feature_mode = feature_mode
        self.nerf_freqs = int(nerf_freqs)       # > 0: xyz queries, NeRF-encoded inside the generic kernel
        self.device = engine.device
        self._simt_ready = False
        self._two_outputs = len(branches) == 2 or int(branches[0].layers[-1].B.shape[0]) == 2
        self.tc = None
        self.tc2 = None
        self.tc3 = None
        self._branches = branches
        if not feature_mode and not self.nerf_freqs and engine.tc_supported:
            from . import tc_pack
            self.tc = tc_pack.bind(engine, branches)

    def _ensure_simt(self):
        """Buffers + descriptor of the generic fp32 kernel, built on first use (the tensor-core paths
        never need them)."""
        if self._simt_ready:
            return
        engine, topo, dev = self.engine, self.engine.topo, self.device
        pack = packer.pack_simt(self._branches)
        self.simt_pack = pack
        if engine.simt_static is None:      # static weights do not depend on the sample
            engine.simt_static = torch.from_numpy(pack.static).to(dev)
        self.simt_sample = torch.from_numpy(pack.sample).to(dev, non_blocking=True)
        d = _lib.SimtDesc()
        d.n_branches, d.n_layers, d.n_outputs = pack.n_branches, pack.n_layers, pack.n_outputs
        d.pre_tanh = int(topo.pre_tanh)
        d.n_class = 0 if topo.classifier is None else int(topo.classifier[0].shape[0])
        d.nerf_freqs = self.nerf_freqs
        for b in range(pack.n_branches):
            d.point_dim[b] = int(pack.point_dim[b])
            idx = (packer.branch_feature_index(topo, topo.branches[b][0]) if self.feature_mode and not self.nerf_freqs
                   else np.arange(3))
            for k, v in enumerate(idx):
                d.point_index[b][k] = int(v)
            for l in range(pack.n_layers):
                for k in range(8):
                    d.table[b][l][k] = int(pack.table[b, l, k])
        self.simt_desc = d
        self._simt_ready = True

    def _tc2_for(self, p_absmax: float):
        """Per-sample block of the v2 tensor-core kernel, valid for |xyz| <= p_absmax (the point
        operand scale is baked into it); returns None when the fp16 ranges cannot be met."""
        if self.feature_mode or self.nerf_freqs or not self.engine.tc_supported:
            return None
        need = max(2.0, float(p_absmax) * 1.01)
        if self.tc2 is None or self.tc2.info["p_absmax"] < need:
            from . import tc2_pack
            try:
                self.tc2 = tc2_pack.bind(self.engine, self._branches, need)
            except ValueError:
                return None
        return self.tc2

    def _tc3_for(self, p_absmax: float):
        """Same for the v3 kernel (its own block: activations are not pre-scaled there)."""
        if self.feature_mode or self.nerf_freqs or not self.engine.tc_supported:
            return None
        need = max(2.0, float(p_absmax) * 1.01)
        if self.tc3 is None or self.tc3.info["p_absmax"] < need:
            from . import tc3_pack
            try:
                self.tc3 = tc3_pack.bind(self.engine, self._branches, need)
            except ValueError:
                return None
        return self.tc3

    # ------------------------------------------------------------------
    def _run(self, q: _lib.Query, n: int, want_cls: bool, bbox: bool, path: str, p_absmax: float = 2.0):
        dev = self.device
        hand = torch.empty(n, dtype=torch.float32, device=dev)
        two = self._two_outputs
        obj = torch.empty(n, dtype=torch.float32, device=dev) if two else None
        cls = torch.empty(n, dtype=torch.int32, device=dev) if want_cls else None
        box = None
        if bbox:
            box = torch.tensor([INT_MAX] * 3 + [-1] * 3 + [INT_MAX] * 3 + [-1] * 3,
                               dtype=torch.int32, device=dev)
        if n == 0:                      # empty query: nothing to launch
            return hand, obj, cls, box
        L = _lib.lib()
        # kernel choice: tensor-core kernels need the shipped topology and no class output
        want = DEFAULT_TC_PATH if path == "auto" else path
        if want == "tc3" and path == "auto" and not self.engine.tc3_in_range:
            want = "tc2"                # this decoder already left the fp8 operand range once: do not try again
        tc2 = tc3 = None
        if want == "tc3":
            tc3 = None if want_cls else self._tc3_for(p_absmax)
            if tc3 is None:
                if path == "tc3":
                    raise AsdfError("tensor-core (v3) path requested but not available for this decoder/query")
                want = "tc2"
        if want == "tc2":
            tc2 = None if want_cls else self._tc2_for(p_absmax)
            if tc2 is None:
                if path == "tc2":
                    raise AsdfError("tensor-core (v2) path requested but not available for this decoder/query")
                want = "tc"
        if want == "tc" and (self.tc is None or want_cls):
            if path == "tc":
                raise AsdfError("tensor-core path requested but not available for this decoder/query")
            want = "simt"
        use_tc3, use_tc2, use_tc = want == "tc3", want == "tc2", want == "tc"
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            if use_tc3:
                status = torch.zeros(1, dtype=torch.int32, device=dev)
                rc = L.asdf_tc3_eval(_lib.ptr(self.engine.tc3_static), _lib.ptr(tc3.sample), C.byref(q),
                                     _lib.ptr(hand), _lib.ptr(obj), _lib.ptr(box), _lib.ptr(status), st)
                _lib.check(rc, "asdf_tc3_eval")
                LAUNCHES["count"] += 1
                if int(status.item()) != 0:
                    # an activation left the range of the fp8 correction operands: the launch's outputs
                    # (and bbox) are not trustworthy -> same query through the all-fp16 kernel
                    if path == "tc3" and os.environ.get("ALIGNSDF_B200_STRICT_PATH"):
                        raise AsdfError("k1_tc3: activation outside the fp8 operand range")
                    FALLBACKS["tc3_to_tc2"] += 1
                    self.engine.tc3_in_range = False
                    return self._run(q, n, want_cls, bbox, "tc2", p_absmax)
            elif use_tc2:
                rc = L.asdf_tc2_eval(_lib.ptr(self.engine.tc2_static), _lib.ptr(tc2.sample), C.byref(q),
                                     _lib.ptr(hand), _lib.ptr(obj), _lib.ptr(box), st)
                _lib.check(rc, "asdf_tc2_eval")
                LAUNCHES["count"] += 1
            elif use_tc:
                rc = L.asdf_tc_eval(C.byref(self.tc.desc), _lib.ptr(self.engine.tc_static),
                                    _lib.ptr(self.tc.sample), C.byref(q), _lib.ptr(hand),
                                    _lib.ptr(obj), _lib.ptr(box), st)
                _lib.check(rc, "asdf_tc_eval")
                LAUNCHES["count"] += 1
            else:
                self._ensure_simt()
                rc = L.asdf_simt_eval(C.byref(self.simt_desc), _lib.ptr(self.engine.simt_static),
                                      _lib.ptr(self.simt_sample), _lib.ptr(self.engine.cls_dev),
                                      C.byref(q), _lib.ptr(hand), _lib.ptr(obj), _lib.ptr(cls),
                                      _lib.ptr(box), st)
                _lib.check(rc, "asdf_simt_eval")
                LAUNCHES["count"] += 1
        return hand, obj, cls, box

    def eval_grid(self, N, voxel, origin, mode="reference", begin=0, end=None, bbox_mask=0,
                  want_cls=False, path=None):
        """Evaluate linear grid indices [begin,end) -> (hand, obj, cls, bbox int32[12] or None)."""
        if self.feature_mode and not self.nerf_freqs:
            raise AsdfError("grid evaluation needs an xyz-folded sample (feature_mode=False)")
        end = N ** 3 if end is None else end
        q = _lib.Query()
        q.mode, q.N, q.begin, q.end = _GRID_MODES[mode], int(N), int(begin), int(end)
        q.voxel = float(voxel)
        for k in range(3):
            q.origin[k] = float(origin[k])
        q.points_dev, q.point_stride, q.bbox_mask = None, 0, int(bbox_mask)
        # bound on |xyz| over the (possibly sheared: up to one extra voxel) lattice
        pmax = max(max(abs(float(origin[k])), abs(float(origin[k]) + (N + 1) * float(voxel))) for k in range(3))
        return self._run(q, end - begin, want_cls, bbox_mask != 0, path or self.engine.path, pmax)

    def eval_points(self, points: torch.Tensor, want_cls=False, path=None):
        """points: CUDA f32 [P, stride]; xyz rows, or embedded feature rows in feature mode."""
        _lib.require_cuda(points, "points")
        pts = points.to(torch.float32).contiguous()
        q = _lib.Query()
        q.mode, q.N, q.begin, q.end = _lib.QUERY_POINTS, 0, 0, int(pts.shape[0])
        q.voxel = 0.0
        q.points_dev, q.point_stride, q.bbox_mask = pts.data_ptr(), int(pts.shape[1]), 0
        pth = path or self.engine.path
        pmax = 2.0
        if pts.shape[0] and not self.feature_mode and pth in ("auto", "tc2", "tc3") and self.engine.tc_supported:
            pmax = float(pts[:, :3].abs().max())
        hand, obj, cls, _ = self._run(q, pts.shape[0], want_cls, False, pth, pmax)
        return hand, obj, cls


class DecoderEngine:
    """Per-decoder state: topology, static weights resident in HBM."""

    def __init__(self, decoder, device):
        self.device = torch.device(device)
        self.topo = packer.decoder_topology(decoder)
        self.simt_static = None
        self.cls_dev = None
        if self.topo.classifier is not None:
            Wc, bc = self.topo.classifier
            self.cls_dev = torch.from_numpy(
                np.concatenate([Wc, bc[:, None]], 1).astype(np.float32)).to(self.device)
        self.path = os.environ.get("ALIGNSDF_B200_PATH", "auto")
        self.tc_static = None
        self.tc2_static = None
        self.tc3_static = None
        self.tc2_scales = None
        self.tc3_scales = None
        self.tc3_in_range = True        # cleared when a k1_tc3 launch reports an activation >= 448 (sticky)
        self.tc_supported = False
        try:
            from . import tc_pack
            self.tc_supported = tc_pack.supported(self.topo) and _lib.lib().asdf_tc_static_bytes() > 0
            if self.tc_supported:
                self.tc_static = tc_pack.pack_static(self)
                from . import tc2_pack
                self.tc2_static = tc2_pack.pack_static(self)
                from . import tc3_pack
                self.tc3_static = tc3_pack.pack_static(self)
        except ImportError:
            self.tc_supported = False

    def bind(self, latent, specs, mano_results, obj_results, feature_mode=False) -> BoundSample:
        # NeRF positional encoding (utils/mesh.py:54-55) is not affine in xyz: the weights are folded as
        # for feature queries and the generic kernel encodes xyz itself
        nerf = 0 if feature_mode else packer.nerf_freqs(specs, mano_results)
        branches = packer.fold_decoder(self.topo, latent, specs, mano_results, obj_results,
                                       feature_mode or nerf > 0)
        return BoundSample(self, branches, feature_mode or nerf > 0, nerf)


def unwrap_decoder(decoder):
    """Legacy API: accept a thin adapter around a decoder (e.g. one that returns only the first
    output); the weights are taken from the first sub-module that owns ``lin*`` layers."""
    for m in decoder.modules():
        if any(n.startswith("lin") for n, _ in m.named_children()):
            return m
    return decoder


_ENGINES: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _version_of(decoder):
    return tuple(int(p._version) for p in decoder.parameters())


def get_engine(decoder, device) -> DecoderEngine:
    """Cached engine per (decoder object, device); rebuilt if any parameter was modified."""
    device = torch.device(device)
    per = _ENGINES.setdefault(decoder, {})
    ver = _version_of(decoder)
    hit = per.get(device)
    if hit is None or hit[0] != ver:
        hit = (ver, DecoderEngine(decoder, device))
        per[device] = hit
    return hit[1]


# ----------------------------------------------------------------------------
# marching cubes
# ----------------------------------------------------------------------------
def _decode_ordered(bits: int) -> float:
    b = bits if bits >= 0 else bits ^ 0x7FFFFFFF
    return float(np.array([b], np.int32).view(np.float32)[0])


def marching_cubes(vol: torch.Tensor, level=0.0, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0),
                   index0_offset=0, want_keys=False, check_range=True):
    """GPU marching cubes.  vol: CUDA f32 [n0,n1,n2].  Returns dict of CUDA tensors
    verts [V,3] (array-axis order x spacing), points [V,3] (= origin + verts), faces [F,3] int32,
    keys [V] int64 (when want_keys).  Raises ValueError like skimage when level is outside the
    data range (the reference catches it, utils/mesh.py:353-358); ``check_range=False`` (z-slabs,
    where an empty slab is normal) returns empty tensors instead."""
    _lib.require_cuda(vol, "vol")
    vol = vol.to(torch.float32).contiguous()
    if vol.dim() != 3:
        raise ValueError("Input volume should be a 3D numpy array.")
    p = _lib.McParams()
    p.n0, p.n1, p.n2 = (int(s) for s in vol.shape)
    p.full1, p.full2 = p.n1, p.n2
    p.index0_offset = int(index0_offset)
    p.iso = float(level)
    for k in range(3):
        p.spacing[k] = float(spacing[k])
        p.origin[k] = float(origin[k])
    L = _lib.lib()
    dev = vol.device
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        scratch = torch.empty(L.asdf_mc_scratch_bytes(C.byref(p)), dtype=torch.uint8, device=dev)
        totals = torch.empty(4, dtype=torch.int64, device=dev)
        _lib.check(L.asdf_mc_count(_lib.ptr(vol), C.byref(p), _lib.ptr(scratch), _lib.ptr(totals), st),
                   "asdf_mc_count")
        LAUNCHES["count"] += 5
        nv, nt, mn, mx = (int(x) for x in totals.cpu())
        if check_range and not (_decode_ordered(mn) <= float(np.float32(level)) <= _decode_ordered(mx)):
            raise ValueError("Surface level must be within volume data range.")
        verts = torch.empty((nv, 3), dtype=torch.float32, device=dev)
        points = torch.empty((nv, 3), dtype=torch.float32, device=dev)
        faces = torch.empty((nt, 3), dtype=torch.int32, device=dev)
        keys = torch.empty(nv, dtype=torch.int64, device=dev) if want_keys else None
        if nv > 0 or nt > 0:
            _lib.check(L.asdf_mc_emit(_lib.ptr(vol), C.byref(p), _lib.ptr(scratch), _lib.ptr(verts),
                                      _lib.ptr(points), _lib.ptr(faces), _lib.ptr(keys), st),
                       "asdf_mc_emit")
            LAUNCHES["count"] += 1
    return dict(verts=verts, points=points, faces=faces, keys=keys)


def ply_face_records(faces: torch.Tensor) -> torch.Tensor:
    """Binary-PLY face records built on the device: uint8 [F,13] = list length 3 + the three int32
    indices (little endian), ready to be written after the header and the vertex block."""
    F = int(faces.shape[0])
    rec = torch.empty((F, 13), dtype=torch.uint8, device=faces.device)
    if F:
        rec[:, 0] = 3
        rec[:, 1:] = faces.contiguous().view(torch.uint8).view(F, 12)
    return rec


def select_component(points: torch.Tensor, faces: torch.Tensor, verts_local: torch.Tensor, dims, spacing):
    """GPU version of trimesh_lite.largest_watertight_component_mc (utils/mesh.py:371-381): connected
    components of a marching-cubes mesh, and -- iff at least two of them are watertight -- the
    largest-area watertight one (first maximum in order of first face), compacted with vertex and
    face order preserved.  points / verts_local: CUDA f32 [V,3], faces: CUDA int32 [F,3].
    Returns (points, faces, info); the inputs themselves when the mesh is kept whole."""
    _lib.require_cuda(points, "points")
    V, F = int(points.shape[0]), int(faces.shape[0])
    info = dict(components=0, watertight=0, kept="whole")
    if F == 0 or V == 0:
        return points, faces, info
    L = _lib.lib()
    dev = points.device
    points = points.contiguous()
    faces = faces.contiguous()
    verts_local = verts_local.contiguous()
    last = (C.c_float * 3)(*[float(np.float32(float(dims[k] - 1) * float(spacing[k]))) for k in range(3)])
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        parent = torch.empty(V, dtype=torch.int32, device=dev)
        _lib.check(L.asdf_cc_label(_lib.ptr(faces), F, V, _lib.ptr(parent), st), "asdf_cc_label")
        area = torch.zeros(V, dtype=torch.float64, device=dev)
        acc = torch.zeros((2, V), dtype=torch.int32, device=dev)          # nfaces, open
        first = torch.full((V,), INT_MAX, dtype=torch.int32, device=dev)
        _lib.check(L.asdf_cc_stats(_lib.ptr(faces), F, _lib.ptr(verts_local), _lib.ptr(points), V, _lib.ptr(parent),
                                   C.byref(last), _lib.ptr(area), _lib.ptr(acc[0]), _lib.ptr(acc[1]),
                                   _lib.ptr(first), st), "asdf_cc_stats")
        LAUNCHES["count"] += 4
        roots = torch.nonzero(acc[0] > 0).flatten()
        n = int(roots.shape[0])
        info["components"] = n
        if n <= 1:
            return points, faces, info
        small = torch.stack([area[roots], acc[0][roots].double(), acc[1][roots].double(), first[roots].double(),
                             roots.double()]).cpu().numpy()
        comp_area, nfaces, is_open, first_face, label = small
        cand = np.nonzero((is_open == 0) & (nfaces >= 4))[0]
        info["watertight"] = int(len(cand))
        if len(cand) <= 1:
            return points, faces, info
        cand = cand[np.argsort(first_face[cand], kind="stable")]      # trimesh orders the pieces by first face
        best = int(label[cand[int(np.argmax(comp_area[cand]))]])      # the reference keeps the first maximum
        keep = torch.empty(V + F, dtype=torch.int32, device=dev)
        _lib.check(L.asdf_cc_mark(_lib.ptr(parent), V, _lib.ptr(faces), F, best, _lib.ptr(keep[:V]), _lib.ptr(keep[V:]), st),
                   "asdf_cc_mark")
        scan_v = torch.cumsum(keep[:V], 0, dtype=torch.int32)
        scan_f = torch.cumsum(keep[V:], 0, dtype=torch.int32)
        nv, nf = int(scan_v[-1]), int(scan_f[-1])
        out_p = torch.empty((nv, 3), dtype=torch.float32, device=dev)
        out_f = torch.empty((nf, 3), dtype=torch.int32, device=dev)
        _lib.check(L.asdf_cc_gather(_lib.ptr(points), _lib.ptr(faces), V, F, _lib.ptr(keep[:V]), _lib.ptr(scan_v),
                                    _lib.ptr(keep[V:]), _lib.ptr(scan_f), _lib.ptr(out_p), _lib.ptr(out_f), st),
                   "asdf_cc_gather")
        LAUNCHES["count"] += 2
    info["kept"] = best
    return out_p, out_f, info


def grid_points(N, voxel, origin, mode="reference", begin=0, end=None, device=None) -> torch.Tensor:
    """Query coordinates of linear indices [begin,end) as the kernels generate them."""
    dev = _device_of(None, device)
    end = N ** 3 if end is None else end
    q = _lib.Query()
    q.mode, q.N, q.begin, q.end = _GRID_MODES[mode], int(N), int(begin), int(end)
    q.voxel = float(voxel)
    for k in range(3):
        q.origin[k] = float(origin[k])
    out = torch.empty((end - begin, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().asdf_grid_points(C.byref(q), _lib.ptr(out), _lib.stream_ptr(dev)),
                   "asdf_grid_points")
    return out


def nerf_embed(xyz: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """[..., 3] CUDA f32 -> [..., 3 + 6 n_freqs] NeRF positional encoding (utils/utils.py:521-533)."""
    _lib.require_cuda(xyz, "xyz")
    flat = xyz.to(torch.float32).reshape(-1, 3).contiguous()
    out = torch.empty((flat.shape[0], 3 + 6 * int(n_freqs)), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().asdf_nerf_embed(_lib.ptr(flat), flat.shape[0], int(n_freqs), _lib.ptr(out),
                                              _lib.stream_ptr(xyz.device)), "asdf_nerf_embed")
    LAUNCHES["count"] += 1
    return out.reshape(*xyz.shape[:-1], out.shape[-1])


def embed_points(xyz: torch.Tensor, specs, mano_results, obj_results) -> torch.Tensor:
    """features[P,pf] = A.xyz + c on the GPU (public kinematic_embedding API)."""
    _lib.require_cuda(xyz, "xyz")
    A, c = packer.embedding_affine(specs, mano_results, obj_results)
    aff = torch.from_numpy(np.concatenate([A, c[:, None]], 1).astype(np.float32)).to(xyz.device)
    x = xyz.to(torch.float32).contiguous()
    out = torch.empty((x.shape[0], A.shape[0]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().asdf_embed_points(_lib.ptr(x), x.shape[0], _lib.ptr(aff), A.shape[0],
                                                _lib.ptr(out), _lib.stream_ptr(x.device)),
                   "asdf_embed_points")
    return out
