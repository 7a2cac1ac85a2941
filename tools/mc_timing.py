"""Marching cubes timing on a decoder field (CUDA events around count / emit).  python tools/mc_timing.py [N]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import _lib, engine, mesh as amesh, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
dec = synthetic.make_decoder(0, init=os.environ.get("MC_DECODER", "engineered"))
s = synthetic.make_sample(0).to(dev)
vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N)
vol = vols[os.environ.get("MC_SURFACE", "hand")].contiguous()
vs = float(vols["voxel"])
L = _lib.lib()
p = _lib.McParams()
p.n0 = p.n1 = p.n2 = N
p.full1 = p.full2 = N
p.index0_offset, p.iso = 0, 0.0
for k in range(3):
    p.spacing[k], p.origin[k] = vs, float(vols["origin"][k])
st = _lib.stream_ptr(dev)
scratch = torch.empty(L.asdf_mc_scratch_bytes(C.byref(p)), dtype=torch.uint8, device=dev)
totals = torch.empty(5, dtype=torch.int64, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tc, te = [], []
for it in range(8):
    ev[0].record()
    _lib.check(L.asdf_mc_count(_lib.ptr(vol), C.byref(p), _lib.ptr(scratch), _lib.ptr(totals), st), "count")
    ev[1].record()
    nv, nt, nseg = int(totals[0]), int(totals[1]), int(totals[4])
    verts = torch.empty((nv, 3), device=dev); pts = torch.empty((nv, 3), device=dev)
    faces = torch.empty((nt, 3), dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.asdf_mc_emit(_lib.ptr(vol), C.byref(p), _lib.ptr(scratch), nseg, _lib.ptr(verts), _lib.ptr(pts),
                              _lib.ptr(faces), None, st), "emit")
    e1.record()
    torch.cuda.synchronize()
    if it >= 3:
        tc.append(ev[0].elapsed_time(ev[1])); te.append(e0.elapsed_time(e1))
bytes_alg = 4 * N ** 3 + 12 * nv + 12 * nt
t = (sum(tc) + sum(te)) / len(tc)
print(f"N={N} V={nv} F={nt}: count {sum(tc) / len(tc) * 1e3:.0f} us, emit {sum(te) / len(te) * 1e3:.0f} us, "
      f"algorithmic {bytes_alg / 1e6:.1f} MB -> {bytes_alg / t / 1e6:.0f} GB/s")
