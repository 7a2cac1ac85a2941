"""Throughput of the decoder / embedding variants next to the flagship configuration, and of the ICP / Chamfer step
after the path:
    python tools/variant_timing.py [N] [reps]
For every listed golden configuration: the two grid passes of utils/mesh.py:24-120 (``mesh.sdf_volumes``) at N^3,
CUDA events around ``reps`` runs after one warm-up -> M hand+obj (or single-output) queries per second and the kernel
kinds that ran.  Then: ICP_T_S.run_icp_f + Chamfer on 30 000 + 30 000 points, against the oracle's wall time on the
host (sklearn KD-trees, the reference's own libraries)."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from alignsdf_b200 import engine, mesh as amesh, trimesh_lite as tl  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
DEV = torch.device("cuda")
CASES = ["sep_default_n32", "sep_both9_n24", "comb_both9_n16", "comb_default_n16", "sep_hand51_n16", "sep_both54_n12",
         "comb_cls_n12", "comb_xyzall_n12", "sep_nerf27_n12", "comb_nerf15_n12", "sep_ln_both9_n12", "comb_ln_both9_n12",
         "sep_pa_both9_n12", "comb_pa_xyz3_n10"]


def time_case(name):
    meta, _, dec, sample = helpers.load_case(name)
    s = sample.to(DEV)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    cam = getattr(s, "cam_intr", None)

    def run():
        return amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N, hb, ob, keep_pass1=False,
                                 cam_intr=cam)
    vols = run()                                                  # warm-up (packs the decoder, calibrates)
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        vols = run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / REPS
    outs = int(vols["hand"] is not None) + int(vols["obj"] is not None)
    eng = engine.get_engine(dec, DEV)
    kinds = sorted(vols["bound"].kinds_used) if hasattr(vols["bound"], "kinds_used") else []
    print(f"{name:22s} {ms:9.2f} ms per sample  {2 * N ** 3 / ms / 1e3:8.1f} M grid points/s x {outs} output(s)  "
          f"level {getattr(eng, 'level', '?')} kinds {kinds}", flush=True)


def time_icp(n=30000):
    from alignsdf_b200.deep_sdf.metrics import chamfer as gch
    from alignsdf_b200.deep_sdf.metrics.icp_trans_scale import ICP_T_S
    from oracle import icp_oracle
    g = np.random.default_rng(0)

    def surf(k):
        d = g.normal(size=(k, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = 0.08 * (1.0 + 0.3 * np.sin(5 * d[:, 0]) * np.cos(3 * d[:, 1]) + 0.2 * d[:, 2] ** 2)
        return d * r[:, None] * np.array([1.0, 0.7, 0.5])
    src, tgt = surf(n), surf(n) * 1.2 + [0.02, -0.01, 0.03]
    none = np.zeros((0, 3), np.int64)

    def gpu():
        icp = ICP_T_S(tl.Mesh(src.copy(), none), tl.Mesh(tgt.copy(), none))
        icp.normalize_points()
        icp.run_icp_f(max_iter=100)
        cd = gch.chamfer_points(icp.points_source * icp.scale + icp.trans, icp.points_target)
        return len(icp.errors), cd
    gpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it, cd = gpu()
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    moved, _ = icp_oracle.normalize(src, tgt)
    scale, trans, errors = icp_oracle.run_icp_f(moved, tgt, max_iter=100)
    cd_o = icp_oracle.chamfer(moved * scale + trans, tgt)
    t_cpu = time.perf_counter() - t0
    print(f"ICP (run_icp_f, {n}+{n} points) + Chamfer: GPU {1e3 * t_gpu:.1f} ms ({it} iterations, {1e3 * t_gpu / it:.2f} ms each, "
          f"chamfer {cd:.6g}) | oracle on the host {1e3 * t_cpu:.0f} ms ({len(errors)} iterations, chamfer {cd_o:.6g})", flush=True)
    # one neighbour search alone, device-timed
    from alignsdf_b200.deep_sdf.metrics.icp_trans_scale import nn_search
    q, r = torch.from_numpy(src).to(DEV), torch.from_numpy(tgt).to(DEV)
    nn_search(q, r)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        nn_search(q, r)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"asdf_nn_search {n} x {n}: {ms:.3f} ms = {n * n / ms / 1e6:.1f} G pairs/s", flush=True)


if __name__ == "__main__":
    for c in CASES:
        try:
            time_case(c)
        except Exception:                                         # keep going: one line per variant
            print(f"{c}: FAILED\n{traceback.format_exc()}", flush=True)
    try:
        time_icp()
    except Exception:
        print(f"icp: FAILED\n{traceback.format_exc()}", flush=True)
