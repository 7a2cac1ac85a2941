#!/usr/bin/env python
"""bench.py -- dense-grid SDF query + marching cubes at 256^3 hand+object (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--N 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = the whole hot path for one synthetic sample (utils/mesh.py::create_mesh_combined_decoder
semantics): pass 1 over the N^3 grid, bbox re-grid, pass 2, marching cubes of the hand and of the
object field.  One step therefore answers 2*N^3 hand+object SDF queries.

  value   M hand+obj queries/s, inputs (latent, pose, weights) already resident in HBM, timed with
          CUDA events on the launching stream, max over ranks.
  e2e     the same work through the public API (create_mesh_combined_decoder[_slab]) with the
          per-step inputs coming from pinned host memory and the meshes read back to the host and
          written as PLY -- H2D/D2H inside the timed region.
  N > 1   z-slab sharding of every sample across the ranks (strong scaling): one all_reduce of the
          bbox (+ kernel flags), neighbour exchange of the boundary planes, one all_gather of the sizes,
          un-padded gather of the mesh pieces (hand surface to rank 0, object surface to rank 1).
  --impl reference   the reference's own torch-CPU path (oracle port: /root/reference cannot travel
          to the GPU box) on all host cores, on a bounded sample of the same workload.
  --samples S        number of distinct synthetic samples cycled through (16 = config #3, 1024 = config #4).
  N > 1 additionally reports ``sample_parallel``: the reference's own multi-GPU mode (dist_reconstruct.py:63-84,
          independent samples per rank, no communication) through the pipelined batch API, end to end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_MIN = 2_086_912          # algorithmic FLOP per hand+obj query after folding
F_REF = 3_147_776          # FLOP per query as the reference executes it
N_SAMPLES = 16             # BASELINE config #3: batch of 16 synthetic latents / poses (--samples 1024: config #4)
DECODER_INIT = "default"   # default-initialised SeparateDecoder, last-layer biases shifted (no wiring)
DECODER_DESC = "SeparateDecoder 5x512, both/9, torch default init + last-layer bias shift (SURVEY 8d)"
NCU_TRAFFIC_BYTES = None   # filled from profiles/ by _ncu_traffic()
STEP_SYNC = os.environ.get("ALIGNSDF_BENCH_STEP_SYNC", "0") == "1"


NCU_TRAFFIC_FILES = ("r02_ncu_tc_eval_f8_256.txt", "r01_ncu_tc3_eval_256.txt")


def _ncu_traffic():
    """dram bytes (read + write) per 256^3 launch of the dominant kernel, from the committed ncu summary."""
    path = next((q for q in (os.path.join(ROOT, "profiles", f) for f in NCU_TRAFFIC_FILES) if os.path.exists(q)), None)
    if path is None:
        return None
    tot, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[2].replace(",", "")) * unit_scale.get(f[1], 1.0)
    return tot or None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), tflops_burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        if os.environ.get("ALIGNSDF_BENCH_NO_CLOCKS"):        # diagnostic: is the sampler itself perturbing the run?
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, power_w_max=max(pw) if pw else None, samples=len(self.rows))


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's torch-CPU path (oracle port)
# ----------------------------------------------------------------------------
def workload_config(N, samples, world):
    """The ``config`` object of the JSON line: the workload only, identical in both arms (run-dependent figures
    live under ``derived``)."""
    return dict(workload=f"{N}^3 hand+obj, 2 passes + 2 marching cubes, {samples} synthetic latents/poses",
                decoder=DECODER_DESC, parallelism=f"zslab{world}" if world > 1 else "single",
                l2="a step writes 4 volumes of %.0f MB (2 outputs x 2 passes) = %.0f MB per rank, which %s the 126 MB L2; "
                   "the inputs of the path (4 MB weight stream, latent, pose) are L2-resident by design"
                   % (4 * N ** 3 / world / 1e6, 16 * N ** 3 / world / 1e6,
                      "exceeds" if 16 * N ** 3 / world > 126e6 else "does NOT exceed"))


def cpu_reference_rate(N, chunks, warm=1):
    """M hand+obj queries/s of the reference's per-chunk work (grid -> embedding -> cat -> decoder,
    chunk = 2**18 points as in reconstruct.py:93) on all host cores."""
    from alignsdf_b200 import synthetic
    from oracle import alignsdf_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    dec = synthetic.make_decoder(0, init=DECODER_INIT)
    s = synthetic.make_sample(0)
    sd = {k: v.detach() for k, v in dec.state_dict().items()}
    cfg = orc.decoder_cfg(dec)
    P = 2 ** 18
    times = []
    with torch.no_grad():
        for c in range(warm + chunks):
            start = (c * P) % max(N ** 3 - P, 1)
            t0 = time.perf_counter()
            orc.eval_volume(sd, cfg, s.latent, s.specs, s.mano_results, s.obj_results, N, 2.0 / (N - 1),
                            [-1, -1, -1], "reference", P, start, start + P)
            if c >= warm:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return P / sec / 1e6, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    rate, sec, cores = cpu_reference_rate(args.N, max(args.steps, 1), max(args.warmup, 1))
    line = dict(
        metric="hand+obj SDF queries/s (2-pass grid + marching cubes)", value=rate, unit="Mq/s",
        impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong" if args.gpus > 1 else "weak", vs_baseline=None, dtype="f32",
        data="synthetic",
        config=workload_config(args.N, args.samples, args.gpus),
        cpu_baseline=dict(value=rate, unit="Mq/s", cores=cores, kind="port",
                          sample=f"{args.steps} chunks of 2^18 grid points of the {args.N}^3 workload through "
                                 "the oracle port of the reference's torch-CPU path (marching cubes excluded: "
                                 "scikit-image is not installable)"),
        e2e=dict(value=rate, unit="Mq/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        wall_s=time.perf_counter() - t_all)
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--N", type=int, default=256)
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--cpu-chunks", type=int, default=6, help="chunks of 2^18 points for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--check", action="store_true",
                    help="N > 1: assert the z-slab mesh of sample 0 equals the single-GPU mesh bit for bit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from alignsdf_b200 import engine, mesh as amesh, slab, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = args.N
    W, K = max(args.warmup, 3), args.steps
    S = max(1, min(args.samples, W + K))         # distinct samples actually visited

    dec = synthetic.make_decoder(0, init=DECODER_INIT)
    host_samples = []
    pin = lambda t: t.contiguous().pin_memory()
    for s in synthetic.make_batch(S):           # per-step inputs live in PINNED host memory
        host_samples.append(synthetic.Sample(pin(s.latent), {k: pin(v) for k, v in s.mano_results.items()},
                                             {k: pin(v) for k, v in s.obj_results.items()}, s.specs))
    h2d_bytes = sum(t.numel() * 4 for t in [host_samples[0].latent, *host_samples[0].mano_results.values(),
                                            *host_samples[0].obj_results.values()])
    eng = engine.get_engine(dec, dev)          # weights packed + resident in HBM once

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(i, bound):
        """Device-resident step: 2 passes + 2 marching cubes (+ slab collectives when world > 1)."""
        if world == 1:
            r = amesh._two_pass_verified(bound, N, 3, "reference", False)
            g = r["grid"][0].tolist()
            for vol in (r["hand"], r["obj"]):
                engine.marching_cubes(vol[0].view(N, N, N), 0.0, [g[0]] * 3, g[1:4], check_range=False)
        else:
            slab.reconstruct_slab(slab.gpu_backend(bound, N, spread=True), N, rank, world, spread=True)
            if STEP_SYNC:
                torch.cuda.synchronize(dev)

    def bind(i):
        s = host_samples[i % S].to(dev)
        return eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)

    # ---------------- device-resident timing ----------------
    bounds = [bind(i) for i in range(S)]
    # untimed set-up: one pass over the batch, so that every sample's packed blocks and calibration exist and the
    # caching allocator has seen every (sample-dependent) mesh buffer size -- otherwise the first visit of each
    # sample, inside the timed region, pays cudaMalloc + device synchronisation on every rank
    for i in range(S):
        step_device(i, bounds[i])
    for i in range(W):
        step_device(i, bounds[i % S])
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    engine.KERNEL_EVENTS, engine.MC_EVENTS = [], []
    l0 = engine.LAUNCHES["count"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    marks = [e0]
    for i in range(K):
        step_device(i, bounds[i % S])
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    e1.record()
    barrier()
    step_ms = sorted(a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:]))
    launches = engine.LAUNCHES["count"] - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    k1, k1_total_ms = {}, 0.0
    for kind, nq, a, b in engine.KERNEL_EVENTS:
        if nq > 4 * engine.CALIB_POINTS:                 # everything but the calibration launches
            t_k = a.elapsed_time(b)
            k1_total_ms += t_k
            if nq >= N ** 3 // max(world, 1) // 2:       # the grid passes themselves
                k1.setdefault(kind, []).append((t_k, nq))
    mc = [(nb, a.elapsed_time(b) + c.elapsed_time(d)) for nb, a, b, c, d in engine.MC_EVENTS]
    engine.KERNEL_EVENTS = engine.MC_EVENTS = None
    kinds_used = sorted(set().union(*[b.kinds_used for b in bounds]))
    product_kind = engine.LEVEL_NAMES[eng.level]

    # ---------------- the same device-resident loop with pass 1 on the exact kind (transparency) ----------------
    exact_bbox = None
    if eng.fast_tau(N) is not None:
        engine.FAST_BBOX = False
        try:
            for i in range(2):
                step_device(i, bounds[i % S])
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for i in range(K):
                step_device(i, bounds[i % S])
            x1.record()
            barrier()
            exact_bbox = x0.elapsed_time(x1)
        finally:
            engine.FAST_BBOX = True

    # ---------------- the all-fp16 kind (the fallback of the calibrated kind), one grid pass, for the record ----------------
    f16x3_ms = None
    if world == 1 and bounds[0].tc_ok:
        from alignsdf_b200 import _lib
        qf = engine.make_query(_lib.QUERY_GRID_REFERENCE, N, 0, N ** 3, 2.0 / (N - 1), (-1.0, -1.0, -1.0), bbox_mask=3)
        for _ in range(2):
            y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            y0.record()
            bounds[0].launch_tc(engine.F16X3, qf, N ** 3, True, engine.new_bbox(dev))
            y1.record()
        torch.cuda.synchronize(dev)
        f16x3_ms = y0.elapsed_time(y1)
        bounds[0]._pending.clear()

    # ---------------- 16 samples in ONE launch per pass (config #3; single GPU) ----------------
    batched = None
    if world == 1 and S >= 2 and N <= 256:
        SB = min(S, 16)
        bb = eng.bind_batch([(s.latent, s.specs, s.mano_results, s.obj_results) for s in (h.to(dev) for h in host_samples[:SB])])
        for _ in range(2):
            r = amesh._two_pass_verified(bb, N, 3, "reference", False)
        torch.cuda.synchronize(dev)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        r = bb.two_pass(N, 3, "reference", eng.level if eng.level < engine.LEVEL_SIMT else None, False)
        b1.record()
        torch.cuda.synchronize(dev)
        bms = b0.elapsed_time(b1)
        batched = dict(samples_per_launch=SB, ms_two_passes=bms, Mq_per_s=2.0 * N ** 3 * SB / (bms * 1e-3) / 1e6,
                       note="both grid passes of the whole batch as two launches (P-tile base indexed by sample), "
                            "marching cubes not included")
        del r, bb

    # ---------------- end-to-end through the public API ----------------
    tmp = tempfile.mkdtemp(prefix="asdf_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    d2h_bytes = [0]

    def step_e2e(i):
        s = host_samples[i % S].to(dev)          # H2D of this step's inputs (pinned -> device)
        prefix = os.path.join(tmp, f"s{rank}_{i % 2}")
        if world == 1:
            res = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results,
                                                     s.obj_results, None, s.specs, prefix, N=N)
        else:
            # hand surface stitched / filtered / written by rank 0, object surface by rank 1
            res = slab.create_mesh_combined_decoder_slab(True, True, False, dec, s.latent, s.mano_results,
                                                         s.obj_results, None, s.specs, prefix, N=N, spread=True)
        if res is not None:
            d2h_bytes[0] = sum(m.vertices.nbytes + m.faces.nbytes for m in res.values() if m is not None)

    for i in range(W):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---------------- end-to-end, batch API: host stages overlapped across samples ----------------
    # world == 1: the K samples of the step loop.  world > 1: SAMPLE-PARALLEL mode (dist_reconstruct.py:63-84):
    # rank r reconstructs samples r, r + world, ... of the same K on its own GPU, no communication.
    mine = list(range(rank, K, world))
    names = [os.path.join(tmp, f"p{rank}_{j % 2}") for j in range(len(mine) + 2)]
    amesh.create_meshes_pipelined(dec, [host_samples[i % S] for i in range(2)], names[:2], N=N, device=dev)
    barrier()
    t0 = time.perf_counter()
    if mine:
        amesh.create_meshes_pipelined(dec, [host_samples[i % S] for i in mine], names[:len(mine)], N=N, device=dev)
    torch.cuda.synchronize(dev)
    e2e_pipe_s = time.perf_counter() - t0

    check = None
    if args.check and world > 1:
        s = host_samples[0].to(dev)
        a = slab.create_mesh_combined_decoder_slab(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                                   None, s.specs, os.path.join(tmp, "chk_slab"), N=N)
        if rank == 0:
            b = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                                   None, s.specs, os.path.join(tmp, "chk_one"), N=N)
            import numpy as np
            check = all(a[t] is not None and b[t] is not None and np.array_equal(a[t].vertices, b[t].vertices)
                        and np.array_equal(a[t].faces, b[t].faces) for t in ("hand", "obj"))
            same_files = all(open(os.path.join(tmp, f"chk_slab_{t}.ply"), "rb").read() ==
                             open(os.path.join(tmp, f"chk_one_{t}.ply"), "rb").read() for t in ("hand", "obj"))
            check = dict(slab_mesh_equals_single_gpu=bool(check), ply_bytes_equal=bool(same_files),
                         faces={t: int(b[t].faces.shape[0]) for t in ("hand", "obj")})
            assert check["slab_mesh_equals_single_gpu"] and check["ply_bytes_equal"], check
        barrier()

    # ---------------- max over ranks ----------------
    t = torch.tensor([ms, e2e_s * 1e3, e2e_pipe_s * 1e3, exact_bbox or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nb = torch.tensor([d2h_bytes[0]], dtype=torch.int64, device=dev)       # meshes are read back by two ranks
        dist.all_reduce(nb, op=dist.ReduceOp.SUM)
        d2h_bytes[0] = int(nb[0])
    ms, e2e_ms, pipe_ms, exact_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    queries = 2.0 * N ** 3 * K
    value = queries / (ms * 1e-3) / 1e6
    e2e_value = queries / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        global NCU_TRAFFIC_BYTES
        NCU_TRAFFIC_BYTES = _ncu_traffic() if N == 256 else None
        peaks = measured_peaks()
        dtype = {"f16+2xe4m3": "f16 main + 2 e4m3 correction products (fp32 accumulate)",
                 "f16x3": "3 f16 products (hi*hi + lo*hi + hi*lo, fp32 accumulate)", "simt": "f32"}[product_kind]
        line = dict(
            metric="hand+obj SDF queries/s (2-pass grid + marching cubes)", value=value, unit="Mq/s",
            n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K, higher_is_better=True,
            scaling="strong" if world > 1 else "weak", vs_baseline=None, dtype=dtype,
            data="synthetic", impl="b200",
            config=workload_config(N, args.samples, world),
            derived=dict(samples_visited=S, meshes_per_s=K / (ms * 1e-3), decoder_evals_Mps=2 * value),
            kernel=dict(selected=product_kind, kinds_launched=kinds_used, calibration_err=eng.calib,
                        bbox_pass=("f16x1 (single fp16 product) with sign threshold tau = %.3e + exact re-evaluation of "
                                   "the points within tau of zero" % eng.fast_tau()) if eng.fast_tau() else product_kind,
                        f16x3_ms_per_launch=f16x3_ms, stats=dict(engine.STATS)),
            e2e=dict(value=e2e_value, unit="Mq/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=d2h_bytes[0],
                     ms_per_step=e2e_ms / K, meshes_per_s=K / (e2e_ms * 1e-3)),
            gpu_launches=launches, clocks=clocks,
            step_ms=dict(min=step_ms[0], median=step_ms[len(step_ms) // 2], max=step_ms[-1]))
        pipe = dict(value=queries / (pipe_ms * 1e-3) / 1e6, unit="Mq/s", ms_per_step=pipe_ms / K,
                    meshes_per_s=K / (pipe_ms * 1e-3),
                    api="mesh.create_meshes_pipelined: same files, consecutive samples overlapped (host inputs, PLY written)")
        if world == 1:
            line["e2e"]["pipelined_batch"] = pipe
        else:
            pipe["mode"] = (f"sample-parallel (dist_reconstruct.py:63-84): {K} samples dealt round-robin to {world} ranks, "
                            "no communication, end to end, max over ranks")
            line["sample_parallel"] = pipe
        if exact_ms > 0:
            line["exact_bbox_pass"] = dict(
                value=queries / (exact_ms * 1e-3) / 1e6, unit="Mq/s", ms_per_step=exact_ms / K,
                note="the same device-resident loop with ALIGNSDF_B200_FAST_BBOX=0: pass 1 on the exact kind instead of "
                     "the single-product kind + exact re-evaluation of the shell (identical bounding boxes and meshes)")
        if batched is not None:
            line["batched"] = batched
        if check is not None:
            line["check"] = check
        if k1:
            # the step's tensor-core work: both grid passes (pass 1 on the single-product kind + exact
            # re-evaluation of the shell around the surface once the decoder has an error bound for it; pass 2 on the
            # calibrated kind) over ALL tensor-core kernel time of the step, in algorithmic FLOPs of the two passes
            per_kind = {k: dict(ms_per_launch=sum(t for t, _ in v) / len(v), launches=len(v),
                                queries_per_launch=sum(q for _, q in v) / len(v)) for k, v in k1.items()}
            nq_step = 2.0 * N ** 3 / world                   # queries of this rank per step
            tms = k1_total_ms / K
            ach = nq_step * F_MIN / (tms * 1e-3) / 1e12
            p2 = k1.get(product_kind)
            line["roofline"] = dict(bound="tensor",
                                    kernel="tc_eval_kernel: pass 1 <%s>%s, pass 2 <%s>" % (
                                        "f16x1" if "f16x1" in k1 else product_kind,
                                        " + exact re-evaluation of the shell" if "f16x1" in k1 else "", product_kind),
                                    achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"],
                                    traffic=NCU_TRAFFIC_BYTES,
                                    traffic_note="dram__bytes_read+write of one 256^3 launch, ncu --set full "
                                                 "(profiles/); algorithmic HBM bytes = 8 B/query",
                                    tensor_kernel_ms_per_step=tms, queries_per_step=nq_step, flop_per_query=F_MIN,
                                    frac_of_burst=ach / peaks["tflops_burst"], peak_source=peaks["source"],
                                    per_kind=per_kind)
            if p2:
                t2 = sum(t for t, _ in p2) / len(p2)
                q2 = sum(q for _, q in p2) / len(p2)
                a2 = q2 * F_MIN / (t2 * 1e-3) / 1e12
                # issued tensor work in fp16-MMA time on padded shapes: 3 fp16 products, or 1 fp16 + 2 fp8 at twice the rate
                issued = a2 * (2.0 if product_kind == "f16+2xe4m3" else 3.0) * (2 * 2 * 524288) / F_MIN
                line["roofline"]["exact_pass"] = dict(kernel=f"tc_eval_kernel<{product_kind}>", ms_per_launch=t2,
                                                      queries_per_launch=q2, achieved=a2, frac=a2 / peaks["tflops"],
                                                      issued_tflops_f16_equiv=issued,
                                                      Mq_per_s_kernel=q2 / (t2 * 1e-3) / 1e6)
        if mc:
            nb = sum(b for b, _ in mc) / len(mc)
            tm = sum(t for _, t in mc) / len(mc)
            line["roofline_mc"] = dict(bound="hbm", kernel="mc_classify + mc_scan_* + mc_compact + mc_emit",
                                       achieved=nb / (tm * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s",
                                       frac=nb / (tm * 1e-3) / 1e9 / peaks["hbm"], bytes_alg_per_surface=nb,
                                       ms_per_surface=tm, formula="4 N^3 + 12 V + 12 F per surface")
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, cores = cpu_reference_rate(N, args.cpu_chunks)
            line["cpu_baseline"] = dict(value=rate, unit="Mq/s", cores=cores, kind="port",
                                        sample=f"{args.cpu_chunks} chunks of 2^18 grid points of the same {N}^3 workload "
                                               f"({sec:.2f} s/chunk) through the oracle port of the reference's torch-CPU path")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
