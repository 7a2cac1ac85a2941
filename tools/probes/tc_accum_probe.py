"""GPU probe: how does the tensor core round its fp32 accumulator?  Compares k1_tc (both kinds) with the CPU
emulation of the packed bytes with exact accumulation and with the accumulate rounding measured by acc_round_probe.cu (tests/tc_emulate.py) and with the exact-fp32 kernel, on non-engineered decoders.    python tools/probes/tc_accum_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from alignsdf_b200 import engine, synthetic, tc_pack  # noqa: E402
from tests.tc_emulate import emulate  # noqa: E402

dev = torch.device("cuda:0")
xyz = (torch.rand(1536, 3, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(dev)
for init, gain in (("engineered", 1.0), ("default", 1.0), ("plain", 1.0), ("plain", 4.0), ("plain", 16.0)):
    dec = synthetic.make_decoder(31, init=init, out_gain=gain)
    s = synthetic.make_sample(31).to(dev)
    eng = engine.get_engine(dec, dev)
    bound = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
    hs, os_, _ = bound.eval_points(xyz, path="simt")
    ref = torch.stack([hs, os_]).double().cpu().numpy()
    pre_ref = np.arctanh(np.clip(ref, -0.999999, 0.999999))
    for kind, path in ((tc_pack.F16X3, "f16"), (tc_pack.F16_F8, "f8")):
        h, o, _ = bound.eval_points(xyz, path=path)
        got = torch.stack([h, o]).double().cpu().numpy()
        blocks, _ = bound.tc_blocks(kind, 2.0)
        st, sm = eng.tc_static(kind).cpu().numpy(), blocks[0].cpu().numpy()
        line = [f"{init:10s} g{gain:4.0f} {path:3s} |sdf|max {np.abs(ref).max():.3f}  gpu-vs-fp32 {np.abs(got - ref).max():.2e}"]
        pre = np.arctanh(np.clip(got, -0.999999, 0.999999))
        sel = np.abs(ref) < 0.9
        line.append(f"pre-tanh mean signed err x sign(ref) {np.mean(((pre - pre_ref) * np.sign(pre_ref))[sel]):+.2e}")
        for accum, n in (("exact", 1536), ("hw", 96)):
            e = np.stack(emulate(st, sm, xyz.cpu().numpy()[:n], kind, accum=accum)).astype(np.float64)
            line.append(f"emu[{accum}]-vs-fp32 {np.abs(e - ref[:, :n]).max():.2e} gpu-vs-emu {np.abs(e - got[:, :n]).max():.2e} "
                        f"biteq {np.mean(e.astype(np.float32) == got[:, :n].astype(np.float32)):.3f}")
        print("  ".join(line), flush=True)
