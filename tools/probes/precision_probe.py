"""CPU precision probe of the tensor-core kernels' arithmetic on NON-engineered decoders.

Emulates k1_tc (fp16 x3 and fp16 + 2 e4m3) from the packed bytes (tests/tc_emulate.py, which the GPU tests
pin to the kernel) and compares with the oracle's fp32 forward on random points of the cube.

    python tools/probes/precision_probe.py [n_points]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from alignsdf_b200 import packer, synthetic, tc_pack  # noqa: E402
from oracle import alignsdf_oracle as orc  # noqa: E402
from tests import tc_emulate  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    variants = [("engineered", 1.0), ("default", 1.0), ("plain", 1.0), ("plain", 4.0), ("plain", 16.0)]
    for init, gain in variants:
        for seed in (0, 1):
            dec = synthetic.make_decoder(seed, init=init, out_gain=gain)
            s = synthetic.make_sample(seed)
            sd = {k: v.detach() for k, v in dec.state_dict().items()}
            with torch.no_grad():
                ref = orc.decode_points(sd, orc.decoder_cfg(dec), s.latent, torch.from_numpy(xyz), s.specs,
                                        s.mano_results, s.obj_results)
            ref = [ref[0].numpy().reshape(-1), ref[1].numpy().reshape(-1)]
            topo = packer.decoder_topology(dec)
            br = packer.fold_decoder(topo, s.latent, s.specs, s.mano_results, s.obj_results)
            st2, sc2 = tc_pack.pack_static_numpy(topo, tc_pack.F16X3)
            sm2, _ = tc_pack.pack_sample_numpy(br, sc2, 2.0, tc_pack.F16X3)
            o2 = tc_emulate.emulate(st2, sm2, xyz, tc_pack.F16X3)
            st3, sc3 = tc_pack.pack_static_numpy(topo, tc_pack.F16_F8)
            sm3, _ = tc_pack.pack_sample_numpy(br, sc3, 2.0, tc_pack.F16_F8)
            o3, vmax = tc_emulate.emulate(st3, sm3, xyz, tc_pack.F16_F8, want_max=True)
            rng_ = max(np.abs(ref[0]).max(), np.abs(ref[1]).max())
            e2 = max(np.abs(o2[d] - ref[d]).max() for d in range(2))
            e3 = max(np.abs(o3[d] - ref[d]).max() for d in range(2))
            e32 = max(np.abs(o3[d] - o2[d]).max() for d in range(2))
            print(f"{init:10s} gain {gain:4.0f} seed {seed}  |sdf|max {rng_:.3f}  f16x3 {e2:.2e}  f16+2e4m3 {e3:.2e}  "
                  f"(between the kinds {e32:.2e})  max act {vmax:.1f}", flush=True)


if __name__ == "__main__":
    main()
