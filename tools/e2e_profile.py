"""Host-side profile of the end-to-end call (create_mesh_combined_decoder at N^3): where do the
milliseconds outside the kernels go?   python tools/e2e_profile.py [N]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignsdf_b200 import mesh as amesh, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
dec = synthetic.make_decoder(0, init=os.environ.get("E2E_DECODER", "default"))
samples = synthetic.make_batch(4)
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def step(i):
    s = samples[i % 4].to(dev)
    return amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                              None, s.specs, os.path.join(tmp, f"p{i % 2}"), N=N)


for i in range(3):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(4):
    step(i)
torch.cuda.synchronize()
print(f"e2e {1e3 * (time.perf_counter() - t0) / 4:.2f} ms per step")
pr = cProfile.Profile()
pr.enable()
for i in range(4):
    step(i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
