"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/rNN_ncu_launch_summary.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = r[4].split("(")[0].replace("asdf::", "").replace("void ", "").replace("<unnamed>::", "")
    tot[name] += float(r[-1]) / 1e6
    cnt[name] += 1
all_ms = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]} ({len(rows)} launches; cold-cache, serialised: compare SHARES)")
print("# kernel, launches, total_ms, ms_per_launch, share_of_gpu_time")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k}, {cnt[k]}, {tot[k]:.3f}, {tot[k] / cnt[k]:.4f}, {tot[k] / all_ms:.4f}")
