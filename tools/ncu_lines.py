"""Per-source-line instruction counts of one kernel from an ncu report (needs --import-source on, -lineinfo).
python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches of the kernel may follow each other: take the first table
start = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[start]
iN, iS, iI, iW = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
for r in rows[start + 1:]:
    if not r or r[0] in ("File Path", "Function Name", "Line No"):
        if agg:
            break
        continue
    try:
        n = int(r[iN])
    except ValueError:
        continue
    a = agg.setdefault(n, [r[iS], 0, 0])
    num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    a[1] += num(r[iI])
    a[2] += num(r[iW])
tab = [(n, a[0], a[1], a[2]) for n, a in agg.items()]
tot = sum(t[2] for t in tab) or 1
tots = sum(t[3] for t in tab) or 1
print(f"total instructions {tot}, samples {tots}")
for n, src, ins, smp in sorted(tab, key=lambda t: -t[2])[:top]:
    print(f"{n:5d} {100.0 * ins / tot:5.1f}% inst {100.0 * smp / tots:5.1f}% stall  {src.strip()[:110]}")
