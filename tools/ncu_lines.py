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
# one table per source file the kernel's SASS maps to (headers, .cuh, the .cu itself): report each, largest first
num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
tables = []
for start in [i for i, r in enumerate(rows) if r and r[0] == "Line No"]:
    hdr = rows[start]
    iN, iS, iI, iW = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    name = next((rows[j][1] for j in range(start - 1, max(start - 4, -1), -1) if rows[j] and rows[j][0] == "File Path"), "?")
    agg = {}
    for r in rows[start + 1:]:
        if not r or r[0] in ("File Path", "Function Name", "Line No"):
            break
        try:
            n = int(r[iN])
        except ValueError:
            continue
        a = agg.setdefault(n, [r[iS], 0, 0])
        a[1] += num(r[iI])
        a[2] += num(r[iW])
    tables.append((sum(a[1] for a in agg.values()), name, agg))
grand = sum(t[0] for t in tables) or 1
for tot, name, agg in sorted(tables, key=lambda t: -t[0]):
    if tot * 200 < grand:
        continue
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"{name}: {tot} instructions ({100.0 * tot / grand:.1f}% of the kernel), {tots} samples")
    for n, (src, ins, smp) in sorted(agg.items(), key=lambda t: -t[1][1])[:top]:
        print(f"{n:5d} {100.0 * ins / max(tot, 1):5.1f}% inst {100.0 * smp / tots:5.1f}% stall  {src.strip()[:110]}")
