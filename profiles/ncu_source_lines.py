#!/usr/bin/env python
"""Warp-stall samples per CUDA source line of one kernel of an .ncu-rep (needs --import-source on, -lineinfo).
usage: ncu_source_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Warp Stall Sampling (All Samples)" in r)
hdr = rows[hi]
sa, ie, si = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
data = []
fname = ""
for r in rows[hi + 1:]:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    if len(r) <= max(sa, ie) or r[0].startswith("0x"):
        continue
    try:
        data.append((int(r[sa]), int(r[ie]), r[0], r[si][:110]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print(f"{kre}: {tot} samples over {len(data)} source lines")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100 * d[0] / tot:5.1f}%  inst={d[1]:>11}  L{d[2]:>4}  {d[3]}")
