#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall-sample shares from an ncu report compiled with -lineinfo.
    python profiles/by_line.py gpurun_out/prof.ncu-rep [topN]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
iL, iS, iI, iW = 0, 1, hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[2] != "-":
        continue  # keep only the per-source-line aggregate rows (Address == '-')
    try:
        data.append((int(r[iI]), int(r[iW]), r[iL], r[iS]))
    except ValueError:
        pass
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total warp instructions {ti}, stall samples {ts}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for d in sorted(data, key=lambda d: -d[1])[:n]:
    print(f"{100 * d[0] / ti:5.1f}% inst {100 * d[1] / ts:5.1f}% samples | L{d[2]:>4s} {d[3].strip()[:120]}")
