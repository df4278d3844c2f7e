import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:^{kname}$"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
iI, iW = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[2] != "-":
        continue
    try:
        data.append((int(r[iI]), int(r[iW]), r[0], r[1]))
    except ValueError:
        pass
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"{kname}: total warp instructions {ti}, stall samples {ts}")
for d in sorted(data, key=lambda d: -d[0])[:n]:
    print(f"{100 * d[0] / ti:5.1f}% inst {100 * d[1] / ts:5.1f}% samples | L{d[2]:>4s} {d[3].strip()[:130]}")
