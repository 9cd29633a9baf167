#!/usr/bin/env python
"""Top stall-sampled SASS instructions of a kernel in an .ncu-rep, with a little context."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 12; ctxn = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try: data.append([int(r[isamp]), int(r[iex]), int(r[ia], 16), r[isrc].strip()])
    except Exception: pass
tot = sum(d[0] for d in data); base = data[0][2]
print(rows[0][1][:80], "samples", tot, "warp-instr", sum(d[1] for d in data))
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in order:
    print("-" * 100)
    for j in range(max(0, i - ctxn), min(len(data), i + 2)):
        s, e, a, src = data[j]
        print(f"{'>>' if j == i else '  '} {s:7d} {s/tot:6.3f} ex={e:9d} +0x{a-base:05x} {src[:100]}")
