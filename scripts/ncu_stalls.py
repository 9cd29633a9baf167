#!/usr/bin/env python
"""Per-kernel stall breakdown from an .ncu-rep source page: where (which execution-count class of SASS) each stall reason falls.
usage: ncu_stalls.py REP KERNEL_REGEX [top]"""
import csv, subprocess, sys, collections
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 10
out = subprocess.run(["ncu", "-i", rep, "-k", f"regex:{kname}", "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, iex, ia = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}
tot = collections.Counter(); byclass = collections.defaultdict(collections.Counter); data = []
for r in rows[2:]:
    try:
        e = int(r[iex])
    except Exception:
        continue
    rec = {h: int(r[idx[h]] or 0) for h in reasons}
    data.append((int(r[ia], 16), e, r[isrc].strip(), rec))
    for h, v in rec.items():
        tot[h] += v; byclass[e][h] += v
allsum = sum(tot.values())
print(rows[0][1][:70], "total stall samples", allsum)
for h, v in tot.most_common():
    if v: print(f"  {h:24s} {v:8d} {v/allsum:6.3f}")
print("by execution-count class (warp-level exec count of the SASS line): top classes")
for e, c in sorted(byclass.items(), key=lambda kv: -sum(kv[1].values()))[:6]:
    s = sum(c.values()); n = sum(1 for d in data if d[1] == e)
    print(f"  exec={e:9d} lines={n:4d} samples={s:7d} ({s/allsum:5.3f}): " + ", ".join(f"{h[6:]}={v}" for h, v in c.most_common(5)))
base = data[0][0]
print("top lines")
for a, e, src, rec in sorted(data, key=lambda d: -sum(d[3].values()))[:top]:
    s = sum(rec.values())
    print(f"  +0x{a-base:05x} ex={e:9d} samples={s:6d} {src[:70]:70s} " + ", ".join(f"{h[6:]}={v}" for h, v in collections.Counter(rec).most_common(3) if v))
