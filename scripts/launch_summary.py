#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X): scripts/launch_summary.py X [title]"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1.0)
    tot[r[ik]] += v
    cnt[r[ik]] += 1
allv = sum(tot.values())
print("# %s\n\n| kernel | launches | total us | share |\n|---|---|---|---|" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
for k, v in tot.most_common():
    print("| `%s` | %d | %.1f | %.3f |" % (k[:110], cnt[k], v, v / allv))
