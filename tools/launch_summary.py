#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total and mean time per kernel.
usage: launch_summary.py launches.csv [markdown]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "").replace("nrv::", "")
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1e-3)
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print("| %s | %d | %.3f | %.1f | %.1f %% |" % (k, cnt[k], v / 1e3, v / cnt[k], 100 * v / s))
print("\nTotal %.1f ms." % (s / 1e3))
