"""Per-kernel time of one training step from an ncu launch list (gpu__time_duration.sum); usage:
NRV_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <count> --csv --log-file X.csv python tools/bench_train.py
python tools/train_launch_summary.py X.csv <steps captured> > profiles/<tag>_train_launches.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
t, n = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    t[name] += float(r[vi].replace(",", "")); n[name] += 1
tot = sum(t.values())
print("# training step, batch 512 x window 13: kernels by device time (ncu launch list, cold-cache serialised launches; %d launches = %.2f steps)\n" % (sum(n.values()), steps))
print("| kernel | launches / step | ms / step | share |\n|---|---|---|---|")
for k, v in t.most_common():
    print("| %s | %.0f | %.3f | %.1f %% |" % (k, n[k] / steps, v / 1e6 / steps, 100 * v / tot))
print("| **total** | %.0f | %.3f | |" % (sum(n.values()) / steps, tot / 1e6 / steps))
