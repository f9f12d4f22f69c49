#!/usr/bin/env python
"""Per-kernel counts of the Blackwell tensor / TMA instructions in libnrv.so (cuobjdump -sass), as a markdown table.
usage: sass_summary.py [out.md]      (runs here: cuobjdump needs no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "nanoreviser_b200", "libnrv.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR", "HMMA", "LDSM", "SYNCS"]
counts, cur = collections.OrderedDict(), None
for ln in sass.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0].replace("nrv::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
used = [k for k in MN if any(c[k] for c in counts.values())]
out = ["# SASS summary of nanoreviser_b200/libnrv.so (`cuobjdump -sass`, sm_100a): tensor-core / tensor-memory / TMA instructions per kernel", "",
       "UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA tensor "
       "load / store, UTMAPF = TMA prefetch, UBLKCP = bulk copy (cp.async.bulk), UTCBAR = tcgen05.commit, HMMA = mma.sync, LDSM = ldmatrix.", "",
       "| kernel | instructions | " + " | ".join(used) + " |", "|---|---|" + "---|" * len(used)]
for k, c in counts.items():
    if any(c[m] for m in used):
        out.append("| `%s` | %d | " % (k, c["_total"]) + " | ".join(str(c[m]) if c[m] else "" for m in used) + " |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out.append("| **all %d kernels** | %d | " % (len(counts), tot["_total"]) + " | ".join(str(tot[m]) for m in used) + " |")
text = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
