#!/usr/bin/env python
"""Per-SOURCE-LINE stall samples of one kernel from an .ncu-rep: joins ncu's SASS page (address, samples, stall reasons)
with nvdisasm's line info of the same kernel in libnrv.so (built with -lineinfo).
usage: ncu_lines.py rep kernel_substring [n]     (kernel_substring matches the mangled name, e.g. 'ILi192ELi128E')"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile

rep, ksub = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "nanoreviser_b200", "libnrv.so")], cwd=tmp, capture_output=True)
line_of = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if ksub not in txt:
        continue
    cur, inside = None, False
    for ln in txt.split("\n"):
        if ln.startswith(".text."):
            inside = ksub in ln
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# a report with several kernels is a sequence of blocks ("Kernel Name", name / header / lines): take the block whose demangled name
# contains NCU_KERNEL (default: the first block)
blocks, cur_b = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_b = [r]
        blocks.append(cur_b)
    elif cur_b is not None:
        cur_b.append(r)
want = os.environ.get("NCU_KERNEL", "")
blk = next((b for b in blocks if want in b[0][1]), blocks[0])
hdr, data = blk[1], [r for r in blk[2:] if r]
ia, isamp, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
base = min(int(r[ia], 16) for r in data)
agg = collections.defaultdict(lambda: [0, collections.Counter(), ""])
tot = 0
for r in data:
    s = int(r[isamp] or 0)
    tot += s
    key = line_of.get(int(r[ia], 16) - base, ("?", 0))
    a = agg[key]
    a[0] += s
    for i, h in stalls:
        a[1][h[6:]] += int(r[i] or 0)
    if s and (not a[2] or s > a[3]):
        a[2] = r[isrc][:60]
        if len(a) < 4:
            a.append(s)
        else:
            a[3] = s
print("total samples", tot)
files = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    if f not in files:
        p = os.path.join(root, "nanoreviser_b200", "csrc", f)
        files[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = files[f][l - 1].strip()[:90] if 0 < l <= len(files[f]) else ""
    top = ", ".join("%s %d" % kv for kv in a[1].most_common(2))
    print("%6d %5.1f%%  %s:%d  [%s]  %s" % (a[0], 100.0 * a[0] / max(tot, 1), f, l, top, text))
