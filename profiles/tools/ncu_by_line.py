#!/usr/bin/env python
"""Aggregate an ncu report's SASS page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <object-or-cubin with -lineinfo> <kernel-name-substring> [top_n] [inner|outer]
       inner: attribute to the innermost inlined frame (default); outer: to the line of the kernel body itself

ncu's `--page source --csv` gives per-SASS-instruction counters; nvdisasm -g gives the (file, line) of each SASS
address (innermost inlined frame).  The join shows, per source line: share of warp instructions, average active lanes,
share of stall samples."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
mode = sys.argv[5] if len(sys.argv) > 5 else "inner"
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    obj = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", obj], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l][0]
addr2line, cur, group, in_group = {}, None, [], False
for l in sass[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not in_group:
            group, in_group = [], True
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        cur = group[0] if mode == "inner" else group[-1]
        continue
    in_group = False
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ino = hdr.index("stall_no_inst")
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    base = a if base is None else base
    k = addr2line.get(a - base) or ("?", 0)
    v = agg[k]
    v[0] += int(r[ie]); v[1] += int(r[it]); v[2] += int(r[isamp]); v[3] += int(r[ino] or 0)
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values()); thr = sum(v[1] for v in agg.values())
print("kernel %s: warp-inst %.3e thread-inst %.3e avg lanes %.2f samples %d" % (kname, tot, thr, thr / tot, tots))
byf = collections.defaultdict(lambda: [0, 0, 0, 0])
for k, v in agg.items():
    for i in range(4):
        byf[k[0]][i] += v[i]
for k, v in sorted(byf.items(), key=lambda x: -x[1][0]):
    print("%-28s inst%%=%5.1f lanes=%5.1f samp%%=%5.1f noinst%%=%5.1f" % (k, 100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / tots, 100 * v[3] / max(v[2], 1)))
print()
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%-22s %4d inst%%=%5.2f lanes=%5.1f samp%%=%5.2f" % (k[0], k[1], 100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / tots))
