#!/usr/bin/env python
"""Opcode histogram of one kernel from an ncu report's SASS page.  usage: ncu_opcodes.py <report.ncu-rep> [top_n]"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ie, te = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ops, opt = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        n, t = float(r[ie]), float(r[te])
    except (ValueError, IndexError):
        continue
    toks = r[si].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    ops[op] += n; opt[op] += t
tot = sum(ops.values())
print("total warp-inst %.3e  thread-inst %.3e" % (tot, sum(opt.values())))
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%-10s warp-inst %5.2f%%  lanes %4.1f" % (op, 100 * n / tot, opt[op] / max(n, 1)))
