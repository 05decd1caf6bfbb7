#!/usr/bin/env python
"""Opcode histogram (executed warp-instructions) from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot = collections.Counter(), collections.Counter(), 0
nwarps = None
for r in rows[h + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    parts = r[ia].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    n = int(r[ie])
    nwarps = nwarps or n
    ops[op] += n
    tot += n
    samp[op] += int(r[isamp] or 0)
print(f"total warp-instructions {tot}; per launched warp {tot / nwarps:.1f}")
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:12s} {v:14d} {100 * v / tot:6.2f}%  stall-samples {samp[k]}")
