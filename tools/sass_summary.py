#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source sass` output by opcode:
instructions, shared-memory wavefronts (actual / ideal) and stall samples per element.
usage: sass_summary.py file.csv NE"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
NE = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hd = rows[hi]
ix = {h: i for i, h in enumerate(hd)}
tot, ideal, inst, stall = (collections.Counter() for _ in range(4))
allinst = allw = 0


def num(v):
    try:
        return int(float(v))
    except ValueError:
        return 0


for r in rows[hi + 1:]:
    if len(r) < len(hd) or r[0] == "Address":
        continue
    parts = r[ix["Source"]].split()
    if not parts:
        continue
    tok = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    base = tok.split(".")[0]
    op = ".".join(tok.split(".")[:3]) if base in ("LDS", "STS", "LDG", "RED", "LDL", "STL", "ATOMS", "SYNCS",
                                                 "LDGSTS", "UBLKCP", "LDC", "BAR") else base
    w, wi, n = num(r[ix["L1 Wavefronts Shared"]]), num(r[ix["L1 Wavefronts Shared Ideal"]]), num(r[ix["Instructions Executed"]])
    tot[op] += w; ideal[op] += wi; inst[op] += n; allinst += n; allw += w
    stall[op] += num(r[ix["# Samples"]])
print(f"warp-instructions/element {allinst / NE:.1f}   shared wavefronts/element {allw / NE:.1f}")
for op, _ in inst.most_common(30):
    print(f"{op:22s} inst/elem {inst[op] / NE:8.1f}  smem wf/elem {tot[op] / NE:7.1f} ideal {ideal[op] / NE:7.1f}  "
          f"stall samples {stall[op]}")
