#!/usr/bin/env python
"""tools/ncu_sass.py -- per-stage digest of an `ncu --set full --import-source on` capture.

    python tools/ncu_sass.py gpurun_out/prof.ncu-rep [--per-element NE] [--top 12]

Splits the kernel's SASS at its BAR.SYNC instructions (= the contraction stages) and prints per
stage: executed warp instructions (total / FP64 / shared-memory / global), shared-memory
wavefronts against the conflict-free ideal, stall samples; then the instructions with the most
excessive shared-memory wavefronts and the most stall samples."""
import argparse
import csv
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--per-element", type=float, default=0.0, help="divide counts by this many elements")
ap.add_argument("--top", type=int, default=12)
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], check=True,
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print("#", rows[0][1][:150] if rows[0][0] == "Kernel Name" else "")
hd = rows[hi]
col = {h: i for i, h in enumerate(hd)}


def num(r, k):
    try:
        return float(r[col[k]].replace(",", "") or 0)
    except (ValueError, KeyError, IndexError):
        return 0.0


div = a.per_element or 1.0
stages, cur = [], dict(n=0, inst=0, fp64=0, sh=0, gl=0, wf=0, ideal=0, samples=0, first=None)
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(hd):
        continue
    s = r[col["Source"]]
    op = s.split()[0] if s.split() else ""
    if op.startswith("@"):
        op = s.split()[1] if len(s.split()) > 1 else op
    ex = num(r, "Instructions Executed")
    wf, ideal, exc = num(r, "L1 Wavefronts Shared"), num(r, "L1 Wavefronts Shared Ideal"), num(r, "L1 Wavefronts Shared Excessive")
    smp = num(r, "# Samples")
    cur["inst"] += ex
    cur["n"] += 1
    if op[:2] in ("DF", "DA", "DM") or op.startswith("DSETP"):
        cur["fp64"] += ex
    if op.startswith(("LDS", "STS", "ATOMS")):
        cur["sh"] += ex
    if op.startswith(("LDG", "STG", "RED", "ATOMG", "UBLKCP", "LDGSTS")):
        cur["gl"] += ex
    cur["wf"] += wf
    cur["ideal"] += ideal
    cur["samples"] += smp
    ins.append((len(stages), r[col["Address"]], s, ex, wf, ideal, exc, smp, r))
    if op.startswith("BAR"):
        stages.append(cur)
        cur = dict(n=0, inst=0, fp64=0, sh=0, gl=0, wf=0, ideal=0, samples=0, first=None)
stages.append(cur)
tot = sum(s["inst"] for s in stages)
tots = sum(s["samples"] for s in stages) or 1
unit = "per element" if a.per_element else "total"
print(f"# warp instructions {unit}; stages are the regions between BAR.SYNC")
print(f"{'stage':>5} {'sass':>5} {'inst':>10} {'fp64':>9} {'smem':>9} {'global':>8} {'sh.wavefr':>10} {'ideal':>9} {'samples%':>8}")
for i, s in enumerate(stages):
    print(f"{i:5d} {s['n']:5d} {s['inst']/div:10.1f} {s['fp64']/div:9.1f} {s['sh']/div:9.1f} {s['gl']/div:8.1f} "
          f"{s['wf']/div:10.1f} {s['ideal']/div:9.1f} {100*s['samples']/tots:8.1f}")
print(f"{'all':>5} {sum(s['n'] for s in stages):5d} {tot/div:10.1f} {sum(s['fp64'] for s in stages)/div:9.1f} "
      f"{sum(s['sh'] for s in stages)/div:9.1f} {sum(s['gl'] for s in stages)/div:8.1f} "
      f"{sum(s['wf'] for s in stages)/div:10.1f} {sum(s['ideal'] for s in stages)/div:9.1f}")
print(f"# top {a.top} instructions by excessive shared-memory wavefronts")
for st, ad, s, ex, wf, ideal, exc, smp, r in sorted(ins, key=lambda t: -t[6])[:a.top]:
    if exc > 0:
        print(f"  stage {st} {ad[-5:]} wf={wf/div:9.2f} ideal={ideal/div:9.2f} x{wf/max(ideal,1):4.1f}  {s[:90]}")
print(f"# top {a.top} instructions by stall samples")
reasons = [h for h in hd if h.startswith("stall_") and "Not Issued" not in h]
for st, ad, s, ex, wf, ideal, exc, smp, r in sorted(ins, key=lambda t: -t[7])[:a.top]:
    rs = sorted(((num(r, k), k[6:]) for k in reasons), reverse=True)[:2]
    print(f"  stage {st} {ad[-5:]} {100*smp/tots:5.1f}%  {s[:70]:70s} {rs[0][1]}:{rs[0][0]:.0f} {rs[1][1]}:{rs[1][0]:.0f}")
