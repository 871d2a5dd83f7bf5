#!/usr/bin/env python
"""tools/opbench.py -- time the fused operator kernel alone (y = R^T A R x) per kernel variant.

    python tools/opbench.py --bp 3 --p 4 --n 96 --variants c0,c1,c2+pf0,v20 [--reps 20]
    python tools/opbench.py --sweep 3,1,5 --variants c0,c1,c2     (p = 1..8)

A variant is a '+'-joined list of cN (elements-per-CTA candidate B2F_NEB_CAND), pfN (L2
prefetch distance B2F_PF, 0 = off), persist (B2F_PERSIST=1), tpe / notpe (B2F_TPE: thread-per-element
kernel of the lowest orders) and g (general pencil kernel).

Prints one line per variant: kernel ms (CUDA events around the launches), GDOF/s and the
achieved algorithmic GB/s (SURVEY.md section 8d bytes) against MEASURED_PEAKS.json.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mfem_b200 as m  # noqa: E402
from mfem_b200.capi import call  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bp", type=int, default=3)
ap.add_argument("--p", type=int, default=4)
ap.add_argument("--n", type=int, default=96)
ap.add_argument("--variants", default="0")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--vdim", type=int, default=1, help="3: vector problem (BP2/4/6), byVDIM ordering")
ap.add_argument("--sweep", default="", help="comma list of BPs: sweep p=1..8")
ap.add_argument("--dofs", type=float, default=20e6)
ap.add_argument("--maxn", type=int, default=100)
ap.add_argument("--numbering", default="h1", help="h1 | dg: probe with a private, contiguous dof range per "
                "element (perfectly coalesced gather/scatter, no sharing) -- an upper bound for index layouts")
a = ap.parse_args()

dev = m.Device("cuda")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    peak = 6650.0


def run_case(bp, p, n, variants, reps):
    gll = bp == 5
    fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
    ir = m.IntegrationRules(gll).Get(3, 2 * p - 1 if gll else 2 * p + 3)
    it = (m.MassIntegrator if bp == 1 else m.DiffusionIntegrator)(1.0, ir)
    vd = a.vdim
    if vd > 1:
        # vector space on the same mesh: the scalar tables with vdim components, byVDIM
        R = m.ElementRestriction.from_space(dev, fes)
        gm, off, ind = R.tables()
        R.close()
        Rv = m.ElementRestriction.from_tables(dev, fes.ne, fes.dof, fes.ndofs, gm, off, ind, vdim=vd, byvdim=True)
        fes.GetElementRestriction = lambda d: Rv
    if a.numbering == "dg":
        ntot = fes.ne * fes.dof
        gm = np.arange(ntot, dtype=np.int32)
        Rd = m.ElementRestriction.from_tables(dev, fes.ne, fes.dof, ntot, gm, np.arange(ntot + 1, dtype=np.int32), gm)
        fes.GetElementRestriction = lambda d: Rd
        fes.ndofs = ntot
    form = m.BilinearForm(dev, fes)
    form.AddDomainIntegrator(it)
    form.Assemble()
    N = fes.ndofs * vd
    x = torch.rand(N, dtype=torch.float64, device=dev.tdev)
    y = dev.empty(N)
    yref = None
    k = 1 if bp == 1 else 6
    ab = (k * ir.nq1d ** 3 * 8 + 4 * (p + 1) ** 3) * fes.ne + 16 * N  # D and indices once per element
    print(f"BP{bp} p={p} n={n}: NE={fes.ne} dofs={N} algorithmic bytes/apply={ab/1e9:.3f} GB", flush=True)
    for v in variants:
        for key in ("B2F_VARIANT", "B2F_PF", "B2F_NEB_CAND", "B2F_GENERAL_KERNEL", "B2F_PERSIST", "B2F_TPE"):
            os.environ.pop(key, None)
        for kv in v.split("+"):
            if kv.startswith("c"):
                os.environ["B2F_NEB_CAND"] = kv[1:]
            elif kv.startswith("pf"):
                os.environ["B2F_PF"] = kv[2:]
            elif kv == "persist":
                os.environ["B2F_PERSIST"] = "1"
            elif kv == "tpe":
                os.environ["B2F_TPE"] = "1"   # thread-per-element kernel (lowest orders)
            elif kv == "notpe":
                os.environ["B2F_TPE"] = "0"
            elif kv == "g":
                os.environ["B2F_GENERAL_KERNEL"] = "1"
            elif kv.startswith("v"):
                os.environ["B2F_VARIANT"] = kv[1:]
        try:
            for _ in range(3):
                form.Mult(x, y)
            dev.sync()
            if yref is None:
                yref = y.clone()
            err = float(torch.linalg.norm(y - yref) / torch.linalg.norm(yref))
            call("b2f_ctx_kernel_timing", dev.ctx, 1)
            for _ in range(reps):
                form.Mult(x, y)
            tot, cnt = C.c_double(), C.c_int()
            call("b2f_ctx_kernel_time", dev.ctx, C.byref(tot), C.byref(cnt))
            call("b2f_ctx_kernel_timing", dev.ctx, 0)
            ms = tot.value / cnt.value
            print(f"  BP{bp} p={p} {v:>10}: {ms:8.4f} ms  {N/ms/1e6:7.2f} GDOF/s  {ab/ms/1e6:8.1f} GB/s  "
                  f"{100*ab/ms/1e6/peak:5.1f}% of {peak:.0f}  relerr_vs_first={err:.1e}", flush=True)
        except Exception as ex:
            print(f"  BP{bp} p={p} {v:>10}: FAILED {ex}", flush=True)
    form.close(); it.close()
    del x, y, yref
    torch.cuda.empty_cache()


variants = a.variants.split(",")
if a.sweep:
    for bp in (int(b) for b in a.sweep.split(",")):
        for p in range(1, 9):
            n = max(4, min(round((a.dofs) ** (1 / 3) / p), a.maxn))
            run_case(bp, p, n, variants, a.reps)
else:
    run_case(a.bp, a.p, a.n, variants, a.reps)
