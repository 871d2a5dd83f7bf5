#!/usr/bin/env python
"""tools/pcg_sweep.py -- the BASELINE metric as a table: CEED BP1 / BP3 / BP5 Jacobi-PCG GDOF/s
(dofs x CG iterations / s, 32 fixed iterations per solve as bench.py) and the fused operator's
share of the HBM roofline, for orders p = 1..8 on one B200.

    python tools/pcg_sweep.py [--dofs 10e6] [--bps 3,1,5] [--steps 2] [--out gpurun_out/pcg_sweep.jsonl]

One JSON line per case, flushed as soon as it is measured (the run may be cut by a time limit)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mfem_b200 as m  # noqa: E402
from mfem_b200.capi import call  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dofs", type=float, default=10e6)
ap.add_argument("--bps", default="3,1,5")
ap.add_argument("--orders", default="4,1,2,3,5,6,7,8")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--iters", type=int, default=32)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "pcg_sweep.jsonl"))
a = ap.parse_args()

dev = m.Device("cuda")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    peak = 6650.0
os.makedirs(os.path.dirname(a.out), exist_ok=True)
out = open(a.out, "a")


def case(bp, p):
    t0 = time.time()
    gll = bp == 5
    n = max(2, round(a.dofs ** (1 / 3) / p))
    fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
    ir = m.IntegrationRules(gll).Get(3, 2 * p - 1 if gll else 2 * p + 3)
    it = (m.MassIntegrator if bp == 1 else m.DiffusionIntegrator)(1.0, ir)
    form = m.BilinearForm(dev, fes)
    form.AddDomainIntegrator(it)
    form.Assemble()
    nd = fes.ndofs
    ess = fes.GetEssentialTrueDofs()
    b = m.LinearForm(dev, fes).Assemble()
    A, _, B = form.FormLinearSystem(ess, dev.zeros(nd), b)
    M = m.OperatorJacobiSmoother(form, ess)
    cg = m.CGSolver()
    cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(a.iters)
    cg.SetOperator(A); cg.SetPreconditioner(M)
    cg.check_every = a.iters
    X = dev.zeros(nd)
    for _ in range(a.warmup):
        cg.Mult(B, X)
    dev.sync()
    call("b2f_ctx_kernel_timing", dev.ctx, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        cg.Mult(B, X)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tot, cnt = C.c_double(), C.c_int()
    call("b2f_ctx_kernel_time", dev.ctx, C.byref(tot), C.byref(cnt))
    call("b2f_ctx_kernel_timing", dev.ctx, 0)
    k_ms = tot.value / max(cnt.value, 1)
    k = 1 if bp == 1 else 6
    ab = (k * ir.nq1d ** 3 * 8 + 4 * (p + 1) ** 3) * fes.ne + 16 * nd
    line = {"bp": bp, "p": p, "n": n, "dofs": nd, "elements": fes.ne, "q1d": ir.nq1d,
            "pcg_gdofs": nd * a.iters * a.steps / (ms * 1e-3) / 1e9, "ms_per_solve": ms / a.steps,
            "iterations": cg.GetNumIterations(), "operator_ms": k_ms,
            "operator_gdofs": nd / (k_ms * 1e-3) / 1e9,
            "operator_frac_of_hbm": ab / (k_ms * 1e-3) / 1e9 / peak,
            "operator_share_of_solve": tot.value / ms, "case_seconds": round(time.time() - t0, 1)}
    s = json.dumps(line)
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()
    form.close(); it.close()
    del X, B, b, A, M, cg, form, fes
    torch.cuda.empty_cache()


for bp in (int(v) for v in a.bps.split(",")):
    for p in (int(v) for v in a.orders.split(",")):
        try:
            case(bp, p)
        except Exception as ex:  # keep going: one line per case
            print(json.dumps({"bp": bp, "p": p, "error": str(ex)}), flush=True)
