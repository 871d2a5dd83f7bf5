#!/usr/bin/env python
"""tools/diagbench.py -- time BilinearForm::AssembleDiagonal (fused element diagonal + unsigned scatter).

    python tools/diagbench.py [--bp 3] [--p 4] [--n 133] [--reps 5]
Prints ms per call (CUDA events) and the rate against the quadrature-data bytes it must read."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mfem_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bp", type=int, default=3)
ap.add_argument("--p", type=int, default=4)
ap.add_argument("--n", type=int, default=133)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = m.Device("cuda")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    peak = 6650.0
gll = a.bp == 5
fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(a.n, a.n, a.n), m.H1_FECollection(a.p, 3))
ir = m.IntegrationRules(gll).Get(3, 2 * a.p - 1 if gll else 2 * a.p + 3)
it = (m.MassIntegrator if a.bp == 1 else m.DiffusionIntegrator)(1.0, ir)
form = m.BilinearForm(dev, fes)
form.AddDomainIntegrator(it)
form.Assemble()
d = dev.empty(fes.ndofs)
for _ in range(2):
    form.AssembleDiagonal(d)
dev.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    form.AssembleDiagonal(d)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
k = 1 if a.bp == 1 else 6
nbytes = k * ir.nq1d ** 3 * 8 * fes.ne + 4 * (a.p + 1) ** 3 * fes.ne + 16 * fes.ndofs
print(f"BP{a.bp} p={a.p} n={a.n}: {fes.ndofs} dofs, AssembleDiagonal {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s "
      f"= {100 * nbytes / ms / 1e6 / peak:.1f}% of {peak:.0f}")
