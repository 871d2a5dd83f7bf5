#!/usr/bin/env python
"""tools/lorbench.py -- time the LOR batched assembly (csrc/lor.cu) and, beside it, the reference's own
CUDA backend (oracle/_ref_cuda/ref_driver_cuda bench --mode lor) on the same problem.

    python tools/lorbench.py --p 4 --n 40 [--reps 5] [--kind diffusion|mass|both] [--no-ref]

Prints one JSON line per arm: dofs, nnz, seconds of Assemble (stencil array + CSR + boundary
elimination; CUDA events for this library, host clock around device syncs for the reference), the
norm of A x for x = Vector::Randomize(1) (a checksum both arms compute), and the algorithmic bytes
(SURVEY-style: vertex coordinates read, CSR written: 3*8 B per E-dof + 12 B per nonzero + 4 B per row).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mfem_b200 as m  # noqa: E402
from mfem_b200.capi import call  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--p", type=int, default=4)
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--kind", default="diffusion")
ap.add_argument("--no-ref", action="store_true")
args = ap.parse_args()

dev = m.Device("cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6448.1) \
    if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6448.1
p, n = args.p, args.n
fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
N = fes.GetNDofs()
ess = fes.GetEssentialTrueDofs()
fes.GetElementRestriction(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
lor = m.BatchedLORAssembly(dev, fes)
e1.record()
torch.cuda.synchronize()
t_vert = e0.elapsed_time(e1) * 1e-3
cm = 0.0 if args.kind == "diffusion" else 1.0
cd = 0.0 if args.kind == "mass" else 1.0
times, ktimes = [], []
A = None
for r in range(args.reps + 1):
    if A is not None:
        A.close()
    torch.cuda.synchronize()
    e0.record()
    A = lor.Assemble(None, ess, mass_coeff=cm, diff_coeff=cd)
    e1.record()
    torch.cuda.synchronize()
    if r:
        times.append(e0.elapsed_time(e1) * 1e-3)
        ktimes.append(A.kernel_ms() * 1e-3)
xs = np.empty(N)
call("b2f_vector_randomize_h", xs.ctypes.data_as(C.c_void_p), xs.size, 1)
x = dev.to_device(xs)
y = dev.empty(N)
A.Mult(x, y)
dev.sync()
best = min(times)
nnz = A.NumNonZeroElems()
alg = 24.0 * fes.ne * (p + 1) ** 3 + 12.0 * nnz + 4.0 * N
print(json.dumps({"arm": "b200fem", "p": p, "n": n, "ndofs": N, "ne": fes.ne, "nnz": nnz,
                  "assemble_seconds": best, "assemble_seconds_mean": float(np.mean(times)),
                  "kernel_seconds": min(ktimes),
                  "vertex_seconds": t_vert, "mdofs_per_s": N / best / 1e6,
                  "algorithmic_gb": alg / 1e9, "algorithmic_gbs": alg / best / 1e9, "frac_of_hbm_peak": alg / best / 1e9 / peak,
                  "ynorm": float(torch.linalg.norm(y))}))
A.close()
drv = os.path.join(ROOT, "oracle", "_ref_cuda", "ref_driver_cuda")
if not args.no_ref and os.path.isfile(drv):
    r = subprocess.run([drv, "bench", "--mode", "lor", "--p", str(p), "--n", str(n), "--reps", str(args.reps),
                        "--lorkind", args.kind], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=dict(os.environ, REF_DRIVER_DEVICE="cuda"), timeout=900)
    if r.returncode == 0:
        j = json.loads(r.stdout.decode().strip().splitlines()[-1])
        j["arm"] = "mfem_cuda"
        j["mdofs_per_s"] = j["ndofs"] / j["assemble_seconds"] / 1e6
        print(json.dumps(j))
    else:
        print(json.dumps({"arm": "mfem_cuda", "failed": r.stderr.decode(errors="replace").strip().splitlines()[-3:]}))
