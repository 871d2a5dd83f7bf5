#!/usr/bin/env python
"""oracle/make_checksums.py -- TEST INFRASTRUCTURE: reference checksums at sizes that fill a GPU.

Runs the UNMODIFIED reference (oracle/_ref/ref_driver, built by oracle/Makefile from
/root/reference) on structured hex meshes too large for full-vector fixtures and records the
norms a parity check needs (SURVEY.md section 6 style):

  op_norm      ||A x||_2, x = Vector::Randomize(1), A = BilinearForm::Mult (PA, unconstrained)
  pcg_*        32 fixed Jacobi-PCG iterations (tests/benchmarks/bench_ceed.cpp:222-243 protocol with
               the Jacobi smoother), b = (1, phi_i), all-boundary Dirichlet, x0 = 0:
               ||X||_2, CGSolver::GetFinalNorm(), GetInitialNorm()

    python oracle/make_checksums.py            # writes tests/golden/checksums.json
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

# (bp, p, n, curved)
CASES = [(3, 4, 16, 0), (3, 4, 24, 0), (3, 4, 40, 0), (3, 1, 64, 0), (3, 2, 32, 0), (3, 8, 8, 0),
         (3, 6, 12, 0), (3, 7, 10, 0),
         (1, 4, 16, 0), (1, 4, 40, 0), (1, 6, 12, 0), (1, 8, 8, 0),
         (5, 7, 10, 0), (5, 7, 8, 1), (5, 7, 24, 1), (5, 4, 24, 1), (5, 1, 64, 0), (5, 1, 48, 1), (1, 1, 64, 0),
         (1, 2, 32, 0), (5, 2, 32, 0),
         # the rest of the metric table (BP1 / BP3 / BP5 x p = 1..8): every cell has a reference checksum
         (3, 3, 21, 0), (3, 5, 13, 0), (1, 3, 21, 0), (1, 5, 13, 0), (1, 7, 10, 0),
         (5, 3, 21, 0), (5, 4, 16, 0), (5, 5, 13, 0), (5, 6, 12, 0), (5, 8, 8, 0)]


def run(mode, bp, p, n, curved):
    cmd = [DRV, "bench", "--mode", mode, "--dim", "3", "--n", str(n), "--p", str(p),
           "--kind", "mass" if bp == 1 else "diffusion", "--quad", "gll" if bp == 5 else "gl",
           "--curved", str(curved), "--iters", "32", "--reps", "1", "--jacobi", "1"]
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE,
                         env=dict(os.environ, MFEM_DEVICE="cpu")).stdout.decode()
    return json.loads(out.strip().splitlines()[-1])


def main():
    path = os.path.join(ROOT, "tests", "golden", "checksums.json")
    res = {}
    if os.path.isfile(path) and "--all" not in sys.argv:
        with open(path) as f:
            res = json.load(f)["cases"]  # keep what is there, add what is missing (--all: redo everything)
    for bp, p, n, curved in CASES:
        if f"bp{bp}_p{p}_n{n}" + ("_curved" if curved else "") in res:
            continue
        m, c = run("mult", bp, p, n, curved), run("pcg", bp, p, n, curved)
        key = f"bp{bp}_p{p}_n{n}" + ("_curved" if curved else "")
        res[key] = dict(bp=bp, p=p, n=n, curved=curved, ndofs=m["ndofs"], ne=m["ne"], op_norm=m["norm"],
                        pcg_iters=c["iters"], pcg_x_norm=c["norm"], pcg_final_norm=c["final_norm"],
                        pcg_initial_norm=c["initial_norm"])
        print(key, res[key], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "checksums.json"), "w") as f:
        json.dump(dict(generator="oracle/make_checksums.py (unmodified reference, MFEM_DEVICE=cpu)",
                       cases=res), f, indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
