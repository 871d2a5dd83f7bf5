"""GPU tests of the rank-parallel path: the single-rank halo (no NCCL) in-process, and --
when the box has two or more GPUs -- the NCCL halo + all-reduced PCG under torchrun
(tools/partest.py), checked against the serial problem on the union mesh."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_single_rank_halo_is_transparent():
    import torch
    import mfem_b200 as m
    dev = m.Device("cuda")
    p, n = 3, 5
    mesh = m.Mesh(3, n, n, n, True, part=(1, 1, 1), rank=0)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3))
    it = m.DiffusionIntegrator(1.0, m.IntRules.Get(3, 2 * p + 3))
    form = m.BilinearForm(dev, fes)
    form.AddDomainIntegrator(it)
    form.Assemble()
    N = fes.ndofs
    x = torch.rand(N, dtype=torch.float64, device=dev.tdev)
    y0, y1 = dev.empty(N), dev.empty(N)
    form.Mult(x, y0)
    halo = m.Halo(dev, 0, 1, N, fes.halo_tables)
    assert halo.n_tdofs == N
    form.SetHalo(halo)
    form.Mult(x, y1)
    dev.sync()
    assert float(torch.linalg.norm(y1 - y0)) <= 1e-13 * float(torch.linalg.norm(y0))
    ess = fes.GetEssentialTrueDofs()
    b = m.LinearForm(dev, fes).Assemble()
    A, _, B = form.FormLinearSystem(ess, dev.zeros(N), b)
    M = m.OperatorJacobiSmoother(form, ess)
    cg = m.CGSolver()
    cg.SetRelTol(1e-8); cg.SetMaxIter(500); cg.SetOperator(A); cg.SetPreconditioner(M)
    X = dev.zeros(N)
    cg.Mult(B, X)
    assert cg.GetConverged()
    assert abs(halo.Dot(x, y0) - float(torch.dot(x, y0))) < 1e-12 * abs(float(torch.dot(x, y0)))
    form.close(); it.close(); halo.close(); dev.close()


@pytest.mark.parametrize("world,kind,p,vdim", [(2, "diffusion", 3, 1), (2, "mass", 2, 1), (4, "diffusion", 2, 1),
                                               (8, "diffusion", 4, 1), (2, "diffusion", 3, 3), (2, "mass", 2, 3)])
def test_nccl_halo_matches_serial(world, kind, p, vdim):
    """vdim = 3: vector space (CEED BP2/BP4), vdof-level halo tables, Ordering::byVDIM"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + world * 7 + p + 20 * vdim
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "partest.py"), "--order", str(p), "--elems", "5", "--integrator", kind,
           "--vdim", str(vdim)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0 and "PARTEST OK" in out, out[-3000:]
