#!/usr/bin/env python
"""tools/partest.py -- multi-GPU parity of the rank-parallel path (run under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/partest.py [--order 3] [--elems 6] [--integrator diffusion]

Every rank owns one block of a Cartesian partition, attaches the NCCL halo and checks, against
the SERIAL problem on the union mesh solved on its own GPU:  P reproduces the global field,
y = P^T A P x, the constrained operator, the assembled diagonal, and Jacobi-PCG (iteration
count within +-1, solution to solver tolerance).  Exit code 0 and "PARTEST OK" on success.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import mfem_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--order", dest="p", type=int, default=3)
ap.add_argument("--elems", dest="n", type=int, default=6, help="elements per side per rank")
ap.add_argument("--integrator", dest="kind", default="diffusion")
ap.add_argument("--vdim", type=int, default=1, help="3: vector space (CEED BP2/BP4), Ordering::byVDIM")
a = ap.parse_args()
vd = a.vdim

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = m.Device(f"cuda:{lr}")
grid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
N = [a.n * g for g in grid]
p = a.p


def build(mesh):
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3), vd, "byVDIM" if vd > 1 else "byNODES")
    it = (m.MassIntegrator if a.kind == "mass" else m.DiffusionIntegrator)(1.0, m.IntRules.Get(3, 2 * p + 3))
    form = m.BilinearForm(dev, fes)
    form.AddDomainIntegrator(it)
    form.Assemble()
    return fes, form


# parallel problem
fes, form = build(m.Mesh(3, N[0], N[1], N[2], True, part=grid, rank=rank))
ids = [m.Halo.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
halo = m.Halo(dev, rank, world, fes.ndofs, fes.halo_tables, ids[0], vdim=vd)
form.SetHalo(halo)
# serial problem on the union mesh (every rank solves it on its own GPU as the reference)
sfes, sform = build(m.Mesh(3, N[0], N[1], N[2], True, part=(1, 1, 1), rank=0))
slat = sfes.halo_tables["ldof_lattice"]
lat = fes.halo_tables["ldof_lattice"]
own = fes.halo_tables["ltdof_ldof"]
nlat = len(slat)
s_of_lat = np.empty(nlat, np.int64)
s_of_lat[slat] = np.arange(nlat)
def vdofs(ids):
    """scalar dof ids -> vdof ids (byVDIM)"""
    return (ids[:, None] * vd + np.arange(vd)[None, :]).ravel()


t2s = torch.from_numpy(vdofs(s_of_lat[lat[own]])).to(dev.tdev)   # my true (v)dofs -> serial (v)dof ids
l2s = torch.from_numpy(vdofs(s_of_lat[lat])).to(dev.tdev)
nlat *= vd
NT, NL = fes.ntdofs * vd, fes.ndofs * vd


def rel(u, v):
    return float(torch.linalg.norm(u - v) / torch.linalg.norm(v))


gen = torch.Generator(device="cpu").manual_seed(7)
xs = torch.rand(nlat, dtype=torch.float64, generator=gen).to(dev.tdev)  # same on every rank
x_t = xs[t2s].contiguous()
# P
x_l = dev.empty(NL)
halo.Mult(x_t, x_l)
dev.sync()
assert torch.equal(x_l, xs[l2s]), "P does not reproduce the global field"
def check(mode):
    global form
    # y = P^T A P x
    y_t, ys = dev.empty(NT), dev.empty(nlat)
    form.Mult(x_t, y_t)
    sform.Mult(xs, ys)
    dev.sync()
    e1 = rel(y_t, ys[t2s])
    assert e1 < 1e-12, f"RAP operator parity {e1}"
    # constrained operator + diagonal
    ess, sess = fes.GetEssentialTrueDofs(), sfes.GetEssentialTrueDofs()
    b_l = m.LinearForm(dev, fes).Assemble()
    b_t = dev.empty(NT)
    halo.MultTranspose(b_l, b_t)
    bs = m.LinearForm(dev, sfes).Assemble()
    dev.sync()
    assert rel(b_t, bs[t2s]) < 1e-12
    A, _, B = form.FormLinearSystem(ess, dev.zeros(NT), b_t)
    As, _, Bs = sform.FormLinearSystem(sess, dev.zeros(nlat), bs)
    A.Mult(x_t, y_t)
    As.Mult(xs, ys)
    dev.sync()
    e2 = rel(y_t, ys[t2s])
    assert e2 < 1e-12, f"constrained operator parity {e2}"
    dg, dgs = dev.empty(NT), dev.empty(nlat)
    form.AssembleDiagonal(dg)
    sform.AssembleDiagonal(dgs)
    dev.sync()
    e3 = rel(dg, dgs[t2s])
    assert e3 < 1e-12, f"diagonal parity {e3}"
    # all-reduced dot
    d_par = halo.Dot(x_t, y_t)
    d_ser = float(torch.dot(xs, ys))
    assert abs(d_par - d_ser) < 1e-12 * abs(d_ser), (d_par, d_ser)
    # PCG
    M, Ms = m.OperatorJacobiSmoother(form, ess), m.OperatorJacobiSmoother(sform, sess)
    res = []
    for (AA, MM, BB, n) in ((A, M, B, NT), (As, Ms, Bs, nlat)):
        cg = m.CGSolver()
        cg.SetRelTol(1e-8); cg.SetMaxIter(1000); cg.SetOperator(AA); cg.SetPreconditioner(MM)
        X = dev.zeros(n)
        cg.Mult(BB, X)
        res.append((cg.GetNumIterations(), cg.GetConverged(), X))
    (it_p, conv_p, X_p), (it_s, conv_s, X_s) = res
    assert conv_p and conv_s and abs(it_p - it_s) <= 1, (it_p, it_s)
    e4 = rel(X_p, X_s[t2s])
    assert e4 < 1e-6, f"PCG solution {e4}"
    # Chebyshev smoother: power method (all-reduced dots), action, and CG preconditioned with it
    start = np.random.default_rng(3).random(nlat)
    Ss = m.OperatorChebyshevSmoother(As, dgs, 3, None, 10, 1e-8, start=start)
    Sp = m.OperatorChebyshevSmoother(A, dg, 3, None, 10, 1e-8, start=start[t2s.cpu().numpy()])
    assert abs(Sp.max_eig_estimate - Ss.max_eig_estimate) < 1e-10 * Ss.max_eig_estimate, \
        (Sp.max_eig_estimate, Ss.max_eig_estimate)
    z_t, zs = dev.empty(NT), dev.empty(nlat)
    Sp.Mult(x_t, z_t)
    Ss.Mult(xs, zs)
    dev.sync()
    e5 = rel(z_t, zs[t2s])
    assert e5 < 1e-10, f"Chebyshev action {e5}"
    res = []
    for (AA, MM, BB, n) in ((A, Sp, B, NT), (As, Ss, Bs, nlat)):
        cg = m.CGSolver()
        cg.SetRelTol(1e-8); cg.SetMaxIter(1000); cg.SetOperator(AA); cg.SetPreconditioner(MM)
        X = dev.zeros(n)
        cg.Mult(BB, X)
        res.append((cg.GetNumIterations(), cg.GetConverged(), X))
    (ic_p, cc_p, Xc_p), (ic_s, cc_s, Xc_s) = res
    assert cc_p and cc_s and abs(ic_p - ic_s) <= 1, (ic_p, ic_s)
    assert rel(Xc_p, Xc_s[t2s]) < 1e-6
    Sp.close(); Ss.close()
    return e1, e2, e3, it_p, it_s, e4


out = {}
out["nccl"] = check("nccl")
halo.connect_p2p(dist)  # from here on P / P^T are peer-memory stores + epoch flags, no NCCL send/recv
out["p2p"] = check("p2p")
e1, e2, e3, it_p, it_s, e4 = out["p2p"]
dist.barrier()
if rank == 0:
    print(f"PARTEST OK (nccl and p2p halo) world={world} grid={grid} p={p} n/rank={a.n} kind={a.kind} vdim={vd}: RAP {e1:.1e} constrained {e2:.1e} "
          f"diag {e3:.1e} pcg its {it_p} vs serial {it_s}, sol {e4:.1e}")
halo.close()
dist.destroy_process_group()
