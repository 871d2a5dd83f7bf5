"""World-size-2 (and 4) CPU test of the rank-parallel path's HOST logic over gloo.

The CUDA/NCCL halo (mfem_b200/csrc/halo.cu) cannot run without GPUs; what can be checked
here is everything it is driven by: the partition tables of b2f_cartesian_partition_h and
the exchange protocol they imply (who sends what to whom, in which order, and the fixed
summation order on the owner).  P and P^T are restated in numpy with torch.distributed
point-to-point over gloo, following halo.cu / fem/pfespace.cpp:5760-5927 step by step, and
the identity  P^T A_local P == A_serial   is checked against the
serial C oracle, as is a 2-rank PCG with all-reduced dot products.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_operator(fes, p, kind, ess):
    """Oracle operator (CPU restatement) on one rank's block, from the library's host tables."""
    import mfem_b200 as m
    from oracle import oracle as orc
    dim = fes.dim
    ir = m.IntRules.Get(dim, 2 * p + 3)
    B, G = m.h1_dof_to_quad(p, ir.pts1d)
    gm, off, ind = orc.restriction_build(fes.ne, fes.dof, fes.ndofs, fes.elem_dof, fes.dof_map)
    Bn, Gn = m.h1_dof_to_quad(1, ir.pts1d)
    J, detJ = orc.geometric_factors(dim, 2, ir.nq1d, fes.ne, Bn, Gn, fes.elem_vertices)
    if kind == "mass":
        D = orc.mass_setup(ir.nq1d ** dim, fes.ne, ir.weights, detJ, [1.0])
    else:
        D = orc.diffusion_setup(dim, ir.nq1d, 1, fes.ne, ir.weights, J, [1.0])
    return orc.Operator(dim, p + 1, ir.nq1d, fes.ne, fes.ndofs, gm, off, ind, B, G, D,
                        is_mass=(kind == "mass"), ess=ess)


class GlooHalo:
    """numpy + gloo restatement of b2f_halo (same tables, same order of operations)."""

    def __init__(self, dist, torch, rank, fes):
        self.dist, self.torch, self.rank = dist, torch, rank
        self.t = fes.halo_tables
        self.nl, self.nt = fes.ndofs, fes.ntdofs
        t = self.t
        # unique shared true dofs and their receive-buffer positions, ascending (halo.cu)
        order = np.lexsort((np.arange(len(t["shr_ltdof"])), t["shr_ltdof"]))
        self.unq_pairs = [(int(t["shr_ltdof"][k]), int(k)) for k in order]

    def _exchange(self, send, soff, recv, roff):
        reqs, keep = [], []
        t = self.t
        for k, nb in enumerate(t["nbr_rank"]):
            s = self.torch.from_numpy(np.ascontiguousarray(send[soff[k]:soff[k + 1]]))
            r = self.torch.empty(int(roff[k + 1] - roff[k]), dtype=self.torch.float64)
            keep.append((k, r))
            if s.numel():
                reqs.append(self.dist.isend(s, int(nb)))
            if r.numel():
                reqs.append(self.dist.irecv(r, int(nb)))
        for q in reqs:
            q.wait()
        for k, r in keep:
            recv[roff[k]:roff[k + 1]] = r.numpy()

    def mult(self, x_t):
        t = self.t
        y = np.zeros(self.nl)
        shr = x_t[t["shr_ltdof"]]
        ext = np.zeros(len(t["ext_ldof"]))
        self._exchange(shr, t["shr_offsets"], ext, t["ext_offsets"])
        y[t["ltdof_ldof"]] = x_t
        y[t["ext_ldof"]] = ext
        return y

    def mult_transpose(self, x_l):
        t = self.t
        ext = x_l[t["ext_ldof"]]
        shr = np.zeros(len(t["shr_ltdof"]))
        self._exchange(ext, t["ext_offsets"], shr, t["shr_offsets"])
        y = x_l[t["ltdof_ldof"]].copy()
        for tdof, k in self.unq_pairs:  # fixed order: ascending true dof, then buffer position
            y[tdof] += shr[k]
        return y

    def dot(self, a, b):
        v = self.torch.tensor([float(np.dot(a, b))], dtype=self.torch.float64)
        self.dist.all_reduce(v)
        return float(v.item())


def _worker(rank, world, port, grid, n, p, kind, q):
    try:
        import torch
        import torch.distributed as dist
        import mfem_b200 as m
        from oracle import oracle as orc
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        mesh = m.Mesh(3, n[0], n[1], n[2], True, part=grid, rank=rank)
        fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3))
        halo = GlooHalo(dist, torch, rank, fes)
        lat = fes.halo_tables["ldof_lattice"]
        nlat = (p * n[0] + 1) * (p * n[1] + 1) * (p * n[2] + 1)
        xg = orc.randomize(nlat, 1)                        # the same global vector on every rank
        x_t = xg[lat[fes.halo_tables["ltdof_ldof"]]]
        # essential dofs: local dofs on the global boundary; the masked local operator + P/P^T
        op = _local_operator(fes, p, kind, fes.bdr_dofs)
        x_l = halo.mult(x_t)
        assert np.array_equal(x_l, xg[lat])                # P reproduces the global field locally
        y_t = halo.mult_transpose(op.mult(x_l))
        # constrained: z = x, z[ess] = 0 ; y = P^T A P z ; y[ess] = x[ess]   (on true dofs)
        l2t = np.full(fes.ndofs, -1)
        l2t[fes.halo_tables["ltdof_ldof"]] = np.arange(fes.ntdofs)
        ess_t = l2t[fes.bdr_dofs]
        ess_t = ess_t[ess_t >= 0]

        def A_constrained(v_t):
            z = v_t.copy()
            z[ess_t] = 0.0
            w = halo.mult_transpose(op.mult(halo.mult(z)))
            w[ess_t] = v_t[ess_t]
            return w
        # PCG (linalg/solvers.cpp:869-1050), Jacobi, all-reduced dots
        diag_t = halo.mult_transpose(op.diag())
        dinv = 1.0 / diag_t
        dinv[ess_t] = 1.0
        ones = np.ones(fes.ndofs)
        mop = _local_operator(fes, p, "mass", None)
        b_t = halo.mult_transpose(mop.mult(ones))           # (1, phi_i) assembled across ranks
        b_t[ess_t] = 0.0
        x = np.zeros(fes.ntdofs)
        r = b_t.copy()
        z = dinv * r
        d = z.copy()
        nom0 = nom = halo.dot(d, r)
        r0 = nom0 * 1e-16
        zz = A_constrained(d)
        den = halo.dot(zz, d)
        its = 0
        for i in range(1, 500):
            alpha = nom / den
            x += alpha * d
            r -= alpha * zz
            z = dinv * r
            betanom = halo.dot(r, z)
            its = i
            if betanom <= r0:
                break
            d = z + (betanom / nom) * d
            zz = A_constrained(d)
            den = halo.dot(d, zz)
            nom = betanom
        q.put((rank, lat[fes.halo_tables["ltdof_ldof"]], y_t, x, its, None))
        dist.destroy_process_group()
    except Exception as ex:  # surface the failure in the parent
        import traceback
        q.put((rank, None, None, None, -1, traceback.format_exc()))


@pytest.mark.parametrize("grid,n,p,kind", [((2, 1, 1), (4, 3, 3), 2, "diffusion"),
                                           ((2, 1, 1), (4, 2, 2), 3, "mass"),
                                           ((2, 2, 1), (4, 4, 2), 2, "diffusion")])
def test_rap_operator_and_pcg_match_serial(grid, n, p, kind):
    import torch.multiprocessing as mp
    import mfem_b200 as m
    from oracle import oracle as orc
    world = grid[0] * grid[1] * grid[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, n, p, kind, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for o in out:
        assert o[5] is None, o[5]
    # serial reference on the union mesh (the partition call with a 1x1x1 grid also gives the
    # lattice position of every serial dof)
    smesh = m.Mesh(3, n[0], n[1], n[2], True, part=(1, 1, 1), rank=0)
    sfes = m.FiniteElementSpace(smesh, m.H1_FECollection(p, 3))
    slat = sfes.halo_tables["ldof_lattice"]
    sop = _local_operator(sfes, p, kind, sfes.bdr_dofs)
    nlat = len(slat)
    xg = orc.randomize(nlat, 1)
    y_serial = np.empty(nlat)
    y_serial[slat] = sop.mult(xg[slat])
    y_par = np.full(nlat, np.nan)
    x_par = np.full(nlat, np.nan)
    for rank, lat_t, y_t, x_t, its, _ in out:
        y_par[lat_t] = y_t
        x_par[lat_t] = x_t
    assert not np.isnan(y_par).any()                      # every global dof has exactly one owner
    assert np.linalg.norm(y_par - y_serial) <= 1e-12 * np.linalg.norm(y_serial)
    # serial PCG with the same tolerance
    mop = _local_operator(sfes, p, "mass", None)
    b = mop.mult(np.ones(sfes.ndofs))
    b[sfes.bdr_dofs] = 0.0
    dinv = sop.jacobi(sop.diag())
    X, its_s, conv, fn, hist = sop.pcg(b, dinv, rel_tol=1e-8, max_iter=500)
    its_par = {o[4] for o in out}
    assert len(its_par) == 1                              # all ranks agree (all-reduced scalars)
    assert abs(its_par.pop() - its_s) <= 1
    xs = np.empty(nlat)
    xs[slat] = X
    assert np.linalg.norm(x_par - xs) <= 1e-6 * np.linalg.norm(xs)


@pytest.mark.parametrize("grid,n,p,vdim", [((2, 1, 1), (4, 3, 2), 2, 3), ((2, 2, 1), (4, 4, 2), 3, 2)])
def test_vdof_halo_tables_for_vector_spaces(grid, n, p, vdim):
    """Halo.expand_vdim: the vdof-level tables of a byVDIM vector space (what
    ParFiniteElementSpace's communicator holds for vdim > 1).  Emulates P and P^T for all ranks
    in one process: P reproduces a global vector field on every rank's local vdofs, true vdofs
    stay the head of the local vector, and P^T of the local "ones" counts every global vdof's
    sharers exactly once per owner."""
    import mfem_b200 as m
    world = grid[0] * grid[1] * grid[2]
    spaces = [m.FiniteElementSpace(m.Mesh(3, n[0], n[1], n[2], True, part=grid, rank=r), m.H1_FECollection(p, 3))
              for r in range(world)]
    tabs = [m.Halo.expand_vdim(f.halo_tables, vdim) for f in spaces]
    nlat = (p * n[0] + 1) * (p * n[1] + 1) * (p * n[2] + 1)
    rng = np.random.default_rng(5)
    field = rng.random(nlat * vdim)                     # global field, byVDIM

    def vd(ids):
        return (np.asarray(ids, np.int64)[:, None] * vdim + np.arange(vdim)[None, :]).ravel()
    x_t = []
    for f, t in zip(spaces, tabs):
        nt = f.ntdofs * vdim
        assert np.array_equal(t["ltdof_ldof"], np.arange(nt))          # owned-first prefix survives
        x_t.append(field[vd(f.halo_tables["ldof_lattice"][f.halo_tables["ltdof_ldof"]])])
    # P: owners' shared values travel to the sharers' ext slots (neighbour-pair order must agree)
    counts = np.zeros(nlat * vdim)
    for r, (f, t) in enumerate(zip(spaces, tabs)):
        y = np.full(f.ndofs * vdim, np.nan)
        y[t["ltdof_ldof"]] = x_t[r]
        for k, nb in enumerate(t["nbr_rank"]):
            tn = tabs[nb]
            kk = list(tn["nbr_rank"]).index(r)
            sent = x_t[nb][tn["shr_ltdof"][tn["shr_offsets"][kk]:tn["shr_offsets"][kk + 1]]]
            seg = t["ext_ldof"][t["ext_offsets"][k]:t["ext_offsets"][k + 1]]
            assert len(sent) == len(seg)
            y[seg] = sent
        assert np.array_equal(y, field[vd(f.halo_tables["ldof_lattice"])])
    # P^T of ones: every owner collects one contribution per sharing rank
    for r, (f, t) in enumerate(zip(spaces, tabs)):
        acc = np.ones(f.ntdofs * vdim)
        for k, nb in enumerate(t["nbr_rank"]):
            tn = tabs[nb]
            kk = list(tn["nbr_rank"]).index(r)
            n_in = tn["ext_offsets"][kk + 1] - tn["ext_offsets"][kk]
            seg = t["shr_ltdof"][t["shr_offsets"][k]:t["shr_offsets"][k + 1]]
            assert n_in == len(seg)
            np.add.at(acc, seg, 1.0)
        counts[vd(f.halo_tables["ldof_lattice"][f.halo_tables["ltdof_ldof"]])] += acc
    # each global vdof is owned by exactly one rank; its count = number of ranks holding it
    holders = np.zeros(nlat * vdim)
    for f in spaces:
        holders[vd(f.halo_tables["ldof_lattice"])] += 1
    assert np.array_equal(counts, holders)


def _vector_worker(rank, world, port, grid, n, p, vdim, q):
    """2-rank vector space over gloo: the vdof halo tables drive the exchange; the local block
    operator is the oracle's VectorOperator (byVDIM)."""
    try:
        import torch
        import torch.distributed as dist
        import mfem_b200 as m
        from oracle import oracle as orc
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        mesh = m.Mesh(3, n[0], n[1], n[2], True, part=grid, rank=rank)
        fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3))

        class _V:  # what GlooHalo reads from a space: vdof-level tables and sizes
            pass
        v = _V()
        v.halo_tables = m.Halo.expand_vdim(fes.halo_tables, vdim)
        v.ndofs, v.ntdofs = fes.ndofs * vdim, fes.ntdofs * vdim
        halo = GlooHalo(dist, torch, rank, v)
        lat = fes.halo_tables["ldof_lattice"]
        nlat = (p * n[0] + 1) * (p * n[1] + 1) * (p * n[2] + 1)

        def vd(ids):
            return (np.asarray(ids, np.int64)[:, None] * vdim + np.arange(vdim)[None, :]).ravel()
        xg = orc.randomize(nlat * vdim, 1)
        x_t = xg[vd(lat[fes.halo_tables["ltdof_ldof"]])]
        ess_v = vd(fes.bdr_dofs)
        op = orc.VectorOperator(_local_operator(fes, p, "diffusion", None), vdim, True, ess_v)
        x_l = halo.mult(x_t)
        assert np.array_equal(x_l, xg[vd(lat)])
        y_t = halo.mult_transpose(op.mult(x_l))
        d_t = halo.mult_transpose(op.diag())
        q.put((rank, vd(lat[fes.halo_tables["ltdof_ldof"]]), y_t, d_t, None))
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, None, None, None, traceback.format_exc()))


def test_vector_space_rap_operator_matches_serial():
    import torch.multiprocessing as mp
    import mfem_b200 as m
    from oracle import oracle as orc
    grid, n, p, vdim = (2, 1, 1), (4, 2, 2), 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vector_worker, args=(r, 2, port, grid, n, p, vdim, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    out = [q.get(timeout=240) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
    for o in out:
        assert o[4] is None, o[4]
    smesh = m.Mesh(3, n[0], n[1], n[2], True, part=(1, 1, 1), rank=0)
    sfes = m.FiniteElementSpace(smesh, m.H1_FECollection(p, 3))
    slat = sfes.halo_tables["ldof_lattice"]
    nlat = len(slat)

    def vd(ids):
        return (np.asarray(ids, np.int64)[:, None] * vdim + np.arange(vdim)[None, :]).ravel()
    sop = orc.VectorOperator(_local_operator(sfes, p, "diffusion", None), vdim, True, vd(sfes.bdr_dofs))
    xg = orc.randomize(nlat * vdim, 1)
    y_serial, d_serial = np.empty(nlat * vdim), np.empty(nlat * vdim)
    y_serial[vd(slat)] = sop.mult(xg[vd(slat)])
    d_serial[vd(slat)] = sop.diag()
    y_par, d_par = np.full(nlat * vdim, np.nan), np.full(nlat * vdim, np.nan)
    for rank, ids, y_t, d_t, _ in out:
        y_par[ids] = y_t
        d_par[ids] = d_t
    assert not np.isnan(y_par).any()
    assert np.linalg.norm(y_par - y_serial) <= 1e-12 * np.linalg.norm(y_serial)
    assert np.linalg.norm(d_par - d_serial) <= 1e-12 * np.linalg.norm(d_serial)
