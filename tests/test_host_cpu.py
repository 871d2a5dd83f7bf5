"""CPU-side tests (no GPU): the library loads and exports every symbol include/b200fem.h
declares, and the host C++ code (quadrature, basis tables, structured mesh + H1 numbering,
Cartesian partition) reproduces the reference's tables bit for bit (tests/golden fixtures
dumped from the unmodified reference by oracle/ref_driver.cpp)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, golden_names, relerr


def test_library_loads_and_exports_every_declared_symbol():
    from mfem_b200 import capi
    L = capi.lib()
    header = open(os.path.join(ROOT, "include", "b200fem.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b2f_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) > 50
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/b200fem.h but not exported"
        assert name in capi.PROTOTYPES, f"{name} has no ctypes prototype"
    assert L.b2f_version() == 100


def test_ctypes_prototypes_match_the_header_signatures():
    """Every prototype in mfem_b200/capi.py has the argument count and the integer / pointer /
    floating kinds of its declaration in include/b200fem.h (a drifted ctypes signature would
    only show up as a crash on the GPU box)."""
    from mfem_b200 import capi
    header = open(os.path.join(ROOT, "include", "b200fem.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    header = re.sub(r"//[^\n]*", "", header)
    decls = re.findall(r"\b(?:int|void|int64_t|const char \*)\s*\*?\s*(b2f_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    seen = {}
    for name, args in decls:
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        seen[name] = params
    assert len(seen) > 50
    for name, (res, argtypes) in capi.PROTOTYPES.items():
        assert name in seen, f"{name} has a ctypes prototype but no declaration"
        params = seen[name]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, ctypes {len(argtypes)}"
        for k, (decl, ct) in enumerate(zip(params, argtypes)):
            is_ptr = "*" in decl
            if is_ptr:
                assert ct in (capi.vp, C.c_char_p) or hasattr(ct, "contents") or issubclass(ct, C._Pointer), \
                    f"{name} arg {k} ({decl}): pointer in the header, {ct} in ctypes"
            elif re.match(r"(const )?double\b", decl):
                assert ct is C.c_double, f"{name} arg {k} ({decl})"
            elif re.match(r"(const )?int64_t\b", decl):
                assert ct is C.c_int64, f"{name} arg {k} ({decl})"
            elif re.match(r"(const )?size_t\b", decl):
                assert ct is C.c_size_t, f"{name} arg {k} ({decl})"
            elif re.match(r"(const )?int\b", decl):
                assert ct is C.c_int, f"{name} arg {k} ({decl})"


def test_no_device_is_a_loud_error_not_a_fallback():
    """Without a CUDA device the context cannot be created: there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mfem_b200 import capi
    h = C.c_void_p()
    rc = capi.lib().b2f_ctx_create(0, None, C.byref(h))
    assert rc == 5 and not h.value  # B2F_ERR_NO_DEVICE
    assert b"no CPU path" in capi.lib().b2f_last_error()
    import mfem_b200 as m
    with pytest.raises(capi.B2FError):
        m.Device("cuda")
    with pytest.raises(ValueError):
        m.Device("cpu")


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mfem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def meta(g):
    return (int(g["dim"][0]), int(g["D1D"][0]), int(g["Q1D"][0]), int(g["NE"][0]),
            int(g["ndofs"][0]), bool(g["is_mass"][0]))


def test_quadrature_and_basis_tables_bit_exact(golden):
    import mfem_b200 as m
    dim, D1D, Q1D, NE, ndofs, _ = meta(golden)
    gll = bool(golden["is_gll"][0])
    qp, qw = (m.gauss_lobatto if gll else m.gauss_legendre)(Q1D)
    assert np.array_equal(qp, golden["qpts1d"])
    assert np.array_equal(qw, golden["qwts1d"])
    assert np.array_equal(m.gauss_lobatto(D1D)[0], golden["dofpts1d"])
    B, G = m.h1_dof_to_quad(D1D - 1, qp)
    assert np.array_equal(B, golden["B"])
    assert np.array_equal(G, golden["G"])
    assert np.array_equal(m.tensor_weights(dim, qw), golden["W"])
    nd = int(golden["node_D1D"][0])
    Bn, Gn = m.h1_dof_to_quad(nd - 1, qp)
    assert np.array_equal(Bn, golden["Bn"]) and np.array_equal(Gn, golden["Gn"])
    # IntRules.Get picks the same number of points as the reference
    rules = m.IntegrationRules(gll)
    assert rules.Get(dim, 2 * (D1D - 1) - 1 if gll else 2 * (D1D - 1) + 3).nq1d == Q1D


def make_space(g, name):
    import mfem_b200 as m
    dim, D1D, Q1D, NE, ndofs, _ = meta(g)
    n = int(g["n"][0])
    if name == "diff3d_p4_n234":
        mesh = m.Mesh.MakeCartesian3D(2, 3, 4)
    elif dim == 3:
        mesh = m.Mesh.MakeCartesian3D(n, n, n, sfc_ordering=("nosfc" not in name))
    else:
        mesh = m.Mesh.MakeCartesian2D(n, n)
    if int(g["curved"][0]):
        mesh.SetCurvature(3)
        mesh.Transform(m.curve_map)
    return mesh, m.FiniteElementSpace(mesh, m.H1_FECollection(D1D - 1, dim))


@pytest.mark.parametrize("name", golden_names())
def test_mesh_and_h1_numbering_bit_exact(name):
    g = load_golden(name)
    dim, D1D, Q1D, NE, ndofs, _ = meta(g)
    mesh, fes = make_space(g, name)
    assert fes.ne == NE and fes.ndofs == ndofs
    assert np.array_equal(fes.elem_dof, g["elem_dof"])
    assert np.array_equal(fes.dof_map, g["dof_map"])
    assert np.array_equal(fes.bdr_dofs, g["ess_tdof_list"])
    nodes, nd = mesh.nodes_e(fes.elem_vertices)
    assert nd == int(g["node_D1D"][0])
    if int(g["curved"][0]):
        assert relerr(nodes, g["nodesE"]) < 1e-14
    else:
        assert np.array_equal(nodes, g["nodesE"])


# ---- Cartesian partition tables -------------------------------------------------------------
def partition(dim, n, p, grid, sfc=True):
    import mfem_b200 as m
    px, py, pz = grid
    nr = px * py * pz
    out = []
    for r in range(nr):
        mesh = m.Mesh(dim, n[0], n[1], n[2], sfc, part=grid, rank=r)
        out.append(m.FiniteElementSpace(mesh, m.H1_FECollection(p, dim)))
    return out


@pytest.mark.parametrize("dim,n,p,grid", [(3, (4, 4, 4), 2, (2, 1, 1)), (3, (4, 6, 5), 3, (2, 2, 1)),
                                          (3, (4, 4, 4), 1, (2, 2, 2)), (3, (5, 4, 3), 4, (2, 2, 2)),
                                          (2, (6, 5, 1), 3, (2, 2, 1)), (3, (6, 3, 3), 2, (3, 1, 1))])
def test_partition_tables_are_consistent(dim, n, p, grid):
    import mfem_b200 as m
    ranks = partition(dim, n, p, grid)
    serial = m.FiniteElementSpace(m.Mesh(dim, n[0], n[1], n[2], True), m.H1_FECollection(p, dim))
    L = [p * n[a] + 1 for a in range(3)]
    if dim == 2:
        L[2] = 1
    nlat = L[0] * L[1] * L[2]
    owner = np.full(nlat, -1)
    copies = np.zeros(nlat, int)
    for r, f in enumerate(ranks):
        t = f.halo_tables
        lat = t["ldof_lattice"]
        assert len(np.unique(lat)) == f.ndofs          # one local dof per lattice point
        copies[lat] += 1
        own = lat[t["ltdof_ldof"]]
        assert np.all(owner[own] == -1)                  # exactly one owner per point
        owner[own] = r
        assert np.all(np.diff(t["ltdof_ldof"]) > 0)      # ltdof order = ldof order
    assert np.all(owner >= 0) and nlat == serial.ndofs
    assert sum(f.ntdofs for f in ranks) == serial.ndofs
    # the owner is the lowest rank that holds the point (group master, communication.cpp:120)
    lowest = np.full(nlat, 10 ** 9)
    for r, f in enumerate(ranks):
        np.minimum.at(lowest, f.halo_tables["ldof_lattice"], r)
    assert np.array_equal(owner, lowest)
    # pairwise: what r sends to s (shr) is what s expects from r (ext), in the same order
    for r, f in enumerate(ranks):
        t = f.halo_tables
        covered = np.zeros(f.ndofs, bool)
        covered[t["ltdof_ldof"]] = True
        covered[t["ext_ldof"]] = True
        assert covered.all() and len(t["ext_ldof"]) + f.ntdofs == f.ndofs
        # owned-first numbering and interior-first element order
        assert np.array_equal(t["ltdof_ldof"], np.arange(f.ntdofs))
        ed = f.elem_dof.reshape(f.ne, f.dof)
        touches_ext = (ed >= f.ntdofs).any(axis=1)
        assert not touches_ext[:f.n_interior].any() and touches_ext[f.n_interior:].all()
        for k, s in enumerate(t["nbr_rank"]):
            ts = ranks[s].halo_tables
            ks = list(ts["nbr_rank"]).index(r)
            sent = t["ldof_lattice"][t["ltdof_ldof"][t["shr_ltdof"][t["shr_offsets"][k]:t["shr_offsets"][k + 1]]]]
            recv = ts["ldof_lattice"][ts["ext_ldof"][ts["ext_offsets"][ks]:ts["ext_offsets"][ks + 1]]]
            assert np.array_equal(sent, recv)
    # essential dofs: the local dofs on the GLOBAL boundary
    on_bdr = np.zeros(nlat, bool)
    g = np.arange(nlat)
    gx, gy, gz = g % L[0], (g // L[0]) % L[1], g // (L[0] * L[1])
    on_bdr |= (gx == 0) | (gx == L[0] - 1) | (gy == 0) | (gy == L[1] - 1)
    if dim == 3:
        on_bdr |= (gz == 0) | (gz == L[2] - 1)
    for f in ranks:
        assert np.array_equal(np.sort(f.halo_tables["ldof_lattice"][f.bdr_dofs]),
                              np.sort(f.halo_tables["ldof_lattice"][on_bdr[f.halo_tables["ldof_lattice"]]]))
    assert on_bdr.sum() == len(serial.bdr_dofs)


def test_partition_rejects_bad_arguments():
    import mfem_b200 as m
    from mfem_b200.capi import B2FError
    with pytest.raises(B2FError):
        m.FiniteElementSpace(m.Mesh(3, 2, 2, 2, True, part=(3, 1, 1), rank=0), m.H1_FECollection(2, 3))
    with pytest.raises(B2FError):
        m.FiniteElementSpace(m.Mesh(3, 4, 4, 4, True, part=(2, 1, 1), rank=2), m.H1_FECollection(2, 3))


@pytest.mark.parametrize("dim,n,p,grid", [(3, (4, 4, 4), 2, (2, 2, 2)), (3, (4, 2, 2), 3, (2, 1, 1)),
                                          (2, (4, 4, 1), 3, (2, 2, 1)), (3, (6, 4, 2), 1, (2, 2, 1))])
def test_halo_tables_from_group_tables(dim, n, p, grid):
    """The ParFiniteElementSpace binding (b2f_halo_tables_from_groups_h / include/b200fem_groups.hpp):
    hand-built GroupTopology / GroupCommunicator tables of a Cartesian partition -- groups = sets of
    ranks sharing a dof, master = lowest rank, every rank numbering its groups in its OWN order --
    must flatten to the same per-neighbour lists the Cartesian partitioner produces, and sender and
    receiver must enumerate every pair's dofs in the same order."""
    import ctypes as C
    from mfem_b200.capi import call, vp
    ranks = partition(dim, n, p, grid)
    nr = len(ranks)
    L = [p * n[a] + 1 for a in range(3)]
    nlat = L[0] * L[1] * (L[2] if dim == 3 else 1)
    holders = [[] for _ in range(nlat)]
    for r, f in enumerate(ranks):
        for pt in f.halo_tables["ldof_lattice"]:
            holders[pt].append(r)
    # every rank's groups: the sharing sets among its dofs, numbered by first appearance in a
    # rank-dependent traversal (so group numbers differ from rank to rank, as in MFEM)
    group_id = []
    for r, f in enumerate(ranks):
        ids, order = {}, np.argsort((f.halo_tables["ldof_lattice"] * (2 * r + 3)) % 1009, kind="stable")
        for l in order:
            s = tuple(holders[f.halo_tables["ldof_lattice"][l]])
            if len(s) > 1 and s not in ids:
                ids[s] = len(ids) + 1  # group 0 is the local group
        group_id.append(ids)

    def ptr(a):
        return a.ctypes.data_as(vp)

    out = []
    for r, f in enumerate(ranks):
        t, ids = f.halo_tables, group_id[r]
        lat = t["ldof_lattice"]
        ng = len(ids) + 1
        sets = [(r,)] + [None] * len(ids)
        for s, g in ids.items():
            sets[g] = s
        master_rank = np.array([min(s) for s in sets], np.int32)
        master_group = np.array([0] + [group_id[min(s)][s] for s in sets[1:]], np.int32)
        rank_off = np.zeros(ng + 1, np.int32)
        rank_off[1:] = np.cumsum([len(s) for s in sets])
        ranks_flat = np.array([q for s in sets for q in s], np.int32)
        members = [[] for _ in range(ng)]
        for l in np.argsort(lat, kind="stable"):       # a group's dofs in ascending lattice order
            s = tuple(holders[lat[l]])
            members[ids[s] if len(s) > 1 else 0].append(l)
        members[0] = []                                 # the local group is ignored anyway
        ldof_off = np.zeros(ng + 1, np.int32)
        ldof_off[1:] = np.cumsum([len(mm) for mm in members])
        ldofs_flat = np.array([l for mm in members for l in mm], np.int32)
        ldof_ltdof = np.full(f.ndofs, -1, np.int32)
        ldof_ltdof[t["ltdof_ldof"]] = np.arange(f.ntdofs, dtype=np.int32)
        sizes = (C.c_int * 4)()
        args = (r, ng, ptr(master_rank), ptr(master_group), ptr(rank_off), ptr(ranks_flat), ptr(ldof_off),
                ptr(ldofs_flat), f.ndofs, ptr(ldof_ltdof))
        call("b2f_halo_tables_from_groups_h", *args, sizes, None, None, None, None, None, None)
        nt, nn, ns, nx = list(sizes)
        h = dict(ltdof_ldof=np.empty(nt, np.int32), nbr_rank=np.empty(nn, np.int32),
                 shr_offsets=np.empty(nn + 1, np.int32), shr_ltdof=np.empty(ns, np.int32),
                 ext_offsets=np.empty(nn + 1, np.int32), ext_ldof=np.empty(nx, np.int32))
        call("b2f_halo_tables_from_groups_h", *args, sizes, ptr(h["ltdof_ldof"]), ptr(h["nbr_rank"]),
             ptr(h["shr_offsets"]), ptr(h["shr_ltdof"]), ptr(h["ext_offsets"]), ptr(h["ext_ldof"]))
        out.append(h)
        # against the Cartesian partitioner: same true dofs, same neighbours, same dof SETS per neighbour
        assert np.array_equal(h["ltdof_ldof"], t["ltdof_ldof"])
        assert sorted(h["nbr_rank"]) == sorted(t["nbr_rank"])
        for k, q in enumerate(h["nbr_rank"]):
            j = list(t["nbr_rank"]).index(q)
            mine = h["shr_ltdof"][h["shr_offsets"][k]:h["shr_offsets"][k + 1]]
            theirs = t["shr_ltdof"][t["shr_offsets"][j]:t["shr_offsets"][j + 1]]
            assert sorted(mine) == sorted(theirs)
            mine = h["ext_ldof"][h["ext_offsets"][k]:h["ext_offsets"][k + 1]]
            theirs = t["ext_ldof"][t["ext_offsets"][j]:t["ext_offsets"][j + 1]]
            assert sorted(mine) == sorted(theirs)
    # pairwise order: what s sends to q is, point for point, what q expects from s
    for s in range(nr):
        for k, q in enumerate(out[s]["nbr_rank"]):
            sent = out[s]["shr_ltdof"][out[s]["shr_offsets"][k]:out[s]["shr_offsets"][k + 1]]
            pts_s = ranks[s].halo_tables["ldof_lattice"][out[s]["ltdof_ldof"][sent]]
            j = list(out[q]["nbr_rank"]).index(s)
            got = out[q]["ext_ldof"][out[q]["ext_offsets"][j]:out[q]["ext_offsets"][j + 1]]
            pts_q = ranks[q].halo_tables["ldof_lattice"][got]
            assert np.array_equal(pts_s, pts_q), (s, q)


@pytest.mark.parametrize("p,n", [(1, 5), (3, 4), (4, 3)])
def test_threaded_host_loops_match_the_serial_ones(monkeypatch, p, n):
    """The std::thread fan-out of the numbering / partition loops only engages above 65536
    iterations; B2F_HOST_PAR_MIN=1 forces it on small meshes: identical tables, bit for bit."""
    import mfem_b200 as m

    def tables():
        f = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n + 1), m.H1_FECollection(p, 3))
        out = [f.elem_dof.copy(), f.dof_map.copy(), f.bdr_dofs.copy(), f.elem_vertices.copy()]
        for r in range(4):
            g = m.FiniteElementSpace(m.Mesh(3, n, n, n + 1, True, part=(2, 2, 1), rank=r), m.H1_FECollection(p, 3))
            out += [g.elem_dof.copy(), g.bdr_dofs.copy(), g.elem_vertices.copy()]
            out += [g.halo_tables[k].copy() for k in sorted(g.halo_tables)]
        return out

    monkeypatch.setenv("B2F_HOST_PAR_MIN", str(1 << 40))
    monkeypatch.setenv("B2F_HOST_THREADS", "1")
    serial = tables()
    monkeypatch.setenv("B2F_HOST_PAR_MIN", "1")
    monkeypatch.setenv("B2F_HOST_THREADS", "7")
    threaded = tables()
    assert len(serial) == len(threaded)
    for a, b in zip(serial, threaded):
        assert np.array_equal(a, b)
