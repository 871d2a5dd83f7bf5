"""GPU parity of the LOR batched assembly (csrc/lor.cu, through the C ABI) against the fixtures of the
unmodified reference (tests/golden/lor_*.npz: BatchedLORAssembly on the host) and the numpy oracle
(oracle/oracle_lor.py).  Bars: I and J bit-exact (row order and the column order inside a row are the
reference's serial order), values 1e-12 relative."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "lor_*.npz")))
TOL = 1e-12


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def dev():
    import mfem_b200 as m
    d = m.Device("cuda")
    yield d
    d.close()


def assemble_from_reference_tables(dev, g, ess, x_vert=None):
    import mfem_b200 as m
    from mfem_b200.capi import call, vp
    p, ne, nd, dim = int(g["p"][0]), int(g["NE"][0]), int(g["ndofs"][0]), int(g["dim"][0])
    n1 = p + 1
    R = m.ElementRestriction.from_tables(dev, ne, n1 ** dim, nd, g["gather_map"], g["offsets"], g["indices"])
    xv = dev.to_device(g["X_vert"]) if x_vert is None else x_vert
    cm, cd = dev.to_device(g["c_mass"]), dev.to_device(g["c_diff"])
    ess_d = dev.to_device(ess.astype(np.int32)) if ess.size else None
    h = vp()
    call("b2f_lor_assemble_h1", dev.ctx, R.h, dim, p, vp(xv.data_ptr()), vp(cm.data_ptr()), cm.numel(),
         vp(cd.data_ptr()), cd.numel(), vp(ess_d.data_ptr()) if ess_d is not None else vp(0), int(ess.size), 1,
         C.byref(h))
    return m.LORSparseMatrix(dev, h), R


@pytest.mark.parametrize("name", NAMES)
def test_lor_assembly_from_reference_tables(dev, name):
    g = load(name)
    p, ne, nd = int(g["p"][0]), int(g["NE"][0]), int(g["ndofs"][0])
    A, R = assemble_from_reference_tables(dev, g, np.zeros(0, np.int32))
    V = A.sparse_ij()
    assert np.abs(V - g["sparse_ij"]).max() < TOL * np.abs(g["sparse_ij"]).max()
    assert np.array_equal(A.GetI(), g["I"])
    assert np.array_equal(A.GetJ(), g["J"])
    assert np.abs(A.GetData() - g["data"]).max() < TOL * np.abs(g["data"]).max()
    A.close()
    # with the essential dofs eliminated (DIAG_KEEP), and the matrix applied
    B, _ = assemble_from_reference_tables(dev, g, g["ess_tdof_list"])
    assert np.array_equal(B.GetI(), g["I_bc"]) and np.array_equal(B.GetJ(), g["J_bc"])
    assert np.abs(B.GetData() - g["data_bc"]).max() < TOL * np.abs(g["data"]).max()
    y = dev.empty(nd)
    B.Mult(dev.to_device(g["x"]), y)
    dev.sync()
    assert np.abs(y.cpu().numpy() - g["y_bc"]).max() < TOL * np.abs(g["y_bc"]).max()
    B.close(); R.close()


@pytest.mark.parametrize("name", ["lor_3d_p2_n3", "lor_mesh_fichera_q2_p3", "lor_3d_p6_n2_curved", "lor_2d_p3_n3_curved",
                                  "lor_mesh_star_q3_p4"])
def test_lor_rows_held_by_many_elements(dev, name, monkeypatch):
    """Rows held by more elements than the shared-memory slabs cover take the list-intersecting path
    (unstructured meshes have no bound on the vertex valence); B2F_LOR_MAXE=1 sends every shared row
    down that path."""
    monkeypatch.setenv("B2F_LOR_MAXE", "1")
    g = load(name)
    B, R = assemble_from_reference_tables(dev, g, g["ess_tdof_list"])
    assert np.array_equal(B.GetI(), g["I_bc"]) and np.array_equal(B.GetJ(), g["J_bc"])
    assert np.abs(B.GetData() - g["data_bc"]).max() < TOL * np.abs(g["data"]).max()
    B.close(); R.close()


@pytest.mark.parametrize("name", NAMES)
def test_lor_vertex_coordinates(dev, name):
    import mfem_b200 as m
    from mfem_b200.capi import call, vp
    g = load(name)
    p, ne, k, dim = int(g["p"][0]), int(g["NE"][0]), int(g["node_order"][0]), int(g["dim"][0])
    gp, _ = m.gauss_lobatto(p + 1)
    Bn, _ = m.h1_dof_to_quad(k, gp)
    Bn = np.ascontiguousarray(Bn.reshape(k + 1, p + 1).T)
    nodes = dev.to_device(g["nodes_e"])
    xv = dev.empty(dim * (p + 1) ** dim * ne)
    call("b2f_lor_vertex_coordinates", dev.ctx, dim, p, k + 1, ne, Bn.ctypes.data_as(vp), vp(nodes.data_ptr()),
         vp(xv.data_ptr()))
    assert np.abs(xv.cpu().numpy() - g["X_vert"]).max() < 1e-14 * max(1.0, np.abs(g["X_vert"]).max())


@pytest.mark.parametrize("name,n,kind", [("lor_3d_p1_n3", 3, "diffusion"), ("lor_3d_p4_n2", 2, "diffusion"),
                                         ("lor_3d_p4_n2_mass", 2, "mass"), ("lor_2d_p1_n4", 4, "diffusion")])
def test_lor_end_to_end_own_tables(dev, name, n, kind):
    """Mesh, numbering, restriction tables, vertex coordinates and boundary dofs all from this
    library's own generators (the reference's numbering on generated meshes)."""
    import mfem_b200 as m
    g = load(name)
    p, dim = int(g["p"][0]), int(g["dim"][0])
    mesh = m.Mesh.MakeCartesian3D(n, n, n) if dim == 3 else m.Mesh.MakeCartesian2D(n, n)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, dim))
    a = m.BilinearForm(dev, fes)
    a.AddDomainIntegrator(m.MassIntegrator(2.5) if kind == "mass" else m.DiffusionIntegrator(1.0))
    lor = m.BatchedLORAssembly(dev, fes)
    assert np.abs(lor.GetLORVertexCoordinates().cpu().numpy() - g["X_vert"]).max() < 1e-14
    ess = fes.GetEssentialTrueDofs()
    assert np.array_equal(np.sort(ess), np.sort(g["ess_tdof_list"]))
    A = lor.Assemble(a, ess)
    assert np.array_equal(A.GetI(), g["I_bc"]) and np.array_equal(A.GetJ(), g["J_bc"])
    assert np.abs(A.GetData() - g["data_bc"]).max() < TOL * np.abs(g["data"]).max()
    A.close()


@pytest.mark.parametrize("p,n", [(1, 12), (2, 6), (3, 4), (4, 3), (5, 2), (6, 2), (7, 2), (8, 2),
                                 (1, 64), (2, 40), (4, 24), (6, 12), (8, 10)])
def test_lor_properties_at_size(dev, p, n):
    """Every order the reference has a kernel for, on a curved mesh: the small cases against the numpy
    oracle's CSR stage (on the device's own stencil array), all of them (up to 0.9 M dofs) through
    size-independent properties: constants are in the kernel of the diffusion part (row sums vanish),
    the matrix is symmetric, the vertex-quadrature mass matrix is diagonal and positive."""
    import mfem_b200 as m
    import torch
    from oracle import oracle_lor as L
    mesh = m.Mesh.MakeCartesian3D(n, n, n)
    mesh.SetCurvature(3)
    mesh.Transform(m.curve_map)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3))
    lor = m.BatchedLORAssembly(dev, fes)
    N = fes.GetNDofs()
    one = torch.ones(N, dtype=torch.float64, device=dev.tdev)
    y = dev.empty(N)
    D = lor.Assemble(None, None, mass_coeff=0.0, diff_coeff=1.0)
    D.Mult(one, y)
    dev.sync()
    scale = float(torch.as_tensor(np.abs(D.GetData()).max()))
    assert float(y.abs().max()) < 1e-11 * scale
    gen = torch.Generator(device=dev.tdev).manual_seed(5)
    u = torch.rand(N, dtype=torch.float64, device=dev.tdev, generator=gen)
    v = torch.rand(N, dtype=torch.float64, device=dev.tdev, generator=gen)
    au, av = dev.empty(N), dev.empty(N)
    D.Mult(u, au); D.Mult(v, av)
    dev.sync()
    s1, s2 = float(torch.dot(v, au)), float(torch.dot(u, av))
    assert abs(s1 - s2) < 1e-11 * abs(s1)
    D.close()
    M = lor.Assemble(None, None, mass_coeff=1.0, diff_coeff=0.0, keep_sparse_ij=True)
    M.Mult(one, y)
    dev.sync()
    if N <= 6000:
        R = fes.GetElementRestriction(dev)
        gm, off, ind = R.tables()
        I, J, A = L.fill_csr(p, fes.ne, N, gm, off, ind, M.sparse_ij().reshape(fes.ne, (p + 1) ** 3, 27))
        assert np.array_equal(M.GetI(), I) and np.array_equal(M.GetJ(), J)
        assert np.abs(M.GetData() - A).max() <= 1e-15 * np.abs(A).max()
    # lumped mass: only the diagonal is populated, its sum is the volume of the LOR mesh
    I, J, A = M.GetI(), M.GetJ(), M.GetData()
    row = np.repeat(np.arange(N), np.diff(I))
    assert np.abs(A[row != J]).max() == 0.0
    assert abs(A.sum() - float(y.sum())) < 1e-12 * A.sum() and A[row == J].min() > 0.0
    M.close()


@pytest.mark.parametrize("p,n", [(1, 40), (2, 20), (3, 10), (5, 6), (8, 4), (4, 200)])
def test_lor_2d_properties_at_size(dev, p, n):
    """2D (9-point stencil): every small case against the numpy oracle's CSR stage, all through the
    size-independent properties (constants in the kernel of the diffusion part, symmetry, lumped mass)."""
    import mfem_b200 as m
    import torch
    from oracle import oracle_lor as L
    mesh = m.Mesh.MakeCartesian2D(n, n)
    mesh.SetCurvature(3)
    mesh.Transform(m.curve_map)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 2))
    lor = m.BatchedLORAssembly(dev, fes)
    N = fes.GetNDofs()
    one = torch.ones(N, dtype=torch.float64, device=dev.tdev)
    y = dev.empty(N)
    D = lor.Assemble(None, None, mass_coeff=0.0, diff_coeff=1.0, keep_sparse_ij=True)
    D.Mult(one, y)
    dev.sync()
    assert float(y.abs().max()) < 1e-11 * np.abs(D.GetData()).max()
    gen = torch.Generator(device=dev.tdev).manual_seed(5)
    u = torch.rand(N, dtype=torch.float64, device=dev.tdev, generator=gen)
    v = torch.rand(N, dtype=torch.float64, device=dev.tdev, generator=gen)
    au, av = dev.empty(N), dev.empty(N)
    D.Mult(u, au); D.Mult(v, av)
    dev.sync()
    s1, s2 = float(torch.dot(v, au)), float(torch.dot(u, av))
    assert abs(s1 - s2) < 1e-11 * abs(s1)
    if N <= 6000:
        R = fes.GetElementRestriction(dev)
        gm, off, ind = R.tables()
        I, J, A = L.fill_csr(p, fes.ne, N, gm, off, ind, D.sparse_ij().reshape(fes.ne, (p + 1) ** 2, 9), dim=2)
        assert np.array_equal(D.GetI(), I) and np.array_equal(D.GetJ(), J)
        assert np.abs(D.GetData() - A).max() <= 1e-15 * np.abs(A).max()
    D.close()
    M = lor.Assemble(None, None, mass_coeff=1.0, diff_coeff=0.0)
    I, J, A = M.GetI(), M.GetJ(), M.GetData()
    row = np.repeat(np.arange(N), np.diff(I))
    assert np.abs(A[row != J]).max() == 0.0 and A[row == J].min() > 0.0
    M.close()
