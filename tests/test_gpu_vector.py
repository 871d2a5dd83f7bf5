"""GPU parity of the vector (vdim = 3) path -- CEED BP2 / BP4 / BP6: VectorMassIntegrator /
VectorDiffusionIntegrator on byVDIM and byNODES spaces -- against fixtures dumped from the
unmodified reference (oracle/ref_driver.cpp do_dump_vector)."""
import numpy as np
import pytest

from conftest import load_golden, relerr, vector_golden_names

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def dev():
    import mfem_b200 as m
    d = m.Device("cuda")
    yield d
    d.close()


def host(t):
    return t.cpu().numpy()


def build(dev, g, ess=None):
    import mfem_b200 as m
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    vdim, byvdim, is_mass = int(g["vdim"][0]), int(g["byvdim"][0]), bool(g["is_mass"][0])
    it = m.MassIntegrator() if is_mass else m.DiffusionIntegrator()
    it.assemble_from_factors(dev, dim, D1D, Q1D, NE, g["B"], g["G"], g["W"], J=g["J"], detJ=g["detJ"])
    R = m.ElementRestriction.from_tables(dev, NE, D1D ** dim, nd, g["gather_map"], g["offsets"], g["indices"],
                                         vdim=vdim, byvdim=byvdim)

    class _Fes:
        pass
    fes = _Fes()
    fes.ne, fes.dof, fes.ndofs = NE, D1D ** dim, nd
    fes.GetElementRestriction = lambda d: R
    a = m.BilinearForm(dev, fes)
    a.AddDomainIntegrator(it)
    a.ess = ess if ess is not None else np.zeros(0, np.int32)
    a.Assemble()
    return a, it, R, vdim


@pytest.mark.parametrize("name", vector_golden_names())
def test_vector_restriction_and_element_kernel(dev, name):
    g = load_golden(name)
    a, it, R, vdim = build(dev, g)
    x = dev.to_device(g["x"])
    xe = dev.empty(R.Height())
    R.Mult(x, xe)
    assert np.array_equal(host(xe), g["xE"])          # vdim restriction: bit-exact
    ye = dev.zeros(R.Height())
    it.AddMultPA(xe, ye, vdim=vdim)
    assert relerr(host(ye), g["yE"]) < TOL
    y = dev.empty(len(g["x"]))
    R.MultTranspose(ye, y)
    assert relerr(host(y), g["y"]) < TOL
    a.close(); it.close(); R.close()


@pytest.mark.parametrize("name", vector_golden_names())
def test_vector_operator_and_pcg(dev, name):
    import mfem_b200 as m
    g = load_golden(name)
    ess = g["ess_tdof_list"]
    a, it, R, vdim = build(dev, g, ess=ess)
    n = len(g["x"])
    assert a.Height() == n
    x = dev.to_device(g["x"])
    y, y2, d = dev.empty(n), dev.empty(n), dev.empty(n)
    a.Mult(x, y)
    assert relerr(host(y), g["y"]) < TOL              # fused, one quadrature-data load for 3 components
    a.MultUnfused(x, y2)
    assert relerr(host(y2), g["y"]) < TOL
    a.AssembleDiagonal(d)
    assert relerr(host(d), g["diag"]) < TOL
    A = m.ConstrainedOperator(a)
    A.Mult(x, y)
    assert relerr(host(y), g["y_constrained"]) < TOL
    b = dev.to_device(g["b"])
    A.EliminateRHS(dev.zeros(n), b)
    assert relerr(host(b), g["B_rhs"]) < TOL
    M = m.OperatorJacobiSmoother(a, ess)
    cg = m.CGSolver()
    cg.SetRelTol(float(g["pcg_rtol"][0])); cg.SetMaxIter(2000); cg.SetOperator(A); cg.SetPreconditioner(M)
    X = dev.zeros(n)
    cg.Mult(b, X)
    assert cg.GetConverged() and abs(cg.GetNumIterations() - int(g["pcg_iters"][0])) <= 1
    assert relerr(host(X), g["X_sol"]) < 1e-6
    a.close(); it.close(); R.close()


@pytest.mark.parametrize("name", ["vec_diff3d_p4_n2_byvdim", "vec_mass3d_p3_n2_bynodes", "vec_bp6_3d_p3_n3_byvdim",
                                  "vec_diff3d_p5_n2_curved"])
def test_vector_end_to_end_own_tables(dev, name):
    """mesh -> vector space -> operator -> PCG with the library's own host tables"""
    import mfem_b200 as m
    g = load_golden(name)
    p, n = int(g["p"][0]), int(g["n"][0])
    mesh = m.Mesh.MakeCartesian3D(n, n, n)
    if int(g["curved"][0]):
        mesh.SetCurvature(3)
        mesh.Transform(m.curve_map)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3), vdim=3,
                               ordering="byVDIM" if int(g["byvdim"][0]) else "byNODES")
    gll = bool(g["is_gll"][0])
    ir = m.IntegrationRules(gll).Get(3, 2 * p - 1 if gll else 2 * p + 3)
    it = (m.VectorMassIntegrator if g["is_mass"][0] else m.VectorDiffusionIntegrator)(1.0, ir)
    a = m.BilinearForm(dev, fes)
    a.AddDomainIntegrator(it)
    a.Assemble()
    nv = fes.GetVSize()
    assert nv == len(g["x"])
    ess = fes.GetEssentialTrueDofs()
    assert np.array_equal(ess, g["ess_tdof_list"])
    y = dev.empty(nv)
    a.Mult(dev.to_device(g["x"]), y)
    assert relerr(host(y), g["y"]) < TOL
    b = dev.to_device(g["b"])
    A, _, B = a.FormLinearSystem(ess, dev.zeros(nv), b)
    M = m.OperatorJacobiSmoother(a, ess)
    cg = m.CGSolver()
    cg.SetRelTol(float(g["pcg_rtol"][0])); cg.SetMaxIter(2000); cg.SetOperator(A); cg.SetPreconditioner(M)
    X = dev.zeros(nv)
    cg.Mult(B, X)
    assert abs(cg.GetNumIterations() - int(g["pcg_iters"][0])) <= 1
    assert relerr(host(X), g["X_sol"]) < 1e-6
    a.close(); it.close()


@pytest.mark.parametrize("name", ["diff3d_p7_n2", "diff3d_p8_n2", "bp5_3d_p7_n2", "mass3d_p8_n2"])
@pytest.mark.parametrize("byvdim", [1, 0])
def test_vector_operator_is_the_scalar_one_per_component_at_high_order(dev, name, byvdim):
    """Vector{Diffusion,Mass}Integrator acts on every component with the scalar operator
    (fem/integ/bilininteg_vecdiffusion_pa.cpp, bilininteg_vecmass_pa.cpp).  At D1D >= 8 the fused
    kernel keeps the element's quadrature data in shared memory across the components while the
    y-stage buffer is overlaid on the slab's tail (PipeCfg::ALIAS: the tail is re-fetched per
    component); the scalar fixtures of p = 7, 8 pin that path: y_c = A x_c to 1e-12."""
    import mfem_b200 as m
    g = load_golden(name)
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    vdim = 3
    it = m.MassIntegrator() if bool(g["is_mass"][0]) else m.DiffusionIntegrator()
    it.assemble_from_factors(dev, dim, D1D, Q1D, NE, g["B"], g["G"], g["W"], J=g["J"], detJ=g["detJ"])
    R = m.ElementRestriction.from_tables(dev, NE, D1D ** dim, nd, g["gather_map"], g["offsets"], g["indices"],
                                         vdim=vdim, byvdim=byvdim)

    class _Fes:
        pass
    fes = _Fes()
    fes.ne, fes.dof, fes.ndofs = NE, D1D ** dim, nd
    fes.GetElementRestriction = lambda d: R
    a = m.BilinearForm(dev, fes)
    a.AddDomainIntegrator(it)
    a.Assemble()
    rng = np.random.default_rng(5)
    comps = [g["x"], rng.random(nd), rng.random(nd) - 0.5]
    xv = np.stack(comps, axis=1).ravel() if byvdim else np.concatenate(comps)
    yv = dev.empty(nd * vdim)
    a.Mult(dev.to_device(xv), yv)
    yh = host(yv)
    # scalar reference: the fixture's y for component 0, the same (linear) operator for the others
    Rs = m.ElementRestriction.from_tables(dev, NE, D1D ** dim, nd, g["gather_map"], g["offsets"], g["indices"])
    fs = _Fes()
    fs.ne, fs.dof, fs.ndofs = NE, D1D ** dim, nd
    fs.GetElementRestriction = lambda d: Rs
    s = m.BilinearForm(dev, fs)
    s.AddDomainIntegrator(it)
    s.Assemble()
    for c in range(vdim):
        yc = yh[c::vdim] if byvdim else yh[c * nd:(c + 1) * nd]
        ys = dev.empty(nd)
        s.Mult(dev.to_device(comps[c]), ys)
        assert relerr(yc, host(ys)) < TOL
        if c == 0:
            assert relerr(yc, g["y"]) < TOL
    a.close(); s.close(); it.close(); R.close(); Rs.close()
