"""Pins the C restatement (oracle/oracle_pa.c) against the unmodified reference.

Every array compared here was dumped by oracle/ref_driver.cpp (the reference's own
code, compiled by oracle/Makefile) -- see oracle/make_golden.py.  Integer tables must
be bit-exact; floating point within 1e-12 relative (the tolerance BASELINE.json's
north_star states), and in practice ~1e-15.
"""
import numpy as np
import pytest

from conftest import relerr
from oracle import oracle as orc

TOL = 1e-12


def meta(g):
    return (int(g["dim"][0]), int(g["D1D"][0]), int(g["Q1D"][0]), int(g["NE"][0]),
            int(g["ndofs"][0]), bool(g["is_mass"][0]))


def test_restriction_tables_bit_exact(golden):
    dim, D1D, Q1D, NE, ndofs, _ = meta(golden)
    dof = D1D ** dim
    gm, off, ind = orc.restriction_build(NE, dof, ndofs, golden["elem_dof"], golden["dof_map"])
    assert np.array_equal(gm, golden["gather_map"])
    assert np.array_equal(off, golden["offsets"])
    assert np.array_equal(ind, golden["indices"])


def test_restriction_apply_bit_exact(golden):
    dim, D1D, Q1D, NE, ndofs, _ = meta(golden)
    dof = D1D ** dim
    xE = orc.restriction_mult(NE, dof, golden["gather_map"], golden["x"])
    assert np.array_equal(xE, golden["xE"])
    y = orc.restriction_mult_transpose(ndofs, dof, golden["offsets"], golden["indices"],
                                       golden["yE"])
    assert np.array_equal(y, golden["y_RtyE"])
    t = orc.restriction_abs_mult_transpose(ndofs, dof, golden["offsets"], golden["indices"],
                                           golden["xE"])
    assert np.array_equal(t, golden["absmultT_xE"])
    t2 = orc.restriction_mult_transpose(ndofs, dof, golden["offsets"], golden["indices"],
                                        golden["xE"], y=orc.randomize(ndofs, 3))
    assert np.array_equal(t2, golden["addmultT_xE_seed3"])


def test_randomize_matches_reference(golden):
    assert np.array_equal(orc.randomize(len(golden["x"]), 1), golden["x"])


def test_quadrature_and_basis(golden):
    dim, D1D, Q1D, NE, ndofs, _ = meta(golden)
    gll = bool(golden["is_gll"][0])
    qp, qw = (orc.gauss_lobatto if gll else orc.gauss_legendre)(Q1D)
    assert np.array_equal(qp, golden["qpts1d"])
    assert np.array_equal(qw, golden["qwts1d"])
    nodes, _ = orc.gauss_lobatto(D1D)
    assert np.array_equal(nodes, golden["dofpts1d"])
    B, G = orc.basis_tables(nodes, qp)
    assert np.array_equal(B, golden["B"])
    assert np.array_equal(G, golden["G"])
    W = orc.tensor_weights(dim, qw)
    assert np.array_equal(W, golden["W"])
    if gll:  # BP5: collocated, B is the identity up to 1 ulp (exact for p=7)
        assert np.abs(B.reshape(D1D, Q1D) - np.eye(Q1D)).max() <= 2.3e-16


def test_geometric_factors(golden):
    dim, D1D, Q1D, NE, ndofs, _ = meta(golden)
    J, detJ = orc.geometric_factors(dim, int(golden["node_D1D"][0]), Q1D, NE, golden["Bn"],
                                    golden["Gn"], golden["nodesE"])
    assert relerr(J, golden["J"]) < TOL
    assert relerr(detJ, golden["detJ"]) < TOL


def test_quadrature_data(golden):
    dim, D1D, Q1D, NE, ndofs, is_mass = meta(golden)
    if is_mass:
        D = orc.mass_setup(Q1D ** dim, NE, golden["W"], golden["detJ"], [1.0])
        assert np.array_equal(D, golden["pa_data"])
        return
    D = orc.diffusion_setup(dim, Q1D, 1, NE, golden["W"], golden["J"], [1.0])
    assert np.array_equal(D, golden["pa_data"])
    assert np.array_equal(D, golden["pa_data_seam"])
    cd = dim * (dim + 1) // 2
    Dm = orc.diffusion_setup(dim, Q1D, cd, NE, golden["W"], golden["J"], golden["coeff_matrix"])
    assert np.array_equal(Dm, golden["pa_data_matrix"])
    Dv = orc.diffusion_setup(dim, Q1D, dim, NE, golden["W"], golden["J"], golden["coeff_vector"])
    assert np.array_equal(Dv, golden["pa_data_vector"])


def test_element_apply_and_diag(golden):
    dim, D1D, Q1D, NE, ndofs, is_mass = meta(golden)
    if is_mass:
        yE = orc.mass_apply(dim, D1D, Q1D, NE, golden["B"], golden["pa_data"], golden["xE"])
        dE = orc.mass_diag(dim, D1D, Q1D, NE, golden["B"], golden["pa_data"])
    else:
        yE = orc.diffusion_apply(dim, D1D, Q1D, NE, golden["B"], golden["G"],
                                 golden["pa_data"], golden["xE"])
        dE = orc.diffusion_diag(dim, D1D, Q1D, NE, golden["B"], golden["G"], golden["pa_data"])
    assert relerr(yE, golden["yE"]) < TOL
    assert relerr(dE, golden["diagE"]) < TOL


def make_op(g):
    dim, D1D, Q1D, NE, ndofs, is_mass = meta(g)
    return orc.Operator(dim, D1D, Q1D, NE, ndofs, g["gather_map"], g["offsets"], g["indices"],
                        g["B"], g["G"], g["pa_data"], is_mass=is_mass, ess=g["ess_tdof_list"])


def test_operator_and_constrained(golden):
    op = make_op(golden)
    y = op.mult(golden["x"])
    assert relerr(y, golden["y"]) < TOL
    assert abs(np.linalg.norm(y) - golden["y_norml2"][0]) < TOL * golden["y_norml2"][0]
    assert relerr(op.mult_constrained(golden["x"]), golden["y_constrained"]) < TOL
    assert relerr(op.diag(), golden["diag"]) < TOL
    # homogeneous BC: FormLinearSystem leaves b untouched except b[ess] = 0
    B = op.eliminate_rhs(np.zeros(op.ndofs), golden["b"])
    assert relerr(B, golden["B_rhs"]) < TOL


def test_jacobi_and_pcg(golden):
    op = make_op(golden)
    dinv = op.jacobi(golden["diag"])
    z = np.zeros(op.ndofs) + dinv * golden["jacobi_in_seed5"]
    assert relerr(z, golden["jacobi_out"]) < TOL
    X, its, conv, fn, hist = op.pcg(golden["B_rhs"], dinv, rel_tol=float(golden["pcg_rtol"][0]))
    ref_its = int(golden["pcg_iters"][0])
    assert abs(its - ref_its) <= 1           # north_star: PCG within +-1 iteration
    assert conv == bool(golden["pcg_converged"][0])
    assert relerr(X, golden["X_sol"]) < 1e-6  # solver tolerance, not kernel parity
    n = min(len(hist), len(golden["pcg_history"]), 6)
    # the printed history has 6 significant digits (linalg/solvers.cpp:984)
    assert np.allclose(hist[:n], golden["pcg_history"][:n], rtol=2e-5)
