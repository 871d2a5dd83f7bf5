"""Pins the vector (vdim = 3, CEED BP2/4/6) restatement in oracle/oracle.py against fixtures
dumped from the unmodified reference (VectorDiffusionIntegrator / VectorMassIntegrator on a
vdim-3 H1 space, byVDIM and byNODES orderings; oracle/ref_driver.cpp do_dump_vector)."""
import numpy as np
import pytest

from conftest import load_golden, relerr, vector_golden_names
from oracle import oracle as orc

TOL = 1e-12


def build(g):
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    is_mass = bool(g["is_mass"][0])
    if is_mass:
        D = orc.mass_setup(Q1D ** dim, NE, g["W"], g["detJ"], [1.0])
    else:
        D = orc.diffusion_setup(dim, Q1D, 1, NE, g["W"], g["J"], [1.0])
    sop = orc.Operator(dim, D1D, Q1D, NE, nd, g["gather_map"], g["offsets"], g["indices"], g["B"], g["G"], D,
                       is_mass=is_mass)
    return sop, orc.VectorOperator(sop, int(g["vdim"][0]), int(g["byvdim"][0]), g["ess_tdof_list"])


@pytest.mark.parametrize("name", vector_golden_names())
def test_vector_operator_matches_reference(name):
    g = load_golden(name)
    sop, vop = build(g)
    dim, D1D, NE = int(g["dim"][0]), int(g["D1D"][0]), int(g["NE"][0])
    assert np.array_equal(vop.to_evector(g["x"], NE, D1D ** dim, g["gather_map"]), g["xE"])
    assert relerr(vop.mult(g["x"]), g["y"]) < TOL
    assert relerr(vop.mult_constrained(g["x"]), g["y_constrained"]) < TOL
    assert relerr(vop.diag(), g["diag"]) < TOL
    dinv = 1.0 / vop.diag()
    dinv[g["ess_tdof_list"]] = 1.0
    X, its, conv = vop.pcg(g["B_rhs"], dinv, rel_tol=float(g["pcg_rtol"][0]))
    assert conv and abs(its - int(g["pcg_iters"][0])) <= 1
    assert relerr(X, g["X_sol"]) < 1e-6
