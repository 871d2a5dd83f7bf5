"""Pins the non-symmetric-coefficient branches of the C oracle (dim*dim quadrature-data
components: setup, apply, diagonal) against fixtures dumped from the unmodified reference
(DiffusionIntegrator(MatrixFunctionCoefficient), oracle/ref_driver.cpp do_dump_nonsym)."""
import numpy as np
import pytest

from conftest import load_golden, nonsym_golden_names, relerr
from oracle import oracle as orc

TOL = 1e-12


@pytest.mark.parametrize("name", nonsym_golden_names())
def test_nonsymmetric_setup_apply_diag(name):
    g = load_golden(name)
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    assert int(g["pa_symmetric"][0]) == 0
    D = orc.diffusion_setup(dim, Q1D, dim * dim, NE, g["W"], g["J"], g["coeff_full"])
    assert np.array_equal(D, g["pa_data"])
    yE = orc.diffusion_apply(dim, D1D, Q1D, NE, g["B"], g["G"], D, g["xE"], symmetric=0)
    assert relerr(yE, g["yE"]) < TOL
    dE = orc.diffusion_diag(dim, D1D, Q1D, NE, g["B"], g["G"], D, symmetric=0)
    assert relerr(dE, g["diagE"]) < TOL
    op = orc.Operator(dim, D1D, Q1D, NE, nd, g["gather_map"], g["offsets"], g["indices"], g["B"], g["G"], D,
                      symmetric=0)
    assert relerr(op.mult(g["x"]), g["y"]) < TOL
    assert relerr(op.diag(), g["diag"]) < TOL
