"""Pins the oracle's OperatorChebyshevSmoother / PowerMethod restatement (oracle_pa.c) against
fixtures dumped from the unmodified reference (oracle/ref_driver.cpp do_dump_cheb:
linalg/solvers.cpp:453-658, linalg/operator.cpp:886-945), and checks the product's host-side
coefficient routine (no device needed) against both."""
import numpy as np
import pytest

from conftest import chebyshev_golden_names, load_golden, relerr
from oracle import oracle as orc

TOL = 1e-12


def oracle_operator(g):
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    return orc.Operator(dim, D1D, Q1D, NE, nd, g["gather_map"], g["offsets"], g["indices"], g["B"], g["G"],
                        g["pa_data"], is_mass=bool(int(g["is_mass"][0])), ess=g["ess_tdof_list"])


@pytest.mark.parametrize("name", chebyshev_golden_names())
def test_power_method_and_smoother_action(name):
    g = load_golden(name)
    op = oracle_operator(g)
    assert relerr(op.diag(), g["diag"]) < TOL
    dinv = op.jacobi(g["diag"], 1.0)
    n = len(g["diag"])
    start = orc.randomize(n, int(g["power_seed"][0]))
    assert np.array_equal(start, g["power_start"])
    lam, v = op.power_method(dinv, start, int(g["power_iterations"][0]), float(g["power_tolerance"][0]))
    assert abs(lam - g["max_eig_estimate"][0]) < 1e-12 * abs(lam)
    assert relerr(v, g["power_vector"]) < 1e-10
    x = g["cheb_in_seed11"]
    for order in range(1, 6):
        y = op.chebyshev_mult(dinv, order, float(g["max_eig_estimate"][0]), x)
        assert relerr(y, g["cheb_out_order%d" % order]) < TOL, order


@pytest.mark.parametrize("name", chebyshev_golden_names())
def test_cg_with_chebyshev_preconditioner(name):
    g = load_golden(name)
    op = oracle_operator(g)
    dinv = op.jacobi(g["diag"], 1.0)
    order = int(g["pcg_order"][0])
    x, its, conv, fn, hist = op.pcg_chebyshev(g["B_rhs"], dinv, order, float(g["max_eig_estimate"][0]),
                                              rel_tol=float(g["pcg_rtol"][0]))
    assert its == int(g["pcg_iters"][0]) and conv == bool(g["pcg_converged"][0])
    assert relerr(x, g["X_sol"]) < 1e-9
    ref = g["pcg_history"]
    assert len(hist) == len(ref)
    # the printed history has 6 significant digits
    assert np.allclose(hist, ref, rtol=2e-5, atol=0)


def test_order_outside_1_to_5_is_rejected():
    with pytest.raises(ValueError):
        orc.chebyshev_coeffs(6, 2.0)
    with pytest.raises(ValueError):
        orc.chebyshev_coeffs(0, 2.0)


@pytest.mark.parametrize("lam", [0.7, 2.0, 3.9])
def test_product_coefficients_match_oracle_and_residual_polynomial(lam):
    """b2f_chebyshev_coefficients_h is host-only.  Besides equality with the oracle: the
    coefficients define q with 1 - t q(t) = T_k((theta - t)/delta) / T_k(theta/delta)."""
    import ctypes as C
    from mfem_b200 import capi
    theta, delta = 0.75 * lam, 0.45 * lam
    for order in range(1, 6):
        c = np.zeros(5)
        rc = capi.lib().b2f_chebyshev_coefficients_h(order, C.c_double(lam), c.ctypes.data_as(capi.c_double_p))
        assert rc == 0
        assert np.allclose(c[:order], orc.chebyshev_coeffs(order, lam), rtol=1e-14, atol=0)
        t = np.linspace(0.3 * lam, 1.2 * lam, 41)
        q = sum(c[k] * t ** k for k in range(order))
        cheb = np.polynomial.chebyshev.Chebyshev.basis(order)
        assert np.allclose(1 - t * q, cheb((theta - t) / delta) / cheb(theta / delta), rtol=0, atol=1e-11)
    c = np.zeros(8)
    assert capi.lib().b2f_chebyshev_coefficients_h(6, C.c_double(lam), c.ctypes.data_as(capi.c_double_p)) != 0
