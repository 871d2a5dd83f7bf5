"""Pins the oracle's restatement of attribute-restricted domain integrators
(BilinearForm::AddDomainIntegrator(bfi, elem_marker); fem/bilinearform_ext.cpp:370-454,753-847)
against fixtures dumped from the unmodified reference (oracle/ref_driver.cpp do_dump_markers)."""
import numpy as np
import pytest

from conftest import load_golden, marker_golden_names, relerr
from oracle import oracle as orc

TOL = 1e-12


def marked_form(g):
    dim, D1D, Q1D, NE, nd = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "ndofs"))
    Dd = orc.diffusion_setup(dim, Q1D, 1, NE, g["W"], g["J"], np.array([1.0]))
    Dm = orc.mass_setup(Q1D ** dim, NE, g["W"], g["detJ"], np.array([1.0]))
    Dh = orc.mass_setup(Q1D ** dim, NE, g["W"], g["detJ"], np.array([0.5]))
    assert np.array_equal(Dd, g["pa_data_diffusion"])
    assert np.array_equal(Dm, g["pa_data_mass"])
    assert np.array_equal(Dh, g["pa_data_mass_half"])
    marked = [("diffusion", Dd, g["marker_diffusion"]), ("mass", Dm, g["marker_mass"])]
    plain = [("mass", Dh, None)]
    integ = plain + marked if int(g["marker_order"][0]) == 1 else marked + plain
    return orc.MarkedForm(dim, D1D, Q1D, NE, nd, g["gather_map"], g["offsets"], g["indices"], g["B"], g["G"],
                          g["elem_attributes"], integ)


@pytest.mark.parametrize("name", marker_golden_names())
def test_marked_operator_and_diagonal(name):
    g = load_golden(name)
    f = marked_form(g)
    assert len(np.unique(g["elem_attributes"])) >= 2
    assert relerr(f.mult(g["x"]), g["y"]) < TOL
    assert relerr(f.diag(), g["diag"]) < TOL
    assert relerr(f.mult_constrained(g["x"], g["ess_tdof_list"]), g["y_constrained"]) < TOL


def test_plain_first_diagonal_has_the_reference_zeros():
    """With the unrestricted integrator first, the later marked ones reset its diagonal on their
    unmarked elements (a property of the reference that a drop-in must reproduce)."""
    g = load_golden("mark_3d_p3_n3_plain_first")
    assert (g["diag"] == 0.0).any()
    assert (marked_form(g).diag() == 0.0).sum() == (g["diag"] == 0.0).sum()
