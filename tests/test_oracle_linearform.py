"""Pins the oracle's DomainLFIntegrator / VectorDomainLFIntegrator device-assembly restatement
(oracle_pa.c orc_domain_lf; fem/integ/lininteg_domain_kernels.hpp:76-297,
fem/linearform_ext.cpp:20-66) against fixtures dumped from the unmodified reference
(oracle/ref_driver.cpp do_dump_lf)."""
import numpy as np
import pytest

from conftest import lf_golden_names, load_golden, relerr
from oracle import oracle as orc

TOL = 1e-12


def scatter_vdofs(g, be):
    """LinearFormExtension::Assemble: lf = R^T b_e for a (dofs, vdim, NE) E-vector."""
    dim, D1D, NE, nd, vdim = (int(g[k][0]) for k in ("dim", "D1D", "NE", "ndofs", "vdim"))
    dof = D1D ** dim
    be = be.reshape(NE, vdim, dof)
    out = np.empty(nd * vdim)
    for c in range(vdim):
        comp = orc.restriction_mult_transpose(nd, dof, g["offsets"], g["indices"],
                                              np.ascontiguousarray(be[:, c, :]).ravel())
        if int(g["byvdim"][0]) and vdim > 1:
            out[c::vdim] = comp
        else:
            out[c * nd:(c + 1) * nd] = comp
    return out


@pytest.mark.parametrize("name", lf_golden_names())
def test_domain_lf_device_assembly(name):
    g = load_golden(name)
    dim, D1D, Q1D, NE, vdim = (int(g[k][0]) for k in ("dim", "D1D", "Q1D", "NE", "vdim"))
    be = orc.domain_lf(dim, D1D, Q1D, vdim, NE, g["B"], g["W"], g["detJ"], g["coeff"])
    assert relerr(be, g["b_e"]) < TOL
    bm = orc.domain_lf(dim, D1D, Q1D, vdim, NE, g["B"], g["W"], g["detJ"], g["coeff"], g["markers"])
    assert relerr(bm, g["b_e_marked"]) < TOL
    assert 0 < np.count_nonzero(g["markers"]) < NE
    bc = orc.domain_lf(dim, D1D, Q1D, vdim, NE, g["B"], g["W"], g["detJ"], 1.5 + np.arange(vdim) if vdim > 1
                       else np.array([3.0]))
    assert relerr(bc, g["b_e_const"]) < TOL
    b = scatter_vdofs(g, be)
    assert relerr(b, g["b_fast"]) < TOL
    assert relerr(b, g["b_legacy"]) < 1e-11   # the element-loop assembly sums in another order
