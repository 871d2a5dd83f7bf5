"""CPU: the numpy restatement of the reference's batched LOR assembly (oracle/oracle_lor.py) against
the fixtures generated from the unmodified reference (tests/golden/lor_*.npz, oracle/make_golden.py,
oracle/ref_driver.cpp do_dump_lor)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle_lor as L

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "lor_*.npz")))


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def test_fixtures_present():
    assert len(NAMES) >= 14


@pytest.mark.parametrize("name", NAMES)
def test_reference_batched_equals_reference_legacy(name):
    """What the reference's own unit test checks (tests/unit/fem/test_lor_batched.cpp:95-101)."""
    g = load(name)
    scale = np.abs(g["data_bc"]).max()
    assert g["legacy_minus_batched_maxnorm"][0] < 1e-12 * scale


@pytest.mark.parametrize("name", NAMES)
def test_oracle_lor_against_reference(name):
    g = load(name)
    p, ne, nd, dim = int(g["p"][0]), int(g["NE"][0]), int(g["ndofs"][0]), int(g["dim"][0])
    n1 = p + 1
    X = L.lor_vertex_coordinates(int(g["node_order"][0]), g["nodes_e"], ne, p, dim)
    assert np.abs(X - g["X_vert"]).max() < 1e-14 * max(1.0, np.abs(g["X_vert"]).max())
    assert np.array_equal(L.sparse_mapping(p, dim).reshape(-1), g["sparse_mapping"])
    Vref = g["sparse_ij"].reshape(ne, n1 ** dim, 3 ** dim)
    if ne * p ** dim <= 1000:  # python loops: the larger cases check the CSR stage only
        asm = L.assemble_sparse_ij if dim == 3 else L.assemble_sparse_ij_2d
        V = asm(p, ne, g["X_vert"], g["c_mass"], g["c_diff"])
        assert np.abs(V - Vref).max() < 1e-13 * np.abs(Vref).max()
    I, J, A = L.fill_csr(p, ne, nd, g["gather_map"], g["offsets"], g["indices"], Vref, dim)
    assert np.array_equal(I, g["I"]) and np.array_equal(J, g["J"])      # bit-exact index work
    assert np.abs(A - g["data"]).max() <= 1e-15 * np.abs(A).max()
    Ab = L.eliminate_bc(nd, I, J, A, g["ess_tdof_list"])
    assert np.array_equal(g["I_bc"], I) and np.array_equal(g["J_bc"], J)
    assert np.abs(Ab - g["data_bc"]).max() <= 1e-15 * np.abs(A).max()
    y = L.spmv(nd, I, J, Ab, g["x"])
    assert np.abs(y - g["y_bc"]).max() < 1e-13 * np.abs(g["y_bc"]).max()
    # the legacy (refined mesh, full assembly) matrix of the same form
    if nd <= 2500:
        Dl = L.to_dense(nd, g["I_legacy"], g["J_legacy"], g["data_legacy"])
        Db = L.to_dense(nd, I, J, Ab)
        assert np.abs(Dl - Db).max() < 1e-12 * np.abs(Db).max()
