"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with
the keys the driver reads, on a bounded sample."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.parametrize("bp", [3, 1, 5, 4])
def test_reference_arm_prints_the_contract_line(bp):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bp", str(bp),
           "--steps", "1", "--warmup", "1", "--cpu-n", "4", "--iters", "8"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GDOF/s" and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.decode().strip() == ""


def test_parity_check_maps_the_reference_numbering_onto_partitions():
    """bench.py's parity_check draws x = Vector::Randomize(1) in the REFERENCE numbering and hands
    every rank its true dofs through the lattice tables: the map must be a bijection between the
    serial dofs and the union of the ranks' true dofs, and the identity for a one-rank partition."""
    import sys as _sys
    import numpy as np
    _sys.path.insert(0, ROOT)
    import bench
    import mfem_b200 as m
    n, p = 4, 3
    s_of_lat, nser = bench.serial_dof_of_lattice(m, np, n, p)
    assert nser == (n * p + 1) ** 3 and sorted(s_of_lat) == list(range(nser))
    one = m.FiniteElementSpace(m.Mesh(3, n, n, n, True, part=(1, 1, 1), rank=0), m.H1_FECollection(p, 3))
    assert np.array_equal(s_of_lat[one.halo_tables["ldof_lattice"]], np.arange(nser))  # same numbering
    seen = np.zeros(nser, int)
    for r in range(8):
        f = m.FiniteElementSpace(m.Mesh(3, n, n, n, True, part=(2, 2, 2), rank=r), m.H1_FECollection(p, 3))
        t2s = s_of_lat[f.halo_tables["ldof_lattice"][f.halo_tables["ltdof_ldof"]]]
        seen[t2s] += 1
    assert np.all(seen == 1)  # every reference dof is a true dof of exactly one rank


def test_vector_randomize_matches_the_reference_generator():
    """b2f_vector_randomize_h mirrors Vector::Randomize (srand(seed); rand()/(RAND_MAX+1.0)): checked
    against the oracle's restatement and against a fixture vector drawn by the reference itself."""
    import ctypes as C
    import numpy as np
    from conftest import load_golden
    from mfem_b200.capi import call
    from oracle import oracle as orc
    g = load_golden("diff3d_p4_n3")
    x = np.empty(g["x"].size)
    call("b2f_vector_randomize_h", x.ctypes.data_as(C.c_void_p), x.size, 1)
    assert np.array_equal(x, g["x"]) and np.array_equal(x, orc.randomize(x.size, 1))
