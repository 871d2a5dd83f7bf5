"""The generated shared-memory layout table of the even-odd kernel
(mfem_b200/csrc/pa_layout_table.h) is what tools/gen_layout_table.py produces, and it satisfies
what the kernel relies on: for even Q1D the A-slab stride residue is even (the Q outputs of a line
are stored / loaded as 128-bit words, PipeCfg::VEC1) and slab paddings keep 16-byte alignment."""
import os
import re
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def parse_table():
    src = open(os.path.join(ROOT, "mfem_b200", "csrc", "pa_layout_table.h")).read()
    body = src[src.index("kLayoutTable[2][8][3][32] = {"):]
    ent = [tuple(int(v) for v in m.groups()) for m in re.finditer(r"\{(\d+),(\d+),(\d+)\}", body)]
    assert len(ent) == 2 * 8 * 3 * 32
    return lambda kind, D, Q, NEB: ent[((kind * 8 + (D - 2)) * 3 + (Q - D)) * 32 + (NEB - 1)]


def test_even_q_entries_allow_128_bit_lines():
    get = parse_table()
    for kind in (0, 1):
        for D in range(2, 10):
            for Q in range(D, D + 3):
                for NEB in range(1, 33):
                    if NEB * Q * Q > 1024:
                        continue
                    order, ra, dpad = get(kind, D, Q, NEB)
                    assert 0 <= order <= 3 and 0 <= ra < 16 and dpad % 2 == 0
                    if Q % 2 == 0:
                        assert ra % 2 == 0, (kind, D, Q, NEB, ra)


@pytest.mark.parametrize("kind,D,Q,NEB", [(0, 5, 6, 3), (1, 5, 6, 2), (0, 4, 5, 4), (0, 8, 8, 1), (1, 3, 4, 10),
                                          (0, 2, 3, 20), (1, 9, 10, 1)])
def test_committed_table_is_what_the_generator_produces(kind, D, Q, NEB):
    import gen_layout_table as g
    order, ra = g.best_stage1(D, Q, NEB)
    dpad = g.best_dpad(6 if kind == 0 else 1, D, Q, NEB)
    assert parse_table()(kind, D, Q, NEB) == (order, ra, dpad)
