"""Property test of the Matrix-Market loader: for random small
coordinate files -- every field/symmetry combination the reference accepts, explicit zeros,
negative zeros, exponents, unordered entries -- the product loader (sx_load_mtx_*), the
oracle restatement and, when built, the reference's own read_suitsparse_matrix + CSC_2_CSR
(oracle/_ref) produce the same CSR, entry for entry.  (row, column) pairs are unique inside
a file: the reference orders exact duplicates with an unstable qsort, SURVEY.md appendix B.)"""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
import sextans_b200 as sx

VALUES = st.sampled_from(["1", "-2.5", "0", "-0", "0.0", "-0.0", "1e-3", "-4E2", "3.25e+1", "7", ".5", "1e-40",
                          "123456.789", "-1e30"])


@st.composite
def mtx_files(draw):
    field = draw(st.sampled_from(["real", "integer", "pattern"]))
    symmetry = draw(st.sampled_from(["general", "symmetric", "skew-symmetric"]))
    M = draw(st.integers(1, 12))
    K = M if symmetry != "general" else draw(st.integers(1, 12))
    cells = [(r, c) for r in range(1, M + 1) for c in range(1, K + 1) if symmetry == "general" or c <= r]
    chosen = draw(st.lists(st.sampled_from(cells), unique=True, max_size=min(len(cells), 30)))
    lines = [f"%%MatrixMarket matrix coordinate {field} {symmetry}", "% generated", f"{M} {K} {len(chosen)}"]
    for r, c in chosen:
        if field == "pattern":
            lines.append(f"{r} {c}")
        elif field == "integer":
            lines.append(f"{r} {c} {draw(st.integers(-5, 5))}")
        else:
            lines.append(f"{r}  {c}\t{draw(VALUES)}")
    return "\n".join(lines) + "\n"


def same(a, b):
    return (a[:3] == b[:3] and all(np.array_equal(x, y) for x, y in zip(a[3:5], b[3:5]))
            and np.array_equal(a[5].view(np.uint8), b[5].view(np.uint8)))


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=mtx_files())
def test_three_loaders_agree(tmp_path, text):
    p = os.path.join(tmp_path, "m.mtx")
    with open(p, "w") as f:
        f.write(text)
    for dtype in (np.float32, np.float64):
        mine = sx.load_mtx(p, dtype)
        port = oracle.load_mtx(p, dtype)
        assert same(mine, port[:6]), text
    if oracle.ref() is not None:
        ref = oracle.ref_load_csr(p)
        assert same(sx.load_mtx(p, np.float32), ref), text
