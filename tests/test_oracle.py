"""CPU tests: the oracle restatement is pinned to the reference.

Two anchors: the committed golden fixtures (generated from the reference's own
cpu_spmm_CSR / loader by tests/golden/make_golden.py) and, when the compiled
reference header is present (oracle/_ref), a direct comparison.
"""
import os

import numpy as np
import pytest

import oracle
from helpers import (GOLDEN, SMALL_MTX, SUITESPARSE, mtx_path, perturbed_inputs, random_csr,
                     random_dense, sha)

needs_ref = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built")


@pytest.mark.parametrize("name", SMALL_MTX)
def test_loader_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, "loader_small.npz"))
    M, K, nnz, rp, ci, v, code = oracle.load_mtx(mtx_path(name), np.float32)
    assert [M, K, nnz] == g[name + "_shape"].tolist()
    assert np.array_equal(rp, g[name + "_rowptr"])
    assert np.array_equal(ci, g[name + "_colidx"])
    assert np.array_equal(v.view(np.uint32), g[name + "_val"].view(np.uint32))
    # the f64 loader sees the same structure on these files
    M2, K2, nnz2, rp2, ci2, v2, _ = oracle.load_mtx(mtx_path(name), np.float64)
    assert (M2, K2, nnz2) == (M, K, nnz) and np.array_equal(rp2, rp) and np.array_equal(ci2, ci)
    assert np.allclose(v2, v, rtol=1e-7, atol=0)


def test_loader_semantics_spelled_out():
    # general_real: 14 stored, two exact +0 dropped, -0.0 kept, duplicate kept twice
    M, K, nnz, rp, ci, v, code = oracle.load_mtx(mtx_path("general_real"), np.float32)
    assert (M, K, nnz, code) == (7, 5, 12, "MCRG")
    assert np.diff(rp).tolist() == [3, 1, 2, 0, 2, 2, 2]
    assert ci[rp[0]:rp[1]].tolist() == [0, 2, 3] and v[rp[0]:rp[1]].tolist() == [1.5, -0.0, 2.5]
    assert np.signbit(v[rp[0] + 1])  # the -0.0 entry survives, +0.0 entries do not
    assert ci[rp[4]:rp[5]].tolist() == [1, 1] and sorted(v[rp[4]:rp[5]].tolist()) == [0.875, 4.125]
    # symmetric_real: off-diagonals mirrored, the explicit zero dropped
    M, K, nnz, rp, ci, v, code = oracle.load_mtx(mtx_path("symmetric_real"), np.float32)
    assert (M, K, nnz, code) == (6, 6, 3 + 2 * 5, "MCRS")
    # skew: read as general
    M, K, nnz, rp, ci, v, code = oracle.load_mtx(mtx_path("skew"), np.float32)
    assert (nnz, code) == (4, "MCRK") and (v > 0).sum() == 3
    # integer: parsed with %f, 16777217 rounds to 16777216 in fp32 but not in fp64
    *_, v32, _ = oracle.load_mtx(mtx_path("integer_general"), np.float32)
    *_, v64, _ = oracle.load_mtx(mtx_path("integer_general"), np.float64)
    assert 16777216.0 in v32.tolist() and 16777217.0 in v64.tolist()


def test_loader_errors(tmp_path):
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%NotMatrixMarket matrix coordinate real general\n1 1 1\n1 1 1\n")
    with pytest.raises(RuntimeError, match="banner"):
        oracle.load_mtx(str(bad))
    arr = tmp_path / "arr.mtx"
    arr.write_text("%%MatrixMarket matrix array real general\n1 1 1\n1.0\n")
    with pytest.raises(RuntimeError, match="coordinate"):
        oracle.load_mtx(str(arr))
    cpx = tmp_path / "cpx.mtx"
    cpx.write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 2.0\n")
    with pytest.raises(RuntimeError, match="complex"):
        oracle.load_mtx(str(cpx))
    with pytest.raises(RuntimeError, match="open"):
        oracle.load_mtx(str(tmp_path / "missing.mtx"))
    zero = tmp_path / "zero.mtx"
    zero.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 3.0\n")
    with pytest.raises(RuntimeError, match="index"):
        oracle.load_mtx(str(zero))


@pytest.mark.parametrize("name", SUITESPARSE)
def test_suitesparse_loader_and_default_run_golden(name, golden):
    g = golden["suitesparse"][name]
    M, K, nnz, rp, ci, v, code = oracle.load_mtx(mtx_path(name), np.float32)
    assert (M, K, nnz, code) == (g["M"], g["K"], g["nnz"], "MCPS")
    assert sha(rp) == g["rowptr_sha256"] and sha(ci) == g["colidx_sha256"]
    assert sha(v) == g["val_sha256"]
    lens = np.diff(rp)
    assert (lens.min(), lens.max()) == (g["row_len_min"], g["row_len_max"])
    for run in g["runs"]:
        N = run["N"]
        if run["kind"] == "default":
            B, C = oracle.init_dense(M, K, N, np.float32)
            val = v
        else:
            val, B, C = perturbed_inputs(M, K, N, nnz, np.float32)
        oracle.spmm_csr(M, N, K, rp, ci, val, run["alpha"], B, run["beta"], C)
        assert sha(C) == run["C_sha256"], (name, run["kind"], N)
        assert float(C[0]) == run["C0"] and float(C[-1]) == run["C_last"]
        if run["kind"] == "default":
            # closed form on the shipped inputs (pattern A, B = 1): SURVEY.md 8(c)
            m = np.arange(M, dtype=np.float64)[None, :]
            n = np.arange(N, dtype=np.float64)[:, None]
            cin = ((m + 1) * (n + 1) / M / N).astype(np.float32)
            closed = (np.float32(0.85) * lens.astype(np.float32)[None, :]
                      + np.float32(-2.06) * cin).astype(np.float32)
            assert np.array_equal(closed.ravel(), C)


def test_small_cases_golden():
    g = np.load(os.path.join(GOLDEN, "spmm_small.npz"))
    for tag in "abcdef":
        M, K, N = g[tag + "_dims"].tolist()
        alpha, beta = g[tag + "_ab"].tolist()
        C = g[tag + "_Cin"].copy()
        oracle.spmm_csr(M, N, K, g[tag + "_rowptr"], g[tag + "_colidx"], g[tag + "_val"],
                        alpha, g[tag + "_B"], beta, C)
        assert np.array_equal(C.view(np.uint32), g[tag + "_C"].view(np.uint32)), tag
        # the seeded generators still produce the committed inputs
        spec = {"a": (6, None), "b": (3, 60), "c": (4, None), "d": (20, None), "e": (3, None),
                "f": (0, None)}[tag]
        i = "abcdef".index(tag)
        rp, ci, v = random_csr(M, K, spec[0], 100 + i, np.float32, long_row=spec[1])
        assert np.array_equal(rp, g[tag + "_rowptr"]) and np.array_equal(v, g[tag + "_val"])


def test_threads_and_row_sample_are_bitwise_identical():
    M, K, N = 300, 200, 16
    for dtype in (np.float32, np.float64):
        rp, ci, v = random_csr(M, K, 12, 7, dtype, long_row=150)
        B, Cin = random_dense(M, K, N, 7, dtype)
        C1 = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
        C4 = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy(), threads=4)
        assert np.array_equal(C1, C4)
        rows = np.array([0, 5, 299, 17, 5], dtype=np.int32)
        S = oracle.spmm_csr_rows(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin, rows)
        assert np.array_equal(S, C1.reshape(N, M).T[rows])


def test_f64_oracle_agrees_with_f32_reference_within_rounding(golden):
    name = "nasa4704"
    M, K, nnz, rp, ci, v, _ = oracle.load_mtx(mtx_path(name), np.float64)
    B, C = oracle.init_dense(M, K, 16, np.float64)
    oracle.spmm_csr(M, 16, K, rp, ci, v, float(np.float32(0.85)), B, float(np.float32(-2.06)), C)
    run = [r for r in golden["suitesparse"][name]["runs"] if r["kind"] == "default" and r["N"] == 16][0]
    assert abs(C[0] - run["C0"]) / abs(run["C0"]) < 1e-6
    assert abs(C.sum() - run["sum"]) / abs(run["sum"]) < 1e-6


def test_verify_criterion():
    cpu = np.array([1.0, 2.0, 0.0, -3.0], dtype=np.float32)
    dev = np.array([1.00005, 2.001, 0.0, -3.0], dtype=np.float32)
    n, pct, ok = oracle.verify_f32(cpu, dev, 2, 2)
    assert n == 1 and pct == 25.0 and not ok
    n, pct, ok = oracle.verify_f32(cpu, cpu, 2, 2)
    assert n == 0 and ok


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_restatement_equals_compiled_reference(seed):
    M, K, N = 257, 193, 8 * seed
    rp, ci, v = random_csr(M, K, 9, seed, np.float32, long_row=120)
    B, Cin = random_dense(M, K, N, seed, np.float32)
    a, b = np.float32(0.85), np.float32(-2.06)
    C_ref = Cin.copy()
    oracle.ref_spmm_csr(M, N, K, rp, ci, v, a, B, b, C_ref)
    C_or = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
    assert np.array_equal(C_ref.view(np.uint32), C_or.view(np.uint32))


@needs_ref
@pytest.mark.parametrize("name", SMALL_MTX + ["nasa4704"])
def test_loader_equals_compiled_reference(name, capfd):
    ref = oracle.ref_load_csr(mtx_path(name))
    mine = oracle.load_mtx(mtx_path(name), np.float32)
    assert ref[:3] == mine[:3]
    for a, b in zip(ref[3:], mine[3:6]):
        assert np.array_equal(a, b)
