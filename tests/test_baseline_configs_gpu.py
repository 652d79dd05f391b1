"""GPU parity at the FULL size of BASELINE.json configs[3] (C4, uniform synthetic, N=128 fp32) and
configs[4] (C5, power-law synthetic, N=16 fp64) -- SURVEY.md 8(d): 10 000 sampled rows including
the longest against the oracle, plus a whole-result checksum from an independent O(nnz) formula:
    sum(C) = alpha * sum_j val[j] * rowsum(B)[col[j]] + beta * sum(C_in).
Inputs are bench.py's own (same generators, same seeds), so these are the matrices the bench times."""
import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from sextans_b200 import workloads as wl

pytestmark = pytest.mark.gpu
ALPHA, BETA = float(np.float32(0.85)), float(np.float32(-2.06))


def _check(M, K, N, rp, ci, v, B, Cin, split_tol, checksum_tol):
    dtype = v.dtype.type
    with sx.Engine(0) as eng:
        eng.upload_csr(M, K, rp, ci, v)
        C = Cin.copy()
        ns = eng.spmm(N, dtype(ALPHA), B, dtype(BETA), C)
        assert ns > 0 and eng.launches > 0
        split_rows = eng.info(sx.INFO_SPLIT_ROWS)
    lens = np.diff(rp)
    rng = np.random.default_rng(2024)
    rows = np.unique(np.concatenate([rng.integers(0, M, size=10_000), np.argsort(lens)[-16:], [0, M - 1]])).astype(np.int32)
    want = oracle.spmm_csr_rows(M, N, K, rp, ci, v, dtype(ALPHA), B, dtype(BETA), Cin, rows)
    got = np.ascontiguousarray(C.reshape(N, M).T[rows])
    short = lens[rows] <= 512                              # SX_OPT_SPLIT_ROW_NNZ: longer rows are summed in pieces
    assert np.array_equal(got[short].view(np.uint8), want[short].view(np.uint8)), "unsplit rows must be bit-exact"
    if (~short).any():
        scale = np.abs(want[~short]).max()
        assert np.abs(got[~short] - want[~short]).max() <= split_tol * scale
    assert (split_rows > 0) == bool((lens > 512).any())
    rel = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)
    assert rel[short].max() == 0.0
    # whole-result checksum
    Bsum = B.reshape(N, K).astype(np.float64).sum(axis=0)                      # row sums of B (B is column-major)
    total = ALPHA * float(np.dot(v.astype(np.float64), Bsum[ci])) + BETA * float(Cin.astype(np.float64).sum())
    mass = abs(ALPHA) * float(np.dot(np.abs(v).astype(np.float64), np.abs(B).reshape(N, K).astype(np.float64).sum(axis=0)[ci])) \
        + abs(BETA) * float(np.abs(Cin).astype(np.float64).sum())
    assert abs(float(C.astype(np.float64).sum()) - total) <= checksum_tol * mass
    return rows.size


def test_config4_uniform_full_size_sampled_rows_and_checksum():
    M = K = 1_000_000
    N = 128
    rp, ci, v = wl.uniform_csr(M, K, 20, 12345, np.float32)
    assert ci.size == 20_000_000
    B, Cin = wl.random_dense(M, K, N, 12345, np.float32)
    assert _check(M, K, N, rp, ci, v, B, Cin, split_tol=1e-5, checksum_tol=1e-6) > 9000


def test_config5_powerlaw_full_size_sampled_rows_and_checksum():
    M = K = 1_000_000
    N = 16
    rp, ci, v = wl.powerlaw_csr(M, K, 100_000_000, 12345, np.float64)
    assert ci.size == 100_000_000 and np.diff(rp).min() >= 1
    B, Cin = wl.random_dense(M, K, N, 12345, np.float64)
    assert _check(M, K, N, rp, ci, v, B, Cin, split_tol=1e-12, checksum_tol=1e-13) > 9000
