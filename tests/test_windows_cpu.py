"""CPU tests of the column-window split (sx_split_col_windows, host only) and of the claim
the window passes rest on: walking the windows in ascending order, each pass continuing
from the running sum the previous one left, performs exactly the oracle's chain of
additions for rows stored in ascending column order."""
import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import random_csr, random_dense


def model_split(M, K, rp, ci, W):
    nwin = max(1, (K + W - 1) // W)
    segs = [[[j for j in range(rp[r], rp[r + 1]) if ci[j] // W == w] for r in range(M)] for w in range(nwin)]
    return nwin, segs


@pytest.mark.parametrize("M,K,avg,W,sort", [(50, 100, 7, 32, True), (50, 100, 7, 100, True), (1, 5, 3, 2, True),
                                            (300, 1000, 20, 256, False), (0, 10, 0, 4, True), (20, 30, 0, 7, True),
                                            (64, 4097, 30, 4096, True), (10, 7, 3, 1, True)])
def test_split_matches_the_definition(M, K, avg, W, sort):
    rp, ci, _ = random_csr(M, K, avg, M + K + W, np.float32, sort=sort)
    wrp, base, order, asc = sx.split_col_windows(M, K, rp, ci, W)
    nwin, segs = model_split(M, K, rp, ci, W)
    assert wrp.shape == (nwin, M + 1) and base.size == nwin + 1
    assert base[0] == 0 and base[-1] == rp[M]
    assert sorted(order.tolist()) == list(range(int(rp[M])))
    for w in range(nwin):
        assert wrp[w, 0] == 0 and wrp[w, M] == base[w + 1] - base[w]
        for r in range(M):
            assert order[base[w] + wrp[w, r]: base[w] + wrp[w, r + 1]].tolist() == segs[w][r]
    if rp[M] > 0:
        assert asc == bool(np.all([np.all(np.diff(ci[rp[r]:rp[r + 1]]) >= 0) for r in range(M)]))


def test_split_is_independent_of_the_thread_count(monkeypatch):
    M, K, W = 20000, 50000, 8192
    rp, ci, _ = random_csr(M, K, 20, 3, np.float32)   # > 2^18 nonzeros: the threaded path
    assert rp[M] > (1 << 18)
    outs = []
    for t in ("1", "5", "16"):
        monkeypatch.setenv("SX_LOADER_THREADS", t)
        outs.append(sx.split_col_windows(M, K, rp, ci, W))
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(o[:3], outs[0][:3])) and o[3] == outs[0][3]
    wrp, base, order, asc = outs[0]
    assert asc
    # every window's entries lie in its column range, rows in order, stored order inside
    for w in range(wrp.shape[0]):
        seg = order[base[w]:base[w + 1]]
        assert np.all(ci[seg] // W == w)
        rows = np.repeat(np.arange(M), np.diff(wrp[w]))
        assert np.all(np.searchsorted(rp, seg, side="right") - 1 == rows)
        same_row = rows[1:] == rows[:-1]
        assert np.all(seg[1:][same_row] > seg[:-1][same_row])


def test_split_error_paths():
    rp = np.array([0, 2], dtype=np.int32)
    ci = np.array([0, 9], dtype=np.int32)
    with pytest.raises(sx.SextansError, match="out of range"):
        sx.split_col_windows(1, 5, rp, ci, 2)
    with pytest.raises(sx.SextansError, match="bad argument"):
        sx.split_col_windows(1, 10, rp, ci, 0)
    with pytest.raises(sx.SextansError, match="4096 windows"):
        sx.split_col_windows(1, 10_000_000, rp, ci, 2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_window_passes_continue_the_oracles_chain_of_additions(dtype):
    """Model of the passes in numpy (separately rounded * and +, like strict mode): first
    pass from 0, later passes from the running sum, epilogue after the last one == oracle."""
    M, K, N, W = 60, 500, 8, 128
    rp, ci, v = random_csr(M, K, 25, 9, dtype)
    B, Cin = random_dense(M, K, N, 9, dtype)
    a, b = dtype(0.85), dtype(-2.06)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
    wrp, base, order, asc = sx.split_col_windows(M, K, rp, ci, W)
    assert asc
    Bm = B.reshape(N, K)
    P = np.zeros((M, N), dtype=dtype)
    for w in range(wrp.shape[0]):
        for r in range(M):
            acc = P[r].copy()
            for j in order[base[w] + wrp[w, r]: base[w] + wrp[w, r + 1]]:
                acc = (acc + (v[j] * Bm[:, ci[j]]).astype(dtype)).astype(dtype)
            P[r] = acc
    C = ((a * P).astype(dtype) + (b * Cin.reshape(N, M).T).astype(dtype)).astype(dtype)
    assert np.array_equal(np.ascontiguousarray(C.T).ravel().view(np.uint8), ref.view(np.uint8))
