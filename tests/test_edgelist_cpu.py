"""Host-side plan of the edge-list kernel (sx_plan_edge_lists; no GPU): the blocks tile the rows,
every block's runs list its distinct columns in ascending order, every nonzero's 16-bit local
column names its column (row-aligned streams: rows start at multiples of 8 entries and are padded to
multiples of 8), and no block exceeds the shared-memory budget."""
import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import mtx_path, random_csr


def check_plan(M, K, rp, ci, row_bytes, elem, budget, blocks, cols, lcol, total, max_smem, prow, rows=32, target=0):
    lens = np.diff(rp)
    assert prow[0] == 0 and np.array_equal(np.diff(prow), (lens + 7) & ~7)      # rows start at multiples of 8, padded to 8
    assert blocks.shape[1] == 8 and lcol.size == prow[M] and cols.size % 4 == 0
    nxt = 0
    tot = 0
    for b in blocks:
        r0, nr, pb, pe, c0, ncols, _, smem = (int(x) for x in b)
        assert r0 == nxt and 1 <= nr <= rows and (r0 // 4096) == ((r0 + nr - 1) // 4096)
        jb, je = int(rp[r0]), int(rp[r0 + nr])
        if target and nr > 1:                 # never further from the target than without its last row
            assert (je - jb) - target <= target - (rp[r0 + nr - 1] - jb)
        assert pb == prow[r0] and pe == prow[r0 + nr] and c0 % 4 == 0
        mine = cols[c0:c0 + ncols]
        assert mine.tolist() == sorted(set(ci[jb:je].tolist()))
        for r in range(r0, r0 + nr):          # entry k of row r at prow[r] + k names its column; pad entries are 0
            n = int(lens[r])
            assert np.array_equal(mine[lcol[prow[r]:prow[r] + n]], ci[rp[r]:rp[r] + n])
            assert not lcol[prow[r] + n:prow[r + 1]].any()
        pad = cols[c0 + ncols:c0 + ((ncols + 3) & ~3)]
        assert ncols == 0 or np.all(pad == mine[-1])                 # pad entries name a real column
        assert smem == ncols * row_bytes + (pe - pb) * (elem + 2) + ((ncols + 3) & ~3) * 4 + 2 * ((nr + 4) & ~3) * 4
        assert smem <= budget and smem <= max_smem
        nxt = r0 + nr
        tot += ncols
    assert nxt == M and tot == total


@pytest.mark.parametrize("name,row_bytes,elem,distinct", [("nasa4704", 128, 8, 20897), ("pcrystk02", 64, 4, 138342)])
def test_suitesparse_plans(name, row_bytes, elem, distinct):
    """BASELINE configs[1] (N=16 fp64) and configs[2] (N=16 fp32) at four blocks per SM."""
    M, K, nnz, rp, ci, v, _ = oracle.load_mtx(mtx_path(name), np.float32)
    blocks, runs, lcol, total, max_smem, prow = sx.plan_edge_lists(M, K, rp, ci, row_bytes, elem, 56192)
    assert len(blocks) == sum((min(M, g + 4096) - g + 31) // 32 for g in range(0, M, 4096)) and total <= distinct + 400   # no block had to be cut
    # cut by nonzeros into ~148 blocks (what the engine does with a small matrix on 148 SMs)
    target = -(-nnz // 148)
    bal = sx.plan_edge_lists(M, K, rp, ci, row_bytes, elem, 115000, max_rows=128, nnz_target=target)
    check_plan(M, K, rp, ci, row_bytes, elem, 115000, *bal, rows=128, target=target)
    sizes = rp[bal[0][:, 0] + bal[0][:, 1]] - rp[bal[0][:, 0]]
    assert 140 <= len(sizes) <= 160 and sizes.max() <= 1.25 * target
    check_plan(M, K, rp, ci, row_bytes, elem, 56192, blocks, runs, lcol, total, max_smem, prow)
    assert total * 2 <= nnz                                          # a staged B row serves >= 2 nonzeros
    # a budget that forces cuts
    small = sx.plan_edge_lists(M, K, rp, ci, row_bytes, elem, 12000)
    assert len(small[0]) > len(blocks) and small[3] >= total - 400
    check_plan(M, K, rp, ci, row_bytes, elem, 12000, *small)


@pytest.mark.parametrize("seed", range(4))
def test_random_matrices_with_empty_rows_and_unsorted_columns(seed):
    rng = np.random.default_rng(seed)
    M, K = int(rng.integers(1, 400)), int(rng.integers(40, 3000))
    rp, ci, v = random_csr(M, K, int(rng.integers(1, 30)), seed, np.float32, long_row=min(K, 200))
    for r in range(M):                       # stored order is arbitrary
        rng.shuffle(ci[rp[r]:rp[r + 1]])
    for row_bytes, elem, budget in ((32, 4, 4096), (64, 8, 20000), (256, 4, 114000)):
        plan = sx.plan_edge_lists(M, K, rp, ci, row_bytes, elem, budget)
        if len(plan[0]) == 0:                # some row alone exceeds the budget
            lens = np.diff(rp)
            worst = int(lens.max())
            assert worst * (row_bytes + 4) + (worst + 14) * (elem + 2) > budget * 0.5
            continue
        check_plan(M, K, rp, ci, row_bytes, elem, budget, *plan)
        plan = sx.plan_edge_lists(M, K, rp, ci, row_bytes, elem, budget, max_rows=128, nnz_target=50)
        if len(plan[0]):
            check_plan(M, K, rp, ci, row_bytes, elem, budget, *plan, rows=128, target=50)


def test_degenerate_inputs():
    rp = np.zeros(1, np.int32)
    b, r, l, t, m, p = sx.plan_edge_lists(0, 5, rp, np.zeros(0, np.int32), 64, 4, 4096)
    assert len(b) == 0 and t == 0
    rp = np.zeros(41, np.int32)              # 40 empty rows: blocks without runs
    b, r, l, t, m, p = sx.plan_edge_lists(40, 5, rp, np.zeros(0, np.int32), 64, 4, 4096)
    assert len(b) == 2 and t == 0 and b[:, 5].sum() == 0
    with pytest.raises(sx.SextansError):
        sx.plan_edge_lists(4, 5, np.zeros(5, np.int32), np.zeros(0, np.int32), 60, 4, 4096)   # row_bytes % 16
