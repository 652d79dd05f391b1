"""Property tests (hypothesis) of host-side pieces of the path that need no GPU:
the FPGA-image decoder against the reference's own encoder, the column-window split against
its definition, and the row-block partitioner's invariants."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle
import sextans_b200 as sx

HAVE_REF = oracle.ref() is not None and hasattr(oracle.ref(), "sxref_build_images")


@st.composite
def csr_matrices(draw, max_rows=150, max_cols=9000):
    M = draw(st.integers(1, max_rows))
    K = draw(st.integers(1, max_cols))
    density = draw(st.sampled_from([0, 1, 3, 8]))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    lens = np.minimum(rng.poisson(density, size=M), K)
    rp = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(K, size=int(n), replace=False)) for n in lens] + [np.zeros(0, dtype=np.int64)])
    v = rng.standard_normal(ci.size).astype(np.float32)
    return M, K, rp, ci.astype(np.int32), v


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libsextans_ref.so not built")
@settings(max_examples=60, deadline=None)
@given(a=csr_matrices())
def test_image_decoder_inverts_the_reference_encoder(a):
    M, K, rp, ci, v = a
    ptr, imgs, num_a_len = oracle.ref_build_images(M, K, rp, ci, v)
    assert ptr.size == (K + 4095) // 4096 + 1 and ptr[-1] == num_a_len
    rp2, ci2, v2 = sx.images_decode_A(ptr, imgs, M, K)
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci)
    assert np.array_equal(v2.view(np.uint32), v.view(np.uint32))


@settings(max_examples=60, deadline=None)
@given(a=csr_matrices(max_cols=600), W=st.integers(1, 700))
def test_column_window_split_partitions_every_row_in_order(a, W):
    M, K, rp, ci, _ = a
    if (K + W - 1) // W > 4096:
        return
    wrp, base, order, asc = sx.split_col_windows(M, K, rp, ci, W)
    nwin = max(1, (K + W - 1) // W)
    assert wrp.shape == (nwin, M + 1) and asc
    assert np.array_equal(np.sort(order), np.arange(rp[M]))
    # concatenating a row's window segments in window order gives the row back
    for r in range(M):
        got = np.concatenate([order[base[w] + wrp[w, r]: base[w] + wrp[w, r + 1]] for w in range(nwin)])
        assert np.array_equal(got, np.arange(rp[r], rp[r + 1]))
    for w in range(nwin):
        seg = order[base[w]:base[w + 1]]
        assert np.all(ci[seg] // W == w)


@settings(max_examples=80, deadline=None)
@given(lens=st.lists(st.integers(0, 50), min_size=0, max_size=200), parts=st.integers(1, 9))
def test_row_partition_invariants(lens, parts):
    rp = np.zeros(len(lens) + 1, dtype=np.int32)
    np.cumsum(np.asarray(lens, dtype=np.int64), out=rp[1:])
    b = sx.partition_rows(rp, parts)
    M, nnz = len(lens), int(rp[-1])
    assert b[0] == 0 and b[-1] == M and np.all(np.diff(b) >= 0)
    if nnz > 0:
        # no block starts before its share of the nonzeros begins, and it begins at most one row late
        for p in range(1, parts):
            target = -(-nnz * p // parts)
            assert rp[b[p]] >= target or b[p] == M
            assert b[p] == 0 or rp[b[p] - 1] < target or b[p] == b[p - 1]


@st.composite
def banded_matrices(draw):
    M = draw(st.integers(1, 400))
    K = draw(st.integers(8, 500))
    half = draw(st.integers(0, 60))
    jitter = draw(st.integers(0, 40))
    per_row = draw(st.integers(0, 6))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    rp = np.zeros(M + 1, dtype=np.int32)
    cols = []
    for r in range(M):
        centre = int(np.clip(r * K // M + rng.integers(-jitter, jitter + 1), 0, K - 1))
        lo, hi = max(0, centre - half), min(K, centre + half + 1)
        n = 0 if rng.random() < 0.15 else min(per_row, hi - lo)
        cols.append(np.sort(rng.choice(np.arange(lo, hi), size=n, replace=False)))
        rp[r + 1] = rp[r] + n
    ci = np.concatenate(cols + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    return M, K, rp, ci


@settings(max_examples=150, deadline=None)
@given(a=banded_matrices(), nchains=st.integers(1, 12))
def test_slide_plan_keeps_every_needed_row_in_the_ring(a, nchains):
    """Worst-case replay of the sliding-window kernel's schedule (spmm_slide_kernel): the loads
    of step s+1 may have landed completely before step s computes, the loads of step s+2 are
    issued right after step s; with a ring of exactly ring_rows rows (tighter than the power
    of two the kernel uses) every column a step touches must be resident when it computes,
    and the A buffers must be large enough."""
    M, K, rp, ci = a
    steps, chains, ring_rows, max_entries = sx.plan_slide(M, rp, ci, nchains)
    nsteps = (M + 31) // 32
    assert steps.shape == (nsteps, 4) and 1 <= chains.shape[0] <= min(nchains, nsteps)
    assert chains[0, 0] == 0 and chains[-1, 1] == nsteps and np.all(chains[1:, 0] == chains[:-1, 1])
    assert np.all(chains[:, 1] > chains[:, 0]) and max_entries % 4 == 0
    for R in (ring_rows, 1 << max(5, int(np.ceil(np.log2(max(ring_rows, 1)))))):
        for b, e in chains:
            ring = np.full(R, -1, dtype=np.int64)

            def load(s):
                lo, hi = int(steps[s, 0]), int(steps[s, 1])
                assert 0 <= lo <= hi <= K
                rows = np.arange(lo, hi)
                ring[rows % R] = rows

            load(b)
            if b + 1 < e:
                load(b + 1)
            for s in range(b, e):
                r0, r1 = s * 32, min(M, s * 32 + 32)
                assert steps[s, 2] == rp[r0] and steps[s, 3] == rp[r1]
                assert steps[s, 3] - (steps[s, 2] & ~3) <= max_entries
                need = ci[rp[r0]:rp[r1]].astype(np.int64)
                assert np.array_equal(ring[need % R], need), (s, R)
                if s + 2 < e:
                    load(s + 2)
