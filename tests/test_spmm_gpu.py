"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and
the committed golden vectors of the reference.

Bars:
  strict mode (default), rows <= split threshold : BIT-EXACT vs cpu_spmm_CSR (fp32) and
                                                   its double restatement (fp64)
  fast mode / split rows                         : max |x-y|/max|y| <= 1e-5 (fp32), 1e-12 (fp64)
  BASELINE contract                              : rel-err <= 1e-6 on the fp64 configs
  always                                         : the reference's own pass criterion
                                                   (src/sextans-host.cpp:262-289)
"""
import os

import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import (GOLDEN, SUITESPARSE, max_rel_err, mtx_path, perturbed_inputs, random_csr,
                     random_dense, scaled_err, sha)

pytestmark = pytest.mark.gpu

A32, B32 = np.float32(0.85), np.float32(-2.06)   # the host program's defaults (host.cpp:29-30)


@pytest.fixture(scope="module")
def eng():
    e = sx.Engine(0)
    yield e
    e.close()


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def run(eng, M, K, N, rp, ci, v, alpha, B, beta, Cin, rp_time=1):
    eng.upload_csr(M, K, rp, ci, v)
    C = Cin.copy()
    ns = eng.spmm(N, alpha, B, beta, C, rp_time)
    return C, ns


def test_native_library_is_the_one_running(eng):
    # loud failure instead of a fallback: the context exists only if CUDA does
    assert os.path.exists(sx.library_path())
    assert eng.info(sx.INFO_LAUNCHES) == 0
    with open("/proc/self/maps") as f:
        assert "libsextans_b200.so" in f.read()


def test_small_golden_cases_bit_exact(eng):
    g = np.load(os.path.join(GOLDEN, "spmm_small.npz"))
    before = eng.launches
    for tag in "abcdef":
        M, K, N = g[tag + "_dims"].tolist()
        alpha, beta = g[tag + "_ab"]
        C, _ = run(eng, M, K, N, g[tag + "_rowptr"], g[tag + "_colidx"], g[tag + "_val"], alpha,
                   g[tag + "_B"], beta, g[tag + "_Cin"])
        assert np.array_equal(bits(C), bits(g[tag + "_C"])), tag
    assert eng.launches > before


@pytest.mark.parametrize("name", SUITESPARSE)
def test_suitesparse_runs_match_reference_golden(eng, golden, name):
    """C1/C3: every golden run of the reference's cpu_spmm_CSR, bit for bit, on matrices
    loaded by the product loader."""
    g = golden["suitesparse"][name]
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path(name), np.float32)
    assert (M, K, nnz) == (g["M"], g["K"], g["nnz"])
    eng.upload_csr(M, K, rp, ci, v)
    for r in g["runs"]:
        N = r["N"]
        if r["kind"] == "default":
            B, C = oracle.init_dense(M, K, N, np.float32)
            eng.upload_csr(M, K, rp, ci, v)
        else:
            val, B, C = perturbed_inputs(M, K, N, nnz, np.float32)
            eng.upload_csr(M, K, rp, ci, val)
        ns = eng.spmm(N, r["alpha"], B, r["beta"], C)
        assert ns > 0
        assert sha(C) == r["C_sha256"], (name, r["kind"], N)
        assert float(C[0]) == r["C0"] and float(C[M - 1]) == r["C_Mm1"] and float(C[-1]) == r["C_last"]


def test_config2_nasa4704_n16_f64(eng, golden):
    """C2: nasa4704, N=16, fp64 on the GPU; rel-err <= 1e-6 vs the CPU path (contract),
    and in fact bit-exact vs the fp64 restatement; ~1e-7 vs the reference's fp32 run."""
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path("nasa4704"), np.float64)
    N = 16
    a, b = float(A32), float(B32)
    B, Cin = oracle.init_dense(M, K, N, np.float64)
    C, ns = run(eng, M, K, N, rp, ci, v, a, B, b, Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
    assert max_rel_err(C, ref) <= 1e-6
    assert np.array_equal(bits(C), bits(ref))
    # against the fp32 reference numbers themselves
    run32 = [r for r in golden["suitesparse"]["nasa4704"]["runs"] if r["kind"] == "default" and r["N"] == 16][0]
    assert abs(C[0] - run32["C0"]) / abs(run32["C0"]) < 1e-6
    assert abs(C.sum() - run32["sum"]) / abs(run32["sum"]) < 1e-6
    # value-sensitive inputs too
    val, B, Cin = perturbed_inputs(M, K, N, nnz, np.float64)
    C, _ = run(eng, M, K, N, rp, ci, val, a, B, b, Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, val, a, B, b, Cin.copy())
    assert np.array_equal(bits(C), bits(ref))
    n_bad, pct, ok = oracle.verify_f32(ref.astype(np.float32), C.astype(np.float32), M, N)
    assert ok and n_bad == 0


SHAPES = [  # M, K, avg nnz/row, N
    (1, 1, 1, 8), (7, 5, 2, 8), (33, 70, 3, 16), (100, 64, 10, 24), (257, 300, 17, 32),
    (64, 1000, 40, 40), (500, 200, 8, 64), (300, 300, 30, 128), (90, 50, 6, 136),
    (130, 77, 9, 256), (41, 60, 5, 520), (1000, 1000, 3, 1), (50, 50, 4, 3), (20, 20, 2, 12),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,avg,N", SHAPES)
def test_random_csr_bit_exact(eng, dtype, M, K, avg, N):
    rp, ci, v = random_csr(M, K, avg, M * 31 + N, dtype)
    B, Cin = random_dense(M, K, N, M * 31 + N, dtype)
    a, b = dtype(0.85), dtype(-2.06)
    C, _ = run(eng, M, K, N, rp, ci, v, a, B, b, Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
    assert np.array_equal(bits(C), bits(ref))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stored_order_unsorted_columns_and_duplicates(eng, dtype):
    # the kernel must follow STORED order, not column order: unsorted rows with
    # repeated columns still match the oracle bit for bit
    rng = np.random.default_rng(5)
    M, K, N = 200, 64, 16
    lens = rng.integers(0, 40, size=M)
    rp = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = rng.integers(0, K, size=rp[-1]).astype(np.int32)
    v = rng.uniform(-1, 1, size=rp[-1]).astype(dtype)
    B, Cin = random_dense(M, K, N, 5, dtype)
    C, _ = run(eng, M, K, N, rp, ci, v, dtype(1.5), B, dtype(0.25), Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(1.5), B, dtype(0.25), Cin.copy())
    assert np.array_equal(bits(C), bits(ref))


@pytest.mark.parametrize("alpha,beta", [(0.0, 1.0), (1.0, 0.0), (0.0, 0.0), (-3.5, 7.25)])
def test_alpha_beta_corner_values(eng, alpha, beta):
    M, K, N = 150, 90, 16
    rp, ci, v = random_csr(M, K, 7, 11, np.float32)
    B, Cin = random_dense(M, K, N, 11, np.float32)
    C, _ = run(eng, M, K, N, rp, ci, v, np.float32(alpha), B, np.float32(beta), Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(alpha), B, np.float32(beta), Cin.copy())
    assert np.array_equal(bits(C), bits(ref))


def test_beta_zero_still_propagates_nan_like_the_reference(eng):
    # the reference always evaluates BETA*C (sparse_helper.h:287): 0*NaN = NaN
    M, K, N = 10, 10, 8
    rp, ci, v = random_csr(M, K, 3, 3, np.float32, empty_frac=0)
    B, Cin = random_dense(M, K, N, 3, np.float32)
    Cin[5] = np.nan
    C, _ = run(eng, M, K, N, rp, ci, v, np.float32(1), B, np.float32(0), Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(1), B, np.float32(0), Cin.copy())
    assert np.isnan(C[5]) and np.isnan(ref[5])
    assert np.array_equal(np.isnan(C), np.isnan(ref))


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
@pytest.mark.parametrize("N", [8, 16, 128])
def test_long_rows_take_the_split_path(eng, dtype, tol, N, kernel):
    M, K = 300, 5000
    rp, ci, v = random_csr(M, K, 20, 77, dtype, long_row=4000)
    B, Cin = random_dense(M, K, N, 77, dtype)
    a, b = dtype(0.85), dtype(-2.06)
    eng.set_option(sx.OPT_KERNEL, kernel)
    try:
        eng.set_option(sx.OPT_SPLIT_ROW_NNZ, 512)
        C, _ = run(eng, M, K, N, rp, ci, v, a, B, b, Cin)
        assert eng.info(sx.INFO_SPLIT_ROWS) == 1
        ref = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
        assert scaled_err(C, ref) <= tol
        long_row = int(np.argmax(np.diff(rp)))
        keep = np.ones(M, dtype=bool)
        keep[long_row] = False
        Cm, Rm = C.reshape(N, M)[:, keep], ref.reshape(N, M)[:, keep]
        assert np.array_equal(bits(np.ascontiguousarray(Cm)), bits(np.ascontiguousarray(Rm)))
        # with splitting disabled the long row is walked in order as well: bit-exact everywhere
        eng.set_option(sx.OPT_SPLIT_ROW_NNZ, 0)
        C2 = Cin.copy()
        eng.spmm(N, a, B, b, C2)
        assert eng.info(sx.INFO_SPLIT_ROWS) == 0
        assert np.array_equal(bits(C2), bits(ref))
    finally:
        eng.set_option(sx.OPT_SPLIT_ROW_NNZ, 512)
        eng.set_option(sx.OPT_KERNEL, 0)


KERNEL_SHAPES = [(7, 5, 2, 8), (257, 300, 17, 32), (64, 1000, 40, 40), (500, 200, 8, 64),
                 (300, 300, 30, 128), (130, 77, 9, 256), (1000, 1000, 3, 1), (2000, 3000, 70, 16)]


@pytest.mark.parametrize("kernel", [1, 2, 3, 5])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,avg,N", KERNEL_SHAPES)
def test_every_kernel_variant_bit_exact(eng, kernel, dtype, M, K, avg, N):
    """Variants 1 (row per lane group), 2 (TMA-staged work items, at several item budgets),
    3 (TMA-staged B window; defers to 1 when a window does not fit in shared memory or
    a dense row is wider than 256 bytes) and 5 (edge lists: a row block's distinct B rows staged
    by TMA, 16-bit local columns; forced here on matrices without any reuse; dense rows wider
    than 256 bytes go to the other variants) all reproduce the oracle bit for bit."""
    rp, ci, v = random_csr(M, K, avg, M * 17 + N + kernel, dtype, long_row=min(K, 300))
    B, Cin = random_dense(M, K, N, M * 17 + N, dtype)
    eng.set_option(sx.OPT_KERNEL, kernel)
    try:
        for item_nnz in ((0, 8, 64) if kernel == 2 else (0,)):
            eng.set_option(sx.OPT_ITEM_NNZ, item_nnz)
            C, _ = run(eng, M, K, N, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin)
            ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
            assert np.array_equal(bits(C), bits(ref)), (kernel, item_nnz)
            family = eng.info(sx.INFO_LAST_KERNEL) // 10000
            if kernel == 5:
                assert family == 8 if N * np.dtype(dtype).itemsize <= 256 else family in (1, 2, 3)
            else:
                assert family in ((kernel,) if kernel != 3 else (3, 1))
    finally:
        eng.set_option(sx.OPT_KERNEL, 0)
        eng.set_option(sx.OPT_ITEM_NNZ, 0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,avg,N", KERNEL_SHAPES + [(3000, 50000, 90, 16), (800, 20000, 120, 128)])
def test_l2_prefetch_of_the_next_batch_changes_nothing(eng, dtype, M, K, avg, N):
    """SX_OPT_PREFETCH only moves data towards L2 earlier: the staged kernel's results stay
    bit-identical to the oracle (long rows included, at several item budgets)."""
    rp, ci, v = random_csr(M, K, avg, M * 19 + N, dtype, long_row=min(K, 300))
    B, Cin = random_dense(M, K, N, M * 19 + N, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.set_option(sx.OPT_KERNEL, 2)
    eng.set_option(sx.OPT_PREFETCH, 1)
    try:
        for item_nnz in (0, 64, 2048):
            eng.set_option(sx.OPT_ITEM_NNZ, item_nnz)
            C, _ = run(eng, M, K, N, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin)
            assert np.array_equal(bits(C), bits(ref)), item_nnz
            assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 2
    finally:
        eng.set_option(sx.OPT_KERNEL, 0)
        eng.set_option(sx.OPT_ITEM_NNZ, 0)
        eng.set_option(sx.OPT_PREFETCH, -1)


def test_auto_prefetch_regime_large_B_narrow_rows(eng):
    # B of 77 MB (> 32 MiB) with 128-byte rows: the auto rule of SX_OPT_PREFETCH switches the
    # prefetch on in the staged kernel; the result is the oracle's, bit for bit
    M, K, N = 4000, 600000, 16
    rp, ci, v = random_csr(M, K, 60, 4242, np.float64)
    B, Cin = random_dense(M, K, N, 4242, np.float64)
    eng.set_option(sx.OPT_KERNEL, 2)
    try:
        C, _ = run(eng, M, K, N, rp, ci, v, 0.85, B, -2.06, Cin)
    finally:
        eng.set_option(sx.OPT_KERNEL, 0)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
    assert np.array_equal(bits(C), bits(ref))


def test_auto_kernel_choice(eng):
    # less than one wave of row groups and narrow column windows: variant 3 (else 1);
    # more rows than that: the staged variant 2
    rp, ci, v = random_csr(3000, 500, 10, 1, np.float32)
    B, Cin = random_dense(3000, 500, 16, 1, np.float32)
    run(eng, 3000, 500, 16, rp, ci, v, A32, B, B32, Cin)
    assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 3
    # same size, but a row block spans more columns than shared memory holds: variant 1
    rp, ci, v = random_csr(3000, 60000, 10, 1, np.float32)
    B, Cin = random_dense(3000, 60000, 16, 1, np.float32)
    C, _ = run(eng, 3000, 60000, 16, rp, ci, v, A32, B, B32, Cin)
    assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 1
    assert np.array_equal(bits(C), bits(oracle.spmm_csr(3000, 16, 60000, rp, ci, v, A32, B, B32, Cin.copy())))
    # many rows, two columns shared by all of them: a staged B row serves many nonzeros, variant 5
    M = 100000
    rp = (np.arange(M + 1) * 2).astype(np.int32)
    ci = np.tile(np.array([1, 7], dtype=np.int32), M)
    v = np.ones(2 * M, np.float32)
    B, Cin = random_dense(M, 16, 16, 2, np.float32)
    C, _ = run(eng, M, 16, 16, rp, ci, v, A32, B, B32, Cin)
    assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 8
    assert np.array_equal(bits(C), bits(oracle.spmm_csr(M, 16, 16, rp, ci, v, A32, B, B32, Cin.copy())))
    # many rows, columns all over a wide B: the nnz-balanced staged variant 2
    K = 500000
    ci = np.random.default_rng(3).integers(0, K, size=2 * M).astype(np.int32)
    B, Cin = random_dense(M, K, 16, 2, np.float32)
    C, _ = run(eng, M, K, 16, rp, ci, v, A32, B, B32, Cin)
    assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 2
    assert np.array_equal(bits(C), bits(oracle.spmm_csr(M, 16, K, rp, ci, v, A32, B, B32, Cin.copy())))


def test_host_paths_zero_copy_and_copy_engine(eng):
    """Page-locked operands are read/written by the kernels directly (no memcpy);
    pageable ones, or anything above SX_OPT_ZEROCOPY_BYTES, take cudaMemcpyAsync.  Same bits."""
    M, K, N = 900, 700, 24
    rp, ci, v = random_csr(M, K, 11, 3, np.float64)
    B, Cin = random_dense(M, K, N, 3, np.float64)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
    eng.upload_csr(M, K, rp, ci, v)
    C = Cin.copy()
    eng.spmm(N, 0.85, B, -2.06, C)
    assert eng.info(sx.INFO_HOST_PATH) == 0 and np.array_equal(bits(C), bits(ref))
    pB, pC = sx.pinned_empty(B.size, np.float64), sx.pinned_empty(Cin.size, np.float64)
    pB[:] = B
    pC[:] = Cin
    eng.spmm(N, 0.85, pB, -2.06, pC, rp_time=3)
    assert eng.info(sx.INFO_HOST_PATH) == 1 and np.array_equal(bits(np.asarray(pC)), bits(ref))
    eng.set_option(sx.OPT_ZEROCOPY_BYTES, 0)
    try:
        pC[:] = Cin
        eng.spmm(N, 0.85, pB, -2.06, pC)
        assert eng.info(sx.INFO_HOST_PATH) == 0 and np.array_equal(bits(np.asarray(pC)), bits(ref))
    finally:
        eng.set_option(sx.OPT_ZEROCOPY_BYTES, 3 << 19)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_fast_arithmetic_within_tolerance(dtype, tol):
    M, K, N = 400, 400, 32
    rp, ci, v = random_csr(M, K, 25, 21, dtype)
    B, Cin = random_dense(M, K, N, 21, dtype)
    with sx.Engine(0, arith=sx.FAST) as e:
        C, _ = run(e, M, K, N, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin)
        assert e.info(sx.INFO_LAST_KERNEL) % 10 == 1
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    assert scaled_err(C, ref) <= tol
    n_bad, pct, ok = oracle.verify_f32(ref.astype(np.float32), C.astype(np.float32), M, N)
    assert ok


def test_result_independent_of_rp_time(eng):
    M, K, N = 333, 222, 16
    rp, ci, v = random_csr(M, K, 9, 4, np.float64)
    B, Cin = random_dense(M, K, N, 4, np.float64)
    C1, ns1 = run(eng, M, K, N, rp, ci, v, 0.85, B, -2.06, Cin, rp_time=1)
    C5, ns5 = run(eng, M, K, N, rp, ci, v, 0.85, B, -2.06, Cin, rp_time=5)
    C0, _ = run(eng, M, K, N, rp, ci, v, 0.85, B, -2.06, Cin, rp_time=0)   # treated as 1
    assert np.array_equal(C1, C5) and np.array_equal(C1, C0)
    assert ns1 > 0 and ns5 > 0


def test_staged_calls_equal_the_one_shot_call(eng):
    M, K, N = 123, 77, 24
    rp, ci, v = random_csr(M, K, 6, 9, np.float32)
    B, Cin = random_dense(M, K, N, 9, np.float32)
    C, _ = run(eng, M, K, N, rp, ci, v, A32, B, B32, Cin)
    eng.stage_B(N, B)
    eng.stage_C(N, Cin)
    assert eng.launch(A32, B32, 3) > 0
    out = np.empty_like(Cin)
    eng.fetch_C(out)
    assert np.array_equal(out, C)
    # pinned host operands go through the same path
    pB, pC = sx.pinned_empty(B.size, np.float32), sx.pinned_empty(Cin.size, np.float32)
    pB[:] = B
    pC[:] = Cin
    eng.spmm(N, A32, pB, B32, pC)
    assert np.array_equal(np.asarray(pC), C)


def test_device_resident_operands_and_layout_changes(eng):
    import torch
    M, K, N = 1000, 800, 16
    rp, ci, v = random_csr(M, K, 12, 13, np.float64)
    B, Cin = random_dense(M, K, N, 13, np.float64)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
    eng.upload_csr(M, K, rp, ci, v)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            dB_cm = torch.from_numpy(B).cuda()          # column-major images, as on the host
            dC_cm = torch.from_numpy(Cin).cuda()
            dB = torch.empty(K * N, dtype=torch.float64, device="cuda")
            dC = torch.empty(M * N, dtype=torch.float64, device="cuda")
            eng.colmajor_to_rowmajor(K, N, dB_cm, dB, N)
            eng.colmajor_to_rowmajor(M, N, dC_cm, dC, N)
            assert torch.equal(dB.view(K, N), dB_cm.view(N, K).t())
            dOut = torch.empty_like(dC)
            eng.spmm_device(N, 0.85, dB, N, -2.06, dC, dOut, N)
            eng.spmm_device(N, 0.85, dB, N, -2.06, dC, dC, N)       # in place
            back = torch.empty_like(dC_cm)
            eng.rowmajor_to_colmajor(M, N, dOut, N, back)
            stream.synchronize()
        assert np.array_equal(back.cpu().numpy(), ref)
        assert torch.equal(dOut, dC)
    finally:
        eng.set_stream(None)


def test_row_blocks_reproduce_the_whole_bitwise(eng):
    """(e) multi-GPU invariant on one device: SpMM of nnz-balanced row blocks, each on its
    own context with the full B, concatenates to exactly the single-context result."""
    name = "pcrystk02"
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path(name), np.float32)
    N = 16
    val, B, Cin = perturbed_inputs(M, K, N, nnz, np.float32)
    whole, _ = run(eng, M, K, N, rp, ci, val, A32, B, B32, Cin)
    bounds = sx.partition_rows(rp, 3)
    Cw = whole.reshape(N, M)
    Cin2 = Cin.reshape(N, M)
    for p in range(3):
        r0, r1 = int(bounds[p]), int(bounds[p + 1])
        sub_rp = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
        sl = slice(rp[r0], rp[r1])
        with sx.Engine(0) as e:
            e.upload_csr(r1 - r0, K, sub_rp, ci[sl], val[sl])
            Cb = np.ascontiguousarray(Cin2[:, r0:r1]).ravel()
            e.spmm(N, A32, B, B32, Cb)
        assert np.array_equal(Cb.reshape(N, r1 - r0), Cw[:, r0:r1])


def test_error_paths(eng):
    with sx.Engine(0) as e:
        with pytest.raises(sx.SextansError, match="STATE"):
            e.dtype = np.dtype(np.float32)
            e.M = e.K = 1
            e.spmm(8, 1.0, np.zeros(8, np.float32), 0.0, np.zeros(8, np.float32))
        with pytest.raises(sx.SextansError, match="out of range"):
            e.upload_csr(2, 2, np.array([0, 1, 2]), np.array([0, 2]), np.ones(2, np.float32))
        with pytest.raises(sx.SextansError, match="rowptr"):
            e.upload_csr(2, 2, np.array([0, 2, 1]), np.array([0]), np.ones(1, np.float32))
        e.upload_csr(2, 2, np.array([0, 1, 2]), np.array([0, 1]), np.ones(2, np.float32))
        with pytest.raises(sx.SextansError, match="INVALID"):
            sx._check(e._L.sx_spmm_f64(e._ctx, 8, 1.0, None, 0.0, None, 1, None))   # f64 call on an f32 matrix
        with pytest.raises(sx.SextansError, match="NO_DEVICE"):
            sx.Engine(1000)
    # an empty matrix and a matrix with only empty rows are fine
    with sx.Engine(0) as e:
        e.upload_csr(3, 4, np.zeros(4, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64))
        C = np.arange(24, dtype=np.float64)
        e.spmm(8, 2.0, np.ones(32), 0.5, C)
        assert np.array_equal(C, np.arange(24) * 0.5)


def test_medium_matrix_sampled_rows_and_linearity(eng):
    """Size-independent properties at a size the full oracle would take long on:
    sampled rows against the oracle, and linearity in B."""
    rng = np.random.default_rng(99)
    M = K = 200_000
    per = 20
    N = 32
    cols = np.sort(rng.integers(0, K - per, size=(M, per)), axis=1) + np.arange(per)
    rp = (np.arange(M + 1) * per).astype(np.int32)
    ci = cols.astype(np.int32).ravel()
    v = rng.uniform(-1, 1, size=ci.size).astype(np.float32)
    B1 = rng.uniform(-1, 1, size=K * N).astype(np.float32)
    Cin = rng.uniform(-1, 1, size=M * N).astype(np.float32)
    C, _ = run(eng, M, K, N, rp, ci, v, A32, B1, B32, Cin)
    rows = np.unique(rng.integers(0, M, size=2000)).astype(np.int32)
    S = oracle.spmm_csr_rows(M, N, K, rp, ci, v, A32, B1, B32, Cin, rows)
    assert np.array_equal(bits(np.ascontiguousarray(C.reshape(N, M).T[rows])), bits(S))
    # linearity: A(2*B) with beta=0 equals 2*(A B) exactly (scaling by 2 is exact)
    z = np.zeros_like(Cin)
    y1 = z.copy(); eng.spmm(N, 1.0, B1, 0.0, y1)
    y2 = z.copy(); eng.spmm(N, 1.0, 2 * B1, 0.0, y2)
    assert np.array_equal(y2, 2 * y1)


# ---- dense-tile (blocked) variant on the FP64 tensor cores -------------------------------
def _tiles_case(eng, M, K, N, rp, ci, v, tau, expect_tiles=True, alpha=0.85, beta=-2.06):
    B, Cin = random_dense(M, K, N, M + N, np.float64)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, alpha, B, beta, Cin.copy())
    eng.set_option(sx.OPT_TILE_MIN_ROWS, tau)
    try:
        eng.upload_csr(M, K, rp, ci, v)
        C = Cin.copy()
        eng.spmm(N, alpha, B, beta, C, rp_time=2)
        tile_nnz, slots, rest = eng.info(sx.INFO_TILE_NNZ), eng.info(sx.INFO_TILE_SLOTS), eng.info(sx.INFO_REST_NNZ)
        assert tile_nnz + rest == ci.size
        if expect_tiles:
            assert tile_nnz > 0 and slots >= tile_nnz and eng.info(sx.INFO_LAST_KERNEL) // 10000 in (1, 2, 4)
        # tolerance-level parity (summation order differs): 1e-12 of the largest entry, and
        # the BASELINE contract of 1e-6 relative with room to spare
        assert scaled_err(C, ref) <= 1e-12, scaled_err(C, ref)
        n_bad, pct, ok = oracle.verify_f32(ref.astype(np.float32), C.astype(np.float32), M, N)
        assert ok and n_bad == 0
        return tile_nnz, slots, rest
    finally:
        eng.set_option(sx.OPT_TILE_MIN_ROWS, 0)


@pytest.mark.parametrize("N", [8, 16, 24, 64, 72, 5])
def test_tiles_fem_like_matrix(eng, N):
    from sextans_b200.workloads import fem_like_csr
    rp, ci, v = fem_like_csr(500, 4, 6, seed=3, band=40, noise_per_row=2)
    M = K = 2000
    tile_nnz, slots, rest = _tiles_case(eng, M, K, N, rp, ci, v, tau=4)
    assert tile_nnz > 0.7 * ci.size and rest > 0          # blocks in tiles, noise in the remainder
    assert tile_nnz / slots > 0.45


@pytest.mark.parametrize("name", SUITESPARSE)
def test_tiles_on_the_shipped_matrices(eng, name):
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path(name), np.float64)
    val, _, _ = perturbed_inputs(M, K, 16, nnz, np.float64)
    tile_nnz, slots, rest = _tiles_case(eng, M, K, 16, rp, ci, val, tau=4)
    assert tile_nnz > (0.3 if name == "nasa4704" else 0.75) * nnz      # FEM matrices: a third / four fifths of A sits in dense panel columns


def test_tiles_edge_cases(eng):
    # M not a multiple of 8, empty rows, duplicates, an all-dense panel, tau extremes, no tiles at all
    rng = np.random.default_rng(8)
    M, K, N = 27, 40, 16
    dense = (rng.random((M, K)) < 0.5)
    dense[8:16] = True                      # a fully dense panel
    dense[16:19] = False                    # empty rows
    rows, cols = np.nonzero(dense)
    rp = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=M), out=rp[1:])
    ci = cols.astype(np.int32)
    v = rng.uniform(-1, 1, ci.size)
    for tau in (1, 4, 8):
        _tiles_case(eng, M, K, N, rp, ci, v, tau)
    # duplicates of one (row, col): the first goes to the tile, the others stay in the remainder
    rp2 = np.array([0, 3, 5, 6, 7, 8, 9, 10, 11], dtype=np.int32)
    ci2 = np.array([2, 2, 2, 2, 5, 2, 2, 2, 2, 2, 2], dtype=np.int32)
    v2 = rng.uniform(-1, 1, ci2.size)
    t, s_, r = _tiles_case(eng, 8, 6, 8, rp2, ci2, v2, tau=4)
    assert t == 8 and r == 3
    # a matrix without any shared column: nothing goes to tiles, plain CSR path, bit-exact
    rp3 = np.arange(9, dtype=np.int32)
    ci3 = np.arange(8, dtype=np.int32)
    t, s_, r = _tiles_case(eng, 8, 8, 8, rp3, ci3, np.ones(8), tau=2, expect_tiles=False)
    assert t == 0 and r == 8
    # fp32 has no exact tensor-core kind: refused loudly
    eng.set_option(sx.OPT_TILE_MIN_ROWS, 4)
    try:
        with pytest.raises(sx.SextansError, match="fp64 only"):
            eng.upload_csr(8, 8, rp3, ci3, np.ones(8, np.float32))
    finally:
        eng.set_option(sx.OPT_TILE_MIN_ROWS, 0)
