"""GPU parity tests of the column-window passes (SX_OPT_COL_WINDOW_ROWS): the staged
kernel run window by window with running sums, against the oracle -- bit-exact in strict
mode for rows stored in ascending column order, tolerance-level for split rows, fast
arithmetic and rows stored out of order."""
import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import random_csr, random_dense, scaled_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    e = sx.Engine(0)
    yield e
    e.close()


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def run_windows(eng, W, M, K, N, rp, ci, v, alpha, B, beta, Cin, rp_time=1):
    eng.set_option(sx.OPT_COL_WINDOW_ROWS, W)
    eng.upload_csr(M, K, rp, ci, v)
    C = Cin.copy()
    eng.spmm(N, alpha, B, beta, C, rp_time)
    return C


# every lane-group shape of the staged kernel: G = 2..32 and 1/2/4 vectors per lane
SHAPES = [  # M, K, avg, N, W
    (257, 300, 17, 8, 64), (500, 2000, 40, 16, 512), (300, 1500, 30, 24, 500), (400, 900, 25, 32, 256),
    (200, 1000, 40, 64, 300), (300, 800, 30, 128, 256), (130, 700, 19, 256, 128), (41, 600, 15, 520, 200),
    (1000, 1000, 3, 1, 10), (50, 50, 4, 3, 7), (64, 4097, 30, 16, 4096), (3000, 5000, 60, 16, 1024),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,avg,N,W", SHAPES)
def test_window_passes_bit_exact(eng, dtype, M, K, avg, N, W):
    rp, ci, v = random_csr(M, K, avg, M * 13 + N, dtype)
    B, Cin = random_dense(M, K, N, M * 13 + N, dtype)
    a, b = dtype(0.85), dtype(-2.06)
    for item_nnz in (0, 8, 64):
        eng.set_option(sx.OPT_ITEM_NNZ, item_nnz)
        C = run_windows(eng, W, M, K, N, rp, ci, v, a, B, b, Cin)
        nwin = (K + W - 1) // W
        assert eng.info(sx.INFO_COL_WINDOWS) == nwin and eng.info(sx.INFO_LAST_KERNEL) == 50000 + nwin
        ref = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
        assert np.array_equal(bits(C), bits(ref)), item_nnz


def test_windows_off_and_single_window_take_the_plain_path(eng):
    M, K, N = 300, 400, 16
    rp, ci, v = random_csr(M, K, 10, 1, np.float32)
    B, Cin = random_dense(M, K, N, 1, np.float32)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(0.85), B, np.float32(-2.06), Cin.copy())
    for W in (0, 400, 100000):
        C = run_windows(eng, W, M, K, N, rp, ci, v, np.float32(0.85), B, np.float32(-2.06), Cin)
        assert eng.info(sx.INFO_COL_WINDOWS) == 0 and eng.info(sx.INFO_LAST_KERNEL) < 50000
        assert np.array_equal(bits(C), bits(ref))
    # the option is read at upload: switching it off afterwards changes nothing until then
    C = run_windows(eng, 64, M, K, N, rp, ci, v, np.float32(0.85), B, np.float32(-2.06), Cin)
    assert eng.info(sx.INFO_COL_WINDOWS) == 7
    eng.set_option(sx.OPT_COL_WINDOW_ROWS, 0)
    C2 = Cin.copy()
    eng.spmm(N, np.float32(0.85), B, np.float32(-2.06), C2)
    assert eng.info(sx.INFO_LAST_KERNEL) == 50007 and np.array_equal(bits(C2), bits(ref))


@pytest.mark.parametrize("alpha,beta", [(0.0, 1.0), (1.0, 0.0), (0.0, 0.0), (-3.5, 7.25)])
def test_window_alpha_beta_corner_values_and_nan(eng, alpha, beta):
    M, K, N = 150, 900, 16
    rp, ci, v = random_csr(M, K, 30, 11, np.float32)
    B, Cin = random_dense(M, K, N, 11, np.float32)
    Cin[5] = np.nan
    C = run_windows(eng, 200, M, K, N, rp, ci, v, np.float32(alpha), B, np.float32(beta), Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(alpha), B, np.float32(beta), Cin.copy())
    assert np.array_equal(np.isnan(C), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(bits(C)[ok], bits(ref)[ok])


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_window_split_rows_and_fast_arithmetic(eng, dtype, tol):
    # one row with 4000 nonzeros: > 512 in every window of 1000 columns -> pieces + finalize
    # in every pass, the running sum going through the finalize kernel
    M, K, N = 300, 5000, 16
    rp, ci, v = random_csr(M, K, 20, 77, dtype, long_row=4000)
    B, Cin = random_dense(M, K, N, 77, dtype)
    a, b = dtype(0.85), dtype(-2.06)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, a, B, b, Cin.copy())
    C = run_windows(eng, 1000, M, K, N, rp, ci, v, a, B, b, Cin)
    assert scaled_err(C, ref) <= tol
    long_row = int(np.argmax(np.diff(rp)))
    keep = np.ones(M, dtype=bool)
    keep[long_row] = False
    assert np.array_equal(bits(np.ascontiguousarray(C.reshape(N, M)[:, keep])),
                          bits(np.ascontiguousarray(ref.reshape(N, M)[:, keep])))
    # without splitting the long row is walked in order through all windows: bit-exact
    eng.set_option(sx.OPT_SPLIT_ROW_NNZ, 0)
    C2 = Cin.copy()
    eng.spmm(N, a, B, b, C2)
    assert np.array_equal(bits(C2), bits(ref))
    eng.set_option(sx.OPT_SPLIT_ROW_NNZ, 512)
    eng.set_option(sx.OPT_ARITH, sx.FAST)
    C3 = Cin.copy()
    eng.spmm(N, a, B, b, C3)
    assert scaled_err(C3, ref) <= tol


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_window_rows_stored_out_of_order_are_within_tolerance(eng, dtype, tol):
    rng = np.random.default_rng(5)
    M, K, N = 200, 640, 16
    lens = rng.integers(0, 40, size=M)
    rp = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = rng.integers(0, K, size=rp[-1]).astype(np.int32)   # unsorted, with duplicates
    v = rng.uniform(-1, 1, size=rp[-1]).astype(dtype)
    B, Cin = random_dense(M, K, N, 5, dtype)
    C = run_windows(eng, 100, M, K, N, rp, ci, v, dtype(1.5), B, dtype(0.25), Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(1.5), B, dtype(0.25), Cin.copy())
    assert scaled_err(C, ref) <= tol


def test_window_result_independent_of_rp_time_and_repeatable(eng):
    M, K, N = 500, 3000, 32
    rp, ci, v = random_csr(M, K, 50, 21, np.float64)
    B, Cin = random_dense(M, K, N, 21, np.float64)
    C1 = run_windows(eng, 700, M, K, N, rp, ci, v, 0.85, B, -2.06, Cin)
    C3 = Cin.copy()
    eng.spmm(N, 0.85, B, -2.06, C3, 3)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
    assert np.array_equal(bits(C1), bits(ref)) and np.array_equal(bits(C3), bits(ref))


def test_window_device_resident_in_place(eng):
    torch = pytest.importorskip("torch")
    M, K, N = 700, 2500, 16
    rp, ci, v = random_csr(M, K, 40, 8, np.float64)
    B, Cin = random_dense(M, K, N, 8, np.float64)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
    eng.set_option(sx.OPT_COL_WINDOW_ROWS, 600)
    eng.upload_csr(M, K, rp, ci, v)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            dB = torch.from_numpy(np.ascontiguousarray(B.reshape(N, K).T)).cuda()     # row-major K x N
            dC = torch.from_numpy(np.ascontiguousarray(Cin.reshape(N, M).T)).cuda()   # row-major M x N
            dOut = torch.empty_like(dC)
            eng.spmm_device(N, 0.85, dB, N, -2.06, dC, dOut, N)
            eng.spmm_device(N, 0.85, dB, N, -2.06, dC, dC, N)                         # in place
            stream.synchronize()
        assert eng.info(sx.INFO_LAST_KERNEL) == 50005
        assert torch.equal(dOut, dC)
        got = np.ascontiguousarray(dC.cpu().numpy().T).ravel()
        assert np.array_equal(bits(got), bits(ref))
    finally:
        eng.set_stream(None)
