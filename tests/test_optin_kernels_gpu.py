"""GPU parity of the opt-in kernel paths: variant 3 under programmatic dependent launch
(SX_OPT_PDL = 1), the sliding-window kernel (SX_OPT_SLIDE + SX_OPT_KERNEL = 4) and the measured
variant choice (SX_OPT_AUTOTUNE).  All of them passed on a B200 in round 2 (profiles/)."""
import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import mtx_path, random_dense

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    e = sx.Engine(0)
    yield e
    e.close()


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def banded_csr(M, K, half_band, per_row, seed, dtype):
    """rows with ascending distinct columns inside [r - half_band, r + half_band]"""
    rng = np.random.default_rng(seed)
    rp = np.zeros(M + 1, dtype=np.int32)
    cols = []
    for r in range(M):
        lo, hi = max(0, r * K // M - half_band), min(K, r * K // M + half_band + 1)
        n = min(per_row if r % 7 else 0, hi - lo)
        cols.append(np.sort(rng.choice(np.arange(lo, hi), size=n, replace=False)))
        rp[r + 1] = rp[r] + n
    ci = np.concatenate(cols).astype(np.int32)
    return rp, ci, rng.uniform(-1, 1, ci.size).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pdl_window_kernel_in_a_dependent_chain(eng, dtype):
    """SX_OPT_PDL: back-to-back launches of variant 3 whose C_in is the previous launch's
    C_out (in place) and whose B was produced by a kernel just before -- the early-started
    prologue must not touch either before the previous kernel is complete."""
    torch = pytest.importorskip("torch")
    M = K = 4096
    N = 16
    rp, ci, v = banded_csr(M, K, 200, 24, 77, dtype)
    B, Cin = random_dense(M, K, N, 77, dtype)
    ref = Cin.copy()
    for _ in range(5):
        ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.5), B, dtype(0.25), ref)
    eng.set_option(sx.OPT_KERNEL, 3)
    eng.set_option(sx.OPT_PDL, 1)
    eng.upload_csr(M, K, rp, ci, v)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            dB_cm = torch.from_numpy(B).cuda()
            dC = torch.from_numpy(np.ascontiguousarray(Cin.reshape(N, M).T)).cuda()   # row-major M x N
            dB = torch.empty(K * N, dtype=dB_cm.dtype, device="cuda")
            for rep in range(3):                         # also exercises re-running the whole chain
                dCw = dC.clone()
                eng.colmajor_to_rowmajor(K, N, dB_cm, dB, N)   # B produced right before the first SpMM
                for _ in range(5):
                    eng.spmm_device(N, dtype(0.5), dB, N, dtype(0.25), dCw, dCw, N)
                stream.synchronize()
                assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 3
                got = np.ascontiguousarray(dCw.cpu().numpy().T).ravel()
                assert np.array_equal(bits(got), bits(ref)), rep
    finally:
        eng.set_stream(None)


@pytest.mark.parametrize("chains_per_sm", [1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,N", [(20000, 20000, 16), (5000, 6000, 8), (999, 1200, 32), (70, 64, 4), (30000, 30000, 64)])
def test_sliding_window_kernel_bit_exact(eng, chains_per_sm, dtype, M, K, N):
    """SX_OPT_SLIDE + SX_OPT_KERNEL = 4: chains of 32-row steps over a shared-memory ring of B
    rows (or the automatic choice where the ring does not fit) reproduce the oracle."""
    rp, ci, v = banded_csr(M, K, 200, 30, M + N + chains_per_sm, dtype)
    B, Cin = random_dense(M, K, N, M + N, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.set_option(sx.OPT_SLIDE, chains_per_sm)
    eng.set_option(sx.OPT_KERNEL, 4)
    eng.upload_csr(M, K, rp, ci, v)
    for rp_time in (1, 3):
        C = Cin.copy()
        eng.spmm(N, dtype(0.85), B, dtype(-2.06), C, rp_time)
        assert np.array_equal(bits(C), bits(ref))
    if N * np.dtype(dtype).itemsize <= 256:
        assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 7
    # without a plan kernel 4 is the automatic choice
    eng.set_option(sx.OPT_SLIDE, 0)
    eng.upload_csr(M, K, rp, ci, v)
    C = Cin.copy()
    eng.spmm(N, dtype(0.85), B, dtype(-2.06), C)
    assert np.array_equal(bits(C), bits(ref)) and eng.info(sx.INFO_LAST_KERNEL) // 10000 != 7


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["banded", "random"])
def test_autotune_picks_a_variant_and_keeps_the_result(eng, dtype, kind):
    """SX_OPT_AUTOTUNE: the first call times the variants that apply and later calls use the
    fastest; whatever it picks, the result is the oracle's."""
    from helpers import random_csr
    M = K = 6000
    N = 16
    if kind == "banded":
        rp, ci, v = banded_csr(M, K, 150, 20, 5, dtype)
    else:
        rp, ci, v = random_csr(M, K, 25, 5, dtype)
    B, Cin = random_dense(M, K, N, 5, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.set_option(sx.OPT_SLIDE, 1)
    eng.set_option(sx.OPT_AUTOTUNE, 1)
    eng.upload_csr(M, K, rp, ci, v)
    assert eng.info(sx.INFO_TUNED_KERNEL) == 0
    for _ in range(3):
        C = Cin.copy()
        eng.spmm(N, dtype(0.85), B, dtype(-2.06), C)
        assert np.array_equal(bits(C), bits(ref))
    choice = eng.info(sx.INFO_TUNED_KERNEL)
    assert choice // 10 in (1, 2, 3, 4)
    family = eng.info(sx.INFO_LAST_KERNEL) // 10000
    assert family == {1: 1, 2: 2, 3: 3, 4: 7}[choice // 10]
    # a new upload forgets the choice
    eng.upload_csr(M, K, rp, ci, v)
    assert eng.info(sx.INFO_TUNED_KERNEL) == 0
