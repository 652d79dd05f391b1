"""GPU parity of variant 5, the edge-list kernel (spmm_edgelist_kernel): a row block's distinct
B rows staged by TMA, 16-bit window-local columns -- the GPU form of the reference's packed
edge words with a window-local column field (src/sparse_helper.h:419-443, src/sextans.cpp:398-402).
Bit-exact against cpu_spmm_CSR in strict mode: one lane group per row, stored order."""
import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import GOLDEN, SUITESPARSE, mtx_path, perturbed_inputs, random_dense, sha

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    e = sx.Engine(0)
    yield e
    e.close()


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def banded_csr(M, K, half_band, per_row, seed, dtype, sort=True):
    rng = np.random.default_rng(seed)
    rp = np.zeros(M + 1, dtype=np.int32)
    cols = []
    for r in range(M):
        lo, hi = max(0, r * K // M - half_band), min(K, r * K // M + half_band + 1)
        n = min(per_row if r % 7 else 0, hi - lo)
        c = rng.choice(np.arange(lo, hi), size=n, replace=False)
        cols.append(np.sort(c) if sort else c)
        rp[r + 1] = rp[r] + n
    ci = np.concatenate(cols).astype(np.int32)
    return rp, ci, rng.uniform(-1, 1, ci.size).astype(dtype)


@pytest.mark.parametrize("name", SUITESPARSE)
def test_auto_takes_edge_lists_on_the_suitesparse_matrices_and_matches_golden(eng, golden, name):
    """C1/C3 through the kernel the auto rule now picks for them: every golden run of the
    reference's cpu_spmm_CSR, bit for bit (SHA-256 of C)."""
    g = golden["suitesparse"][name]
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path(name), np.float32)
    for r in g["runs"]:
        N = r["N"]
        if r["kind"] == "default":
            B, C = oracle.init_dense(M, K, N, np.float32)
            eng.upload_csr(M, K, rp, ci, v)
        else:
            val, B, C = perturbed_inputs(M, K, N, nnz, np.float32)
            eng.upload_csr(M, K, rp, ci, val)
        eng.spmm(N, r["alpha"], B, r["beta"], C, rp_time=2)
        if N * 4 <= 256:
            assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 8, (name, N)
        assert sha(C) == r["C_sha256"], (name, r["kind"], N)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,N,hb,per", [(4704, 4704, 16, 200, 22), (1000, 1000, 8, 40, 9), (999, 1200, 32, 100, 30),
                                          (70, 64, 4, 20, 5), (4000, 4000, 64, 300, 40), (33, 5000, 1, 2000, 60),
                                          (20000, 20000, 24, 30, 12)])
def test_banded_matrices_bit_exact_sorted_and_unsorted(eng, dtype, M, K, N, hb, per):
    for sort in (True, False):
        rp, ci, v = banded_csr(M, K, hb, per, M + N, dtype, sort=sort)
        if not sort and ci.size > 4:
            ci[1] = ci[0]                      # a duplicate (row, column) pair too
        B, Cin = random_dense(M, K, N, M + N, dtype)
        ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
        eng.set_option(sx.OPT_KERNEL, 5)
        eng.upload_csr(M, K, rp, ci, v)
        for rp_time in (1, 3):
            C = Cin.copy()
            eng.spmm(N, dtype(0.85), B, dtype(-2.06), C, rp_time)
            if N * np.dtype(dtype).itemsize <= 256:
                assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 8
            assert np.array_equal(bits(C), bits(ref)), (sort, rp_time)


@pytest.mark.parametrize("pdl", [1, 0])
@pytest.mark.parametrize("prefetch", [1, 0])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dependent_chain_in_place_with_programmatic_launch(eng, dtype, pdl, prefetch):
    """Back-to-back launches whose C_in is the previous launch's C_out (in place) and whose B
    was produced by a kernel just before: the early-started prologue (A slice by TMA, L2
    prefetch of B and C_in) must not consume either before the previous kernel is complete."""
    torch = pytest.importorskip("torch")
    M = K = 4096
    N = 16
    rp, ci, v = banded_csr(M, K, 200, 24, 77, dtype)
    B, Cin = random_dense(M, K, N, 77, dtype)
    ref = Cin.copy()
    for _ in range(6):
        ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.5), B, dtype(0.25), ref)
    eng.set_option(sx.OPT_PDL, pdl)
    eng.set_option(sx.OPT_PREFETCH, prefetch)
    eng.upload_csr(M, K, rp, ci, v)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            dB_cm = torch.from_numpy(B).cuda()
            dC = torch.from_numpy(np.ascontiguousarray(Cin.reshape(N, M).T)).cuda()   # row-major M x N
            dB = torch.empty(K * N, dtype=dB_cm.dtype, device="cuda")
            for rep in range(3):
                dCw = dC.clone()
                dB.zero_()
                eng.colmajor_to_rowmajor(K, N, dB_cm, dB, N)   # B produced right before the first SpMM
                for _ in range(6):
                    eng.spmm_device(N, dtype(0.5), dB, N, dtype(0.25), dCw, dCw, N)
                stream.synchronize()
                assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 8
                got = np.ascontiguousarray(dCw.cpu().numpy().T).ravel()
                assert np.array_equal(bits(got), bits(ref)), rep
            # the same chain as ONE CUDA graph (what the bench replays)
            dCw = dC.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for _ in range(6):
                    eng.spmm_device(N, dtype(0.5), dB, N, dtype(0.25), dCw, dCw, N)
            dCw.copy_(dC)
            g.replay()
            stream.synchronize()
            got = np.ascontiguousarray(dCw.cpu().numpy().T).ravel()
            assert np.array_equal(bits(got), bits(ref))
    finally:
        eng.set_stream(None)


def test_unstructured_matrix_is_left_to_the_other_kernels(eng):
    """No reuse (about one distinct column per nonzero): the auto rule does not build edge lists."""
    from helpers import random_csr
    M, K, N = 20000, 500000, 16
    rp, ci, v = random_csr(M, K, 12, 5, np.float32)
    B, Cin = random_dense(M, K, N, 5, np.float32)
    eng.upload_csr(M, K, rp, ci, v)
    C = Cin.copy()
    eng.spmm(N, np.float32(0.85), B, np.float32(-2.06), C)
    assert eng.info(sx.INFO_LAST_KERNEL) // 10000 in (1, 2)
    assert np.array_equal(bits(C), bits(oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(0.85), B, np.float32(-2.06), Cin.copy())))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,N", [(4704, 4704, 16), (1000, 1000, 8), (997, 1201, 24), (70, 64, 4), (4000, 4000, 32),
                                   (333, 777, 1), (2500, 2500, 64)])
def test_host_facing_call_with_C_carried_by_the_kernel(eng, dtype, M, K, N):
    """sx_spmm_* with page-locked operands and kernel_ns = NULL: B staging + ONE kernel that reads
    C_in from and writes C to the caller's column-major array itself (SX_INFO_HOST_PATH = 2).  Same
    bits as the oracle; with kernel_ns asked for, the three-launch path (1), same bits again."""
    rp, ci, v = banded_csr(M, K, 150, 20, M + N, dtype)
    B, Cin = random_dense(M, K, N, M + N, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.upload_csr(M, K, rp, ci, v)
    hB, hC = sx.pinned_empty(K * N, dtype), sx.pinned_empty(M * N, dtype)
    hB[:] = B
    for rep in range(3):
        hC[:] = Cin
        assert eng.spmm(N, dtype(0.85), hB, dtype(-2.06), hC, want_ns=False) is None
        fused = N * np.dtype(dtype).itemsize <= 256 and (M + K) * N * np.dtype(dtype).itemsize <= (3 << 19)   # SX_OPT_ZEROCOPY_BYTES
        if fused:
            # 3 / family 10: the whole call as one kernel (its grid is resident at once); 2 / family 9: two launches
            assert (eng.info(sx.INFO_HOST_PATH), eng.info(sx.INFO_LAST_KERNEL) // 10000) in ((3, 10), (2, 9))
        assert np.array_equal(bits(np.asarray(hC)), bits(ref)), rep
    hC[:] = Cin
    zero_copy = (M + K) * N * np.dtype(dtype).itemsize <= (3 << 19)
    ns = eng.spmm(N, dtype(0.85), hB, dtype(-2.06), hC)
    assert ns > 0 and eng.info(sx.INFO_HOST_PATH) == (1 if zero_copy else 0) and np.array_equal(bits(np.asarray(hC)), bits(ref))
    eng.set_option(sx.OPT_HOST_FUSED, 0)
    hC[:] = Cin
    eng.spmm(N, dtype(0.85), hB, dtype(-2.06), hC, want_ns=False)
    assert eng.info(sx.INFO_HOST_PATH) == (1 if zero_copy else 0) and np.array_equal(bits(np.asarray(hC)), bits(ref))


def test_host_facing_call_on_the_canned_run(eng, golden):
    """nasa4704 N=16 with the host program's operands through the two-launch path: the golden SHA-256."""
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path("nasa4704"), np.float32)
    run = [r for r in golden["suitesparse"]["nasa4704"]["runs"] if r["kind"] == "default" and r["N"] == 16][0]
    B, Cin = oracle.init_dense(M, K, 16, np.float32)
    eng.upload_csr(M, K, rp, ci, v)
    hB, hC = sx.pinned_empty(K * 16, np.float32), sx.pinned_empty(M * 16, np.float32)
    hB[:] = B
    hC[:] = Cin
    for fused, path in ((-1, 3), (2, 3), (1, 2)):
        eng.set_option(sx.OPT_HOST_FUSED, fused)
        hC[:] = Cin
        eng.spmm(16, run["alpha"], hB, run["beta"], hC, want_ns=False)
        assert eng.info(sx.INFO_HOST_PATH) == path
        assert sha(np.asarray(hC)) == run["C_sha256"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fused", [1, 2])
@pytest.mark.parametrize("groups", [1, 2, 3, 4, 8, 64])
@pytest.mark.parametrize("M,K,N", [(4704, 4704, 16), (997, 1201, 24), (333, 777, 1), (1500, 1400, 10), (9000, 300, 16), (200, 9000, 12)])
def test_host_facing_call_column_pipeline(eng, dtype, fused, groups, M, K, N):
    """SX_OPT_HOST_GROUPS: the fused host-facing call pipelined over column groups of the caller's
    column-major arrays -- inside ONE kernel (SX_OPT_HOST_FUSED = 2: cp.async from host memory, the
    blocks meet behind a counter per group) or as a chain of (B staging, SpMM) pairs (= 1).  Columns
    are independent: the same bits for every group count, including counts that do not divide N and
    more groups than columns; tall and wide shapes (B shares of one row / of many rows per block)."""
    rp, ci, v = banded_csr(M, K, 150, 20, M + N, dtype)
    B, Cin = random_dense(M, K, N, M + N + 1, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.set_option(sx.OPT_HOST_GROUPS, groups)
    eng.set_option(sx.OPT_HOST_FUSED, fused)
    eng.upload_csr(M, K, rp, ci, v)
    hB, hC = sx.pinned_empty(K * N, dtype), sx.pinned_empty(M * N, dtype)
    hB[:] = B
    for rep in range(3):
        hC[:] = Cin
        eng.spmm(N, dtype(0.85), hB, dtype(-2.06), hC, want_ns=False)
        # (1: the matrix does not take the edge-list kernel at all -- wide, hardly any reuse of a staged B row)
        assert eng.info(sx.INFO_HOST_PATH) in ((3 if fused == 2 else 2), 1)
        assert K > 4 * M or eng.info(sx.INFO_HOST_PATH) != 1
        assert np.array_equal(bits(np.asarray(hC)), bits(ref)), rep
    assert eng.info(sx.INFO_EXCHANGE_TIMEOUTS) == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", [0, 2])
@pytest.mark.parametrize("M,K,N,nb", [(4704, 4704, 16, 5), (997, 1201, 8, 3), (70, 64, 4, 17), (600, 600, 40, 2)])
def test_batched_call_matches_one_call_per_operand(eng, dtype, kernel, M, K, N, nb):
    """sx_spmm_device_batch_*: nb (B, C_in, C_out) triples with the same A -- one launch on the edge-list
    kernel (kernel 0 on these banded matrices), one launch per triple otherwise -- every result
    bit-identical to cpu_spmm_CSR; in place as well."""
    import torch
    rp, ci, v = banded_csr(M, K, 150, 20, M + N, dtype)
    eng.set_option(sx.OPT_KERNEL, kernel)
    eng.upload_csr(M, K, rp, ci, v)
    ld = (N + 7) // 8 * 8
    td = torch.float64 if dtype == np.float64 else torch.float32
    dev = torch.device("cuda", 0)
    sB, sC = K * ld + 8, M * ld + 16                      # strides with slack between operands
    dB = torch.zeros(nb * sB, dtype=td, device=dev)
    dCin = torch.zeros(nb * sC, dtype=td, device=dev)
    dCout = torch.full((nb * sC,), 7.0, dtype=td, device=dev)
    refs = []
    for b in range(nb):
        B, Cin = random_dense(M, K, N, 100 * b + N, dtype)
        refs.append(oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy()))
        Brm = np.zeros((K, ld), dtype); Brm[:, :N] = B.reshape(N, K).T
        Crm = np.zeros((M, ld), dtype); Crm[:, :N] = Cin.reshape(N, M).T
        dB[b * sB: b * sB + K * ld] = torch.from_numpy(Brm.ravel()).to(dev)
        dCin[b * sC: b * sC + M * ld] = torch.from_numpy(Crm.ravel()).to(dev)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    l0 = eng.info(sx.INFO_LAUNCHES)
    eng.spmm_device_batch(N, nb, dtype(0.85), dB, ld, sB, dtype(-2.06), dCin, dCout, ld, sC)
    torch.cuda.synchronize()
    if kernel == 0 and N * np.dtype(dtype).itemsize <= 256:
        assert eng.info(sx.INFO_LAST_KERNEL) // 10000 == 8 and eng.info(sx.INFO_LAUNCHES) - l0 == 1
    out = dCout.cpu().numpy()
    for b in range(nb):
        got = out[b * sC: b * sC + M * ld].reshape(M, ld)[:, :N].T.ravel()
        assert np.array_equal(bits(np.ascontiguousarray(got)), bits(refs[b])), b
        assert np.all(out[b * sC + M * ld: (b + 1) * sC] == 7.0)      # the slack between operands is untouched
    eng.spmm_device_batch(N, nb, dtype(0.85), dB, ld, sB, dtype(-2.06), dCin, dCin, ld, sC)   # in place
    torch.cuda.synchronize()
    out = dCin.cpu().numpy()
    for b in range(nb):
        got = out[b * sC: b * sC + M * ld].reshape(M, ld)[:, :N].T.ravel()
        assert np.array_equal(bits(np.ascontiguousarray(got)), bits(refs[b])), b
    with pytest.raises(sx.SextansError):
        eng.spmm_device_batch(N, 2, dtype(0.85), dB, ld, sB, dtype(-2.06), dCin, dCout, ld, 0)


@pytest.mark.parametrize("kernel", [0, 2])
def test_first_call_inside_a_graph_capture_is_refused_not_broken(eng, kernel):
    """The first SpMM of a (matrix, N) builds its plan on the host (allocations, a sync): inside a
    stream capture the library says so (SX_ERR_STATE) before touching the stream, the capture stays
    valid, and after one call outside a capture the same launch is captured and replayed bit-exact."""
    import torch
    M, K, N, dtype = 1000, 1000, 16, np.float64
    rp, ci, v = banded_csr(M, K, 150, 20, 3, dtype)
    B, Cin = random_dense(M, K, N, 4, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.set_option(sx.OPT_KERNEL, kernel)
    eng.upload_csr(M, K, rp, ci, v)
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    eng.set_stream(s.cuda_stream)
    ld = 16
    dB = torch.from_numpy(np.ascontiguousarray(B.reshape(N, K).T)).to(dev)
    dCin = torch.from_numpy(np.ascontiguousarray(Cin.reshape(N, M).T)).to(dev)
    dCout = torch.zeros(M * ld, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        with pytest.raises(sx.SextansError, match="cannot be captured"):
            eng.spmm_device(N, 0.85, dB, ld, -2.06, dCin, dCout, ld)
    with torch.cuda.stream(s):
        eng.spmm_device(N, 0.85, dB, ld, -2.06, dCin, dCout, ld)
    s.synchronize()
    dCout.zero_()
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        eng.spmm_device(N, 0.85, dB, ld, -2.06, dCin, dCout, ld)
    g2.replay()
    torch.cuda.synchronize()
    got = dCout.cpu().numpy().reshape(M, ld)[:, :N].T.ravel()
    assert np.array_equal(bits(np.ascontiguousarray(got)), bits(ref))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("M,K,N", [(4704, 4704, 16), (997, 1201, 24), (333, 777, 1), (200, 9000, 12)])
def test_spmm_on_the_staged_B_image_with_host_C(eng, dtype, pinned, M, K, N):
    """sx_spmm_staged_B_*: B already on the device (sx_stage_B_*), C the caller's host array -- one
    launch with C carried by the kernel when C is page-locked and the matrix runs the edge-list
    kernel, staged / multiplied / fetched otherwise.  Bitwise the oracle either way; B stays staged."""
    rp, ci, v = banded_csr(M, K, 150, 20, M + N, dtype)
    B, Cin = random_dense(M, K, N, M + N + 2, dtype)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
    eng.upload_csr(M, K, rp, ci, v)
    with pytest.raises(sx.SextansError):
        eng.spmm_staged_B(N, dtype(0.85), dtype(-2.06), Cin.copy())     # no B image yet
    eng.stage_B(N, B)
    hC = sx.pinned_empty(M * N, dtype) if pinned else np.empty(M * N, dtype)
    for rep in range(2):
        hC[:] = Cin
        eng.spmm_staged_B(N, dtype(0.85), dtype(-2.06), hC)
        assert np.array_equal(bits(np.asarray(hC)), bits(ref)), rep
        if pinned and K <= 4 * M:
            assert eng.info(sx.INFO_HOST_PATH) == 2 and eng.info(sx.INFO_LAST_KERNEL) // 10000 == 9
        if not pinned:
            assert eng.info(sx.INFO_HOST_PATH) == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_calls_in_flight_on_two_contexts(dtype):
    """sx_spmm_enqueue_*: the host-facing call without its final host sync.  Two contexts (two streams) with
    double-buffered page-locked operands keep two calls in flight -- one call's results leave over PCIe
    while the next call's operands arrive; sx_synchronize hands a buffer back.  Every step bitwise the oracle."""
    M, K, N = 4704, 4704, 16
    rp, ci, v = banded_csr(M, K, 150, 20, 11, dtype)
    engs = [sx.Engine(0), sx.Engine(0)]
    try:
        for e in engs:
            e.upload_csr(M, K, rp, ci, v)
        hB = [sx.pinned_empty(K * N, dtype) for _ in range(2)]
        hC = [sx.pinned_empty(M * N, dtype) for _ in range(2)]
        refs = [None, None]
        for step in range(9):
            j = step % 2
            if step >= 2:
                engs[j].synchronize()                                   # the call that used buffer j two steps ago is complete
                assert engs[j].info(sx.INFO_HOST_PATH) == 3
                assert np.array_equal(bits(np.asarray(hC[j])), bits(refs[j])), step - 2
            B, Cin = random_dense(M, K, N, 300 + step, dtype)
            hB[j][:] = B
            hC[j][:] = Cin
            refs[j] = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
            engs[j].spmm_enqueue(N, dtype(0.85), hB[j], dtype(-2.06), hC[j])
        for j in range(2):
            engs[j].synchronize()
            assert np.array_equal(bits(np.asarray(hC[j])), bits(refs[j]))
    finally:
        for e in engs:
            e.close()
