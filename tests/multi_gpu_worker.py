"""torchrun worker for test_multi_gpu.py: row-block SpMM over WORLD_SIZE GPUs through
ShardedSpMM (NCCL broadcast of B), checked bitwise against the oracle on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle  # noqa: E402
from helpers import mtx_path, perturbed_inputs, random_csr, random_dense  # noqa: E402
import sextans_b200 as sx  # noqa: E402
from sextans_b200.rowblock import ShardedSpMM  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = []
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path("pcrystk02"), np.float32)
    val, B, Cin = perturbed_inputs(M, K, 16, nnz, np.float32)
    cases.append(("pcrystk02 f32 N=16", M, K, 16, rp, ci, val, B, Cin))
    M, K, N = 5000, 4000, 32
    rp, ci, v = random_csr(M, K, 14, 3, np.float64, long_row=3000)
    B, Cin = random_dense(M, K, N, 3, np.float64)
    cases.append(("random f64 N=32 with a long row", M, K, N, rp, ci, v, B, Cin))
    for name, M, K, N, rp, ci, v, B, Cin in cases:
        dtype = v.dtype.type
        for peer_bytes, want in ((8 << 20, "push"), (0, "nccl")):      # both ways of moving B
            sh = ShardedSpMM(M, K, rp, ci, v, local, peer_bytes=peer_bytes)
            sh.engine.set_option(sx.OPT_SPLIT_ROW_NNZ, 0)          # everything in stored order: bitwise
            for rep in range(2):                                       # twice: B is re-staged and re-sent
                Cb = sh.block.take_C(Cin, N)
                ns = sh.spmm(N, dtype(0.85), B if rank == 0 else None, dtype(-2.06), Cb, src=0, rp_time=2)
                assert sh.last_exchange == want, (sh.last_exchange, want)
                full = sh.gather(Cb, N, dst=0)
                if rank == 0:
                    ref = oracle.spmm_csr(M, N, K, rp, ci, v, dtype(0.85), B, dtype(-2.06), Cin.copy())
                    assert full.tobytes() == ref.tobytes(), (name, want)
            if rank == 0:
                print(f"OK {name} via {want}: {world} row blocks == oracle bitwise; rank-0 kernel {ns / 2e3:.1f} us", flush=True)
            sh.close()
    # the push exchange on its own (PushExchange): several steps with a B that changes each time,
    # R = 3 images used round-robin, launches captured in a CUDA graph and replayed (the counters
    # live in device memory), and a different N afterwards through ShardedSpMM (the exchange is rebuilt)
    from sextans_b200.rowblock import PushExchange, RowBlock
    M, K, N = 4000, 3000, 16
    rp, ci, v = random_csr(M, K, 10, 5, np.float64)
    blk = RowBlock(M, K, rp, ci, v, world, rank)
    engs = []
    stream = torch.cuda.Stream()
    for _ in range(3):
        e = sx.Engine(local)
        e.set_stream(stream.cuda_stream)
        e.upload_csr(blk.rows, K, blk.rowptr, blk.colidx, blk.val)
        engs.append(e)
    px = PushExchange(engs, N)
    ld = 16
    for k in range(7):
        j = k % 3
        eng = engs[j]
        B, Cin = random_dense(M, K, N, 100 + k, np.float64)   # every rank can rebuild the root's B to check
        Cb = blk.take_C(Cin, N)
        if rank == 0:
            eng.stage_B(N, B)
        else:
            eng.device_B(N)
        px.before_step(k)
        eng.stage_C(N, Cb)
        eng.launch(0.85, -2.06)
        px.flush()                       # deferred publication (3 images in rotation): a host sync ends the sequence
        eng.fetch_C(Cb)
        ref = oracle.spmm_csr(blk.rows, N, K, blk.rowptr, blk.colidx, blk.val, 0.85, B, -2.06, blk.take_C(Cin, N))
        assert Cb.tobytes() == ref.tobytes(), (rank, k)
        assert eng.info(sx.INFO_EXCHANGE_TIMEOUTS) == 0
    # graph replay: the same three launches (and, on the root, pushes) replayed four times
    dCin = [torch.zeros(blk.rows * ld, dtype=torch.float64, device="cuda") for _ in range(3)]
    dCout = [torch.zeros(blk.rows * ld, dtype=torch.float64, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for k in range(3):
            px.before_step(k)
            engs[k].spmm_device(N, 0.85, engs[k].device_B(N)[0], ld, -2.06, dCin[k], dCout[k], ld)
        px.flush()                       # the last step's publication is part of the graph
    assert px.defer
    for _ in range(4):
        with torch.cuda.stream(stream):
            g.replay()
    torch.cuda.synchronize()
    assert all(e.info(sx.INFO_EXCHANGE_TIMEOUTS) == 0 for e in engs)
    dist.barrier()
    px.close()
    if rank == 0:
        print("OK push of B through peer memory: 7 steps over 3 images bitwise on every rank, 4 graph replays without a time-out", flush=True)
    for e in engs[1:]:
        e.close()
    eng = engs[0]
    # a different N on the same ShardedSpMM: the B image moves, the exchange is rebuilt, no stale mapping
    M, K = 3000, 2500
    rp, ci, v = random_csr(M, K, 9, 8, np.float32)
    sh = ShardedSpMM(M, K, rp, ci, v, local)
    for N in (8, 40, 8):
        B, Cin = random_dense(M, K, N, 50 + N, np.float32)
        Cb = sh.block.take_C(Cin, N)
        sh.spmm(N, np.float32(0.85), B if rank == 0 else None, np.float32(-2.06), Cb)
        assert sh.last_exchange == "push"
        full = sh.gather(Cb, N, dst=0)
        if rank == 0:
            ref = oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(0.85), B, np.float32(-2.06), Cin.copy())
            assert full.tobytes() == ref.tobytes(), N
    sh.close()
    if rank == 0:
        print("OK ShardedSpMM with N = 8, 40, 8: the push exchange follows the B image, bitwise", flush=True)
    # the host-facing call on every rank (want_ns=False): the holder's sx_spmm_* carries the push and its C block,
    # the others' sx_spmm_staged_B_* waits for the push and carries theirs; page-locked operands, FEM-type matrix
    M, K, nnz, rp, ci, v = sx.load_mtx(mtx_path("nasa4704"), np.float64)
    sh = ShardedSpMM(M, K, rp, ci, v, local)
    N = 16
    hB = sx.pinned_empty(K * N, np.float64)
    hC = sx.pinned_empty(sh.block.rows * N, np.float64)
    for rep in range(4):
        B, Cin = random_dense(M, K, N, 70 + rep, np.float64)
        hB[:] = B
        hC[:] = sh.block.take_C(Cin, N)
        assert sh.spmm(N, 0.85, hB if rank == 0 else None, -2.06, hC, want_ns=False) is None
        assert sh.last_exchange == "push" and sh.engine.info(sx.INFO_HOST_PATH) == 2
        assert sh.engine.info(sx.INFO_EXCHANGE_TIMEOUTS) == 0
        full = sh.gather(np.asarray(hC).copy(), N, dst=0)
        if rank == 0:
            ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, Cin.copy())
            assert full.tobytes() == ref.tobytes(), rep
    sh.close()
    if rank == 0:
        print("OK ShardedSpMM host-facing call on every rank (push carried by the holder's SpMM, C by every rank's kernel): bitwise", flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
