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
        for peer_bytes, want in ((8 << 20, "peer"), (0, "nccl")):      # both ways of moving B
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
    # the same exchange without a collective: peers pull the root's B over NVLink, ordered by
    # device-side step counters (PeerBroadcast); several steps with a B that changes each time
    from sextans_b200.rowblock import PeerBroadcast, RowBlock
    M, K, N = 4000, 3000, 16
    rp, ci, v = random_csr(M, K, 10, 5, np.float64)
    blk = RowBlock(M, K, rp, ci, v, world, rank)
    eng = sx.Engine(local)
    eng.upload_csr(blk.rows, K, blk.rowptr, blk.colidx, blk.val)
    eng.device_B(N)
    for fused in (True, False):      # one fused kernel per pull / stream memory ops around a peer copy
        pb = PeerBroadcast([eng], N, fused=fused)
        for k in range(1, 6):
            B, Cin = random_dense(M, K, N, 100 + k + 10 * fused, np.float64)   # every rank can rebuild the root's B to check
            Cb = blk.take_C(Cin, N)
            if rank == 0:
                if k > 1:
                    pb.reclaim(k - 1)                                  # peers are done with the previous B
                eng.stage_B(N, B)
                pb.publish(k)
            else:
                pb.pull(k)
            eng.stage_C(N, Cb)
            eng.launch(0.85, -2.06)
            eng.fetch_C(Cb)
            ref = oracle.spmm_csr(blk.rows, N, K, blk.rowptr, blk.colidx, blk.val, 0.85, B, -2.06, blk.take_C(Cin, N))
            assert Cb.tobytes() == ref.tobytes(), (rank, k, fused)
        if rank == 0:
            pb.reclaim(5)
        eng.synchronize()
        dist.barrier()
        pb.close()
    if rank == 0:
        print("OK peer pull of B over NVLink (fused kernel and memory-op variants): 5 steps each, bitwise on every rank", flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
