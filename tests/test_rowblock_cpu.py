"""Host logic of the multi-GPU path on CPU: the nnz-balanced row-block partition, and a
world_size-2 gloo run of the exchange pattern (broadcast of B from rank 0, independent
row blocks, gather of C) with the oracle standing in for the device kernel -- the N>1
plumbing without a GPU.  The GPU version of the same check is in test_multi_gpu.py."""
import os
import socket

import numpy as np
import pytest

import oracle
from helpers import random_csr, random_dense
from sextans_b200.rowblock import PushExchange, RowBlock


def test_row_blocks_tile_the_matrix():
    M, K = 1000, 800
    rp, ci, v = random_csr(M, K, 9, 5, np.float64, long_row=700)
    for world in (1, 2, 3, 8):
        blocks = [RowBlock(M, K, rp, ci, v, world, r) for r in range(world)]
        assert blocks[0].r0 == 0 and blocks[-1].r1 == M
        assert all(blocks[i].r1 == blocks[i + 1].r0 for i in range(world - 1))
        assert sum(b.nnz for b in blocks) == rp[-1]
        assert np.array_equal(np.concatenate([b.colidx for b in blocks]), ci)
        assert np.array_equal(np.concatenate([b.val for b in blocks]), v)
        for b in blocks:
            assert b.rowptr[0] == 0 and b.rowptr[-1] == b.nnz
            assert np.array_equal(np.diff(b.rowptr), np.diff(rp[b.r0:b.r1 + 1]))
        # balance: no block exceeds the ideal share by more than the longest row
        assert max(b.nnz for b in blocks) <= rp[-1] / world + np.diff(rp).max()


def test_push_tree_reaches_every_rank_once():
    """The tree the push exchange sends B down: every rank but the root has exactly one parent that lists
    it as a child at the index it acknowledges to, nobody sends to more than `fanout` ranks, depth is log."""
    for world in (1, 2, 3, 4, 7, 8, 16):
        for root in (0, world - 1):
            for fanout in (1, 2, 3):
                info = {r: PushExchange.tree(world, r, root, fanout) for r in range(world)}
                assert info[root][0] is None and info[root][3] == 0
                seen = []
                for r, (parent, idx, children, depth) in info.items():
                    assert len(children) <= fanout and r not in children
                    seen += children
                    if r != root:
                        assert info[parent][2][idx] == r and depth == info[parent][3] + 1
                assert sorted(seen) == sorted(r for r in range(world) if r != root)
                if fanout == 2:
                    assert max(v[3] for v in info.values()) <= max(0, (world).bit_length() - 1)


def test_take_and_put_C_round_trip():
    M, K, N = 37, 20, 8
    rp, ci, v = random_csr(M, K, 3, 1, np.float32)
    C = np.arange(M * N, dtype=np.float32)
    out = np.zeros_like(C)
    for r in range(3):
        b = RowBlock(M, K, rp, ci, v, 3, r)
        blk = b.take_C(C, N)
        assert blk.size == b.rows * N
        b.put_C(out, blk, N)
    assert np.array_equal(out, C)


def _worker(rank, world, port, seed):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, K, N = 600, 500, 16
        rp, ci, v = random_csr(M, K, 12, seed, np.float64, long_row=400)   # same on every rank
        Bfull, Cin = random_dense(M, K, N, seed, np.float64)
        blk = RowBlock(M, K, rp, ci, v, world, rank)
        # only rank 0 holds B; everyone receives it through the collective
        B = torch.from_numpy(Bfull.copy()) if rank == 0 else torch.zeros(K * N, dtype=torch.float64)
        dist.broadcast(B, src=0)
        assert np.array_equal(B.numpy(), Bfull)
        Cb = blk.take_C(Cin, N)
        oracle.spmm_csr(blk.rows, N, K, blk.rowptr, blk.colidx, blk.val, 0.85, B.numpy(), -2.06, Cb)
        got = [None] * world if rank == 0 else None
        dist.gather_object(Cb, got, dst=0)
        if rank == 0:
            full = np.empty(M * N)
            for r, b in enumerate(got):
                blk.put_C(full, b, N, rank=r)
            ref = oracle.spmm_csr(M, N, K, rp, ci, v, 0.85, Bfull, -2.06, Cin.copy())
            assert np.array_equal(full.view(np.uint64), ref.view(np.uint64))   # bitwise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_broadcast_blocks_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)
