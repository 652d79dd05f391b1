"""Row-block partition of one SpMM across the GPUs of a box.

Rows of A and C are independent (the reference itself deals rows to its 64 PEs,
src/sparse_helper.h:370), so every rank owns a contiguous block of rows holding about
nnz/world nonzeros, ALL of B, and the matching block of C.  The only exchange step is
the broadcast of B from the rank that has it -- the multi-GPU form of the reference's
daisy chain that hands the B window from PEG to PEG (src/sextans.cpp:909-941).  There
are no reductions, and a row's arithmetic does not depend on the partition, so the
result is bitwise the single-GPU result.

``RowBlock`` is pure host logic (no GPU, no torch); ``ShardedSpMM`` drives one
``Engine`` per process with ``torch.distributed`` (NCCL over NVLink on the GPUs).
"""
from __future__ import annotations

import numpy as np

from . import Engine, partition_rows


class RowBlock:
    """This rank's rows [r0, r1) of a CSR matrix, with the row pointers rebased to 0."""

    def __init__(self, M, K, rowptr, colidx, val, world, rank):
        if not 0 <= rank < world:
            raise ValueError("rank out of range")
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.M, self.K, self.world, self.rank = int(M), int(K), int(world), int(rank)
        self.bounds = partition_rows(rowptr, world)
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        j0, j1 = int(rowptr[self.r0]), int(rowptr[self.r1])
        self.rows = self.r1 - self.r0
        self.rowptr = (rowptr[self.r0:self.r1 + 1] - rowptr[self.r0]).astype(np.int32)
        self.colidx = np.ascontiguousarray(colidx[j0:j1], dtype=np.int32)
        self.val = np.ascontiguousarray(val[j0:j1])
        self.nnz = j1 - j0

    def take_C(self, C_colmajor, N):
        """This rank's block of a column-major M x N operand, as its own column-major array."""
        return np.ascontiguousarray(C_colmajor.reshape(N, self.M)[:, self.r0:self.r1]).ravel()

    def put_C(self, C_colmajor, block, N, rank=None):
        """Write a rank's block back into the full column-major M x N array."""
        r = self.rank if rank is None else rank
        r0, r1 = int(self.bounds[r]), int(self.bounds[r + 1])
        C_colmajor.reshape(N, self.M)[:, r0:r1] = np.asarray(block).reshape(N, r1 - r0)
        return C_colmajor


class PeerBroadcast:
    """B from the root rank's engine(s) to every other rank WITHOUT a collective: each
    rank pulls the root's row-major B image with a copy-engine peer copy over NVLink, and
    the ordering is carried by 32-bit step counters in peer-mapped device memory, written
    and waited on by stream memory operations (sx_flag_write / sx_flag_wait) -- no kernel,
    no host round trip.  For the SuiteSparse-sized configs this replaces ~60 us of NCCL
    launch latency per broadcast by a ~10 us peer copy.

    ``engines``: this rank's Engine objects (one, or several replicas used round-robin),
    each with A uploaded.  Step numbers start at 1 and must increase by one per call.
      root : publish(k)      after B of replica (k-1) % R is staged
             reclaim(k)      before that replica's B is overwritten again (waits until every
                             peer has finished pulling step k)
      peers: pull(k)         waits for publish(k), copies, acknowledges
    """

    def __init__(self, engines, N, group=None, root=0, fused=True):
        import torch.distributed as dist
        self.engines, self.N, self.root, self.fused, self.group = list(engines), N, root, fused, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.R = len(self.engines)
        e0 = self.engines[0]
        self._imported = []
        if self.rank == root:
            self.done = e0.device_alloc(4 * self.world)                 # done[r]: last step rank r pulled
            mine = {"images": [e0_.ipc_export(e0_.device_B(N)[0]) for e0_ in self.engines],
                    "done": e0.ipc_export(self.done)}
        else:
            self.ready = e0.device_alloc(4)                              # last step the root published
            mine = {"ready": e0.ipc_export(self.ready)}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if self.rank == root:
            self.peer_ready = {}
            for r, obj in enumerate(everyone):
                if r != root:
                    self.peer_ready[r] = e0.ipc_import(obj["ready"])
                    self._imported.append(self.peer_ready[r])
        else:
            self.root_images = [e0.ipc_import(h) for h in everyone[root]["images"]]
            self.root_done = e0.ipc_import(everyone[root]["done"])
            self._imported += self.root_images + [self.root_done]
        dist.barrier(group=group)

    def publish(self, k):
        e = self.engines[(k - 1) % self.R]
        ptrs = list(self.peer_ready.values())
        for i in range(0, len(ptrs), 16):              # one small kernel per 16 peers
            e.flag_write_many(ptrs[i:i + 16], k)

    def reclaim(self, k):
        e = self.engines[(k - 1) % self.R]
        for r in self.peer_ready:
            e.flag_wait(self.done + 4 * r, k)

    def pull(self, k):
        j = (k - 1) % self.R
        e = self.engines[j]
        if self.fused:      # one kernel: spin on the local flag, copy over NVLink, acknowledge
            e.pull_B_fused(self.N, self.root_images[j], self.ready, self.root_done + 4 * self.rank, k)
        else:               # stream memory operations around a copy-engine peer copy
            e.flag_wait(self.ready, k)
            e.pull_B(self.N, self.root_images[j])
            e.flag_write(self.root_done + 4 * self.rank, k)

    def close(self):
        """Collective: every rank unmaps what it imported, then the owners free their flags."""
        import torch.distributed as dist
        e0 = self.engines[0]
        e0.synchronize()
        for ptr in self._imported:
            e0.ipc_close(ptr)
        self._imported = []
        dist.barrier(group=self.group)
        for name in ("done", "ready"):
            ptr = getattr(self, name, None)
            if ptr:
                e0.device_free(ptr)
                setattr(self, name, None)


class ShardedSpMM:
    """One process per GPU.  ``spmm`` moves B from the rank that has it to the others --
    by peer copy (PeerBroadcast) when the B image is at most ``peer_bytes``, by one NCCL
    broadcast otherwise -- and runs the local row block; C stays sharded unless ``gather``
    is asked for."""

    def __init__(self, M, K, rowptr, colidx, val, device, group=None, arith=0, peer_bytes=8 << 20):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.block = RowBlock(M, K, rowptr, colidx, val, self.world, self.rank)
        self.engine = Engine(device, arith=arith)
        # one stream for the engine's kernels AND the collective, so that they are ordered
        self.stream = torch.cuda.Stream(device=device)
        self.engine.set_stream(self.stream.cuda_stream)
        self.engine.upload_csr(self.block.rows, K, self.block.rowptr, self.block.colidx, self.block.val)
        self.dtype = self.block.val.dtype
        self.device = device
        self.peer_bytes = peer_bytes
        self._peer = None        # (N, src, PeerBroadcast) once set up
        self._peer_failed = False
        self._step = 0
        self.last_exchange = None

    def close(self):
        if self._peer is not None:
            self._peer[2].close()
            self._peer = None
        self.engine.close()

    def device_B(self, N):
        """torch view of the engine's row-major B image [K, ld] (the broadcast target)."""
        import torch
        ptr, nbytes = self.engine.device_B(N)
        ld = self.engine.info(7)
        n = nbytes // self.dtype.itemsize

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8" if self.dtype == np.float64 else "<f4",
                                        "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Arr(), device=f"cuda:{self.device}").view(self.block.K, ld)

    def _peer_for(self, N, src):
        if self._peer is not None and self._peer[:2] == (N, src):
            return self._peer[2]
        if self._peer is not None or self._peer_failed or self.world == 1:
            return None              # a different N / root than the one the handles were made for
        _, nbytes = self.engine.device_B(N)
        ok = nbytes <= self.peer_bytes
        flags = [None] * self.world
        self.dist.all_gather_object(flags, bool(ok), group=self.group)
        if not all(flags):
            self._peer_failed = True
            return None
        try:
            self._peer = (N, src, PeerBroadcast([self.engine], N, self.group, root=src))
        except Exception:
            self._peer_failed = True
            return None
        return self._peer[2]

    def spmm(self, N, alpha, B_colmajor_root, beta, C_block_colmajor, src=0, rp_time=1):
        """B_colmajor_root: the K x N column-major host B on rank ``src`` (ignored elsewhere).
        C_block_colmajor: this rank's block (in/out).  Returns the local kernel ns."""
        import torch
        pb = self._peer_for(N, src)
        if pb is not None:
            self._step += 1
            if self.rank == src:
                if self._step > 1:
                    pb.reclaim(self._step - 1)             # every peer has finished with the previous B
                self.engine.stage_B(N, B_colmajor_root)    # H2D + layout change on the root only
                pb.publish(self._step)
            else:
                pb.pull(self._step)
            self.last_exchange = "peer"
        else:
            if self.rank == src:
                self.engine.stage_B(N, B_colmajor_root)
            dB = self.device_B(N)
            with torch.cuda.stream(self.stream):
                self.dist.broadcast(dB, src=src, group=self.group)  # NCCL over NVLink / NVSwitch
            self.last_exchange = "nccl"
        self.engine.stage_C(N, C_block_colmajor)
        ns = self.engine.launch(alpha, beta, rp_time)
        self.engine.fetch_C(C_block_colmajor)
        return ns

    def gather(self, C_block_colmajor, N, dst=0):
        """Collect the blocks on rank ``dst`` -> full column-major C there, None elsewhere."""
        blocks = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(C_block_colmajor, blocks, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        out = np.empty(self.block.M * N, dtype=self.dtype)
        for r, b in enumerate(blocks):
            self.block.put_C(out, b, N, rank=r)
        return out
