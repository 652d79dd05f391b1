"""Row-block partition of one SpMM across the GPUs of a box.

Rows of A and C are independent (the reference itself deals rows to its 64 PEs,
src/sparse_helper.h:370), so every rank owns a contiguous block of rows holding about
nnz/world nonzeros, ALL of B, and the matching block of C.  The only exchange step is
the broadcast of B from the rank that has it -- the multi-GPU form of the reference's
daisy chain that hands the B window from PEG to PEG (src/sextans.cpp:909-941).  There
are no reductions, and a row's arithmetic does not depend on the partition, so the
result is bitwise the single-GPU result.

``RowBlock`` is pure host logic (no GPU, no torch); ``ShardedSpMM`` drives one
``Engine`` per process with ``torch.distributed`` (NCCL over NVLink on the GPUs);
``PushExchange`` is the exchange of a small B through peer memory.
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


class PushExchange:
    """B from the root rank to every other rank with NOTHING launched for the exchange on any
    rank: a rank that holds B has its SpMM kernel itself copy the B image it reads into its
    children's images on the way to its rows (sx_spmm_fuse_push: posted 16-byte stores over
    NVLink through CUDA-IPC peer mappings, the step published by a one-warp dependent kernel),
    and a receiving rank's SpMM waits for the push in its own prologue and acknowledges it from
    its last block (sx_spmm_expect_push).  Ranks form a binary tree under the root (children of
    position i: 2i+1, 2i+2), so no GPU sends more than two images per step -- an inner rank's
    kernel waits for its parent's push and forwards the image to its own children in the same
    launch; depth adds latency to the pipeline, not to its period.  All counters are 32-bit words
    in device memory -- per image: ``ready`` and ``epoch`` on a receiver, ``pushes`` and
    ``done[child]`` on a sender -- so captured launches can be replayed.  The multi-GPU form of
    the reference's chain that hands the B window from PEG to PEG (src/sextans.cpp:909-941).

    ``engines``: this rank's Engine objects (one, or R replicas used round-robin), each with A
    uploaded; their own B images (``Engine.device_B(N)``) are the operands.  Step i uses image
    i % R: call ``before_step(i)`` right before the SpMM of step i is enqueued on engine i % R.
    """
    WORDS = 8192          # mailbox: ready[j] at j, epoch[j] at 256 + j, pushes[j] at 512 + j, done[j][c] at 1024 + 16 j + c

    def __init__(self, engines, N, group=None, root=0, fanout=2, defer_publish=None):
        import torch.distributed as dist
        self.engines, self.N, self.root, self.group = list(engines), N, root, group
        # deferred publication (R >= 2 images in rotation on ONE stream): the step is published by the NEXT step's SpMM
        # kernel right after its dependent-launch wait instead of by a one-warp kernel of its own, which would gate
        # that wait (~1.1 us per step); the caller ends every sequence of steps with flush()
        self.defer = (len(self.engines) >= 2) if defer_publish is None else (bool(defer_publish) and len(self.engines) >= 2)
        self._owed = None                      # (engine index, children's ready flags, pushes counter) of the last step
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.R = len(self.engines)
        if self.R > 256 or self.world > 16 or not 1 <= fanout <= 15:
            raise ValueError("PushExchange: at most 256 images, 16 ranks, fan-out 1..15")
        self.parent, self.child_index, self.children, self.depth = self.tree(self.world, self.rank, root, fanout)
        e0 = self.engines[0]
        self.images, self.nbytes = [], None
        for e in self.engines:
            ptr, nb = e.device_B(N)
            self.images.append(ptr)
            self.nbytes = nb
        self.box = e0.device_alloc(4 * self.WORDS)
        mine = {"box": e0.ipc_export_ref(self.box), "images": [e0.ipc_export_ref(p) for p in self.images]}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = {}                      # IPC handle -> base address in this process (a handle is opened once)
        self.child_box = {r: self._open(everyone[r]["box"]) for r in self.children}
        self.child_images = {r: [self._open(ref) for ref in everyone[r]["images"]] for r in self.children}
        self.parent_box = None if self.parent is None else self._open(everyone[self.parent]["box"])
        dist.barrier(group=group)

    @staticmethod
    def tree(world, rank, root=0, fanout=2):
        """(parent, index among the parent's children, children, depth) of ``rank``: the root is
        position 0, the other ranks follow in rank order; children of position i are fanout*i+1 ..."""
        order = [root] + [r for r in range(world) if r != root]
        pos = order.index(rank)
        parent = None if pos == 0 else order[(pos - 1) // fanout]
        child_index = None if pos == 0 else (pos - 1) % fanout
        children = [order[c] for c in range(fanout * pos + 1, fanout * pos + 1 + fanout) if c < world]
        depth = 0
        while pos:
            pos, depth = (pos - 1) // fanout, depth + 1
        return parent, child_index, children, depth

    def _open(self, ref):
        handle, offset = ref
        if handle not in self._opened:
            self._opened[handle] = self.engines[0].ipc_import(handle)
        return self._opened[handle] + offset

    def describe(self):
        return (f"B ({self.nbytes / 1e3:.0f} KB) pushed down a binary tree of ranks over NVLink BY the SpMM kernels themselves (a sender's "
                "blocks copy their share of the image into its <= 2 children's images with posted peer stores; the step is published "
                + ("by the NEXT step's SpMM kernel right after its dependent-launch wait (deferred publication: no kernel of its own in the chain)"
                   if self.defer else "by a one-warp dependent kernel") +
                "; a receiver's kernel waits on the step flag in its prologue, forwards if it has children, "
                "and acknowledges from its last block): no stream and no collective for the exchange on any rank")

    def before_step(self, i):
        from . import INFO_PUSH_PENDING
        j = i % self.R
        e = self.engines[j]
        if self.parent is not None:
            e.expect_push(self.box + 4 * j, self.box + 4 * (256 + j), self.parent_box + 4 * (1024 + 16 * j + self.child_index))
        if self.children:
            ready = [self.child_box[r] + 4 * j for r in self.children]
            args = ([self.child_images[r][j] for r in self.children], ready, self.box + 4 * (1024 + 16 * j), self.box + 4 * (512 + j))
            if not self.defer:
                e.fuse_push(*args)
                return
            if self._owed is not None:
                jp, rp, pp = self._owed
                self._owed = None
                if self.engines[jp].info(INFO_PUSH_PENDING):
                    if jp == j:                # the same image twice in a row: its own launch cannot publish for it
                        self.engines[jp].push_publish(rp, pp)
                    else:
                        e.fuse_publish(rp, pp)
            e.fuse_push_deferred(*args)
            self._owed = (j, ready, self.box + 4 * (512 + j))

    def flush(self):
        """Publish the last step's push if that is still owed (deferred publication): before a host
        sync, at the end of a captured graph, before another stream takes over."""
        from . import INFO_PUSH_PENDING
        if self._owed is not None:
            jp, rp, pp = self._owed
            self._owed = None
            if self.engines[jp].info(INFO_PUSH_PENDING):
                self.engines[jp].push_publish(rp, pp)

    def close(self):
        """Collective: every rank unmaps what it imported, then frees its mailbox."""
        import torch.distributed as dist
        e0 = self.engines[0]
        self.flush()
        for e in self.engines:
            e.synchronize()
        dist.barrier(group=self.group)
        for base in self._opened.values():
            e0.ipc_close(base)
        self._opened = {}
        dist.barrier(group=self.group)
        if self.box:
            e0.device_free(self.box)
            self.box = None


class ShardedSpMM:
    """One process per GPU.  ``spmm`` moves B from the rank that has it to the others --
    pushed through peer memory (PushExchange) when the B image is at most ``peer_bytes``, by one
    NCCL broadcast otherwise -- and runs the local row block; C stays sharded unless ``gather``
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
        self.K = K
        self.device = device
        self.peer_bytes = peer_bytes
        self._xch = None         # (N, src, image address, PushExchange) once set up
        self._xch_failed = False
        self.last_exchange = None

    @classmethod
    def around(cls, engine, dtype, rows, K, stream, device, group=None, peer_bytes=8 << 20):
        """Wrap an Engine that already holds this rank's row block (no second upload)."""
        import torch.distributed as dist
        self = cls.__new__(cls)
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.block = None
        self.engine, self.stream, self.dtype, self.K, self.device = engine, stream, np.dtype(dtype), K, device
        self.engine.set_stream(stream.cuda_stream)
        self.peer_bytes = peer_bytes
        self._xch, self._xch_failed, self.last_exchange = None, False, None
        return self

    def close(self, keep_engine=False):
        if self._xch is not None:
            self._xch[3].close()
            self._xch = None
        if not keep_engine:
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
        return torch.as_tensor(_Arr(), device=f"cuda:{self.device}").view(self.K, ld)

    def _exchange_for(self, N, src):
        """The push exchange for (N, src), rebuilt (collectively) when this rank's B image has
        moved -- a different N re-allocates it, and the peer mappings of the old one would dangle.
        Every rank sees the same sequence of N and holds a B image of the same size, so they all
        take the same decision without asking each other."""
        if self.world == 1 or self._xch_failed:
            return None
        ptr, nbytes = self.engine.device_B(N)
        if self._xch is not None and self._xch[:3] == (N, src, ptr):
            return self._xch[3]
        if self._xch is not None:
            self._xch[3].close()
            self._xch = None
        if nbytes > self.peer_bytes:
            return None
        try:
            self._xch = (N, src, ptr, PushExchange([self.engine], N, group=self.group, root=src))
        except Exception:
            self._xch = None
        flags = [None] * self.world          # at build time only: did CUDA IPC work everywhere?
        self.dist.all_gather_object(flags, self._xch is not None, group=self.group)
        if not all(flags):
            if self._xch is not None:
                self._xch[3].close()
            self._xch, self._xch_failed = None, True
            return None
        return self._xch[3]

    def spmm(self, N, alpha, B_colmajor_root, beta, C_block_colmajor, src=0, rp_time=1, want_ns=True):
        """B_colmajor_root: the K x N column-major host B on rank ``src`` (ignored elsewhere).
        C_block_colmajor: this rank's block (in/out).  Returns the local kernel ns; with
        ``want_ns=False`` (nobody reads the kernel-only time) and a pushed B the call is the
        engine's host-facing call itself on every rank -- the holder's sx_spmm_* (B staging + an
        SpMM kernel that carries the push and C), the others' sx_spmm_staged_B_* (one kernel that
        waits for the push and carries C) -- and None is returned."""
        import torch
        x = self._exchange_for(N, src)
        if x is not None and not want_ns and rp_time <= 1:
            x.before_step(0)
            if self.rank == src:
                self.engine.spmm(N, alpha, B_colmajor_root, beta, C_block_colmajor, want_ns=False)
            else:
                self.engine.device_B(N)                    # marks B as present: the push fills it
                self.engine.spmm_staged_B(N, alpha, beta, C_block_colmajor)
            self.last_exchange = "push"
            return None
        if x is not None:
            if self.rank == src:
                self.engine.stage_B(N, B_colmajor_root)    # H2D + layout change on the root only
                x.before_step(0)                           # the launch below also pushes B (once the peers are done with the previous one)
            else:
                self.engine.device_B(N)                    # marks B as staged: the push fills it
                x.before_step(0)                           # the launch below waits for the push and acknowledges it
            self.last_exchange = "push"
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


class PanelPipeline:
    """A LARGE B at N > 1: B travels, and is consumed, in column panels.

    The reference handles N in passes of 8 columns, re-streaming A once per pass
    (``rp_time_N = rp_time * ((N + 7) >> 3)``, src/sextans.cpp:57,84,328,474), and hands the B
    window down its PEG chain while the PEGs compute (:909-941).  The multi-GPU form: the dense
    operands are held as P column panels of ``panel_cols`` columns (row-major images of their
    own: a column-major K x N host array IS its panels, back to back); panel p is broadcast from
    the rank that holds B on a communication stream (NCCL over NVLink / NVSwitch) while the SpMM
    of panel p-1 runs on the compute stream, so a step costs about  max(sum of broadcasts, sum
    of SpMMs) + one panel  instead of broadcast + SpMM.  Every output element is computed by the
    same chain of operations as in one pass, so results are bitwise those of the unpanelled call.

    Device-resident: ``load`` stages the operands once; ``step`` enqueues one whole SpMM
    (broadcasts included); ``result`` brings this rank's C block back (column-major).
    """

    def __init__(self, engine, rows, K, N, dtype, device, stream, panel_cols=None, group=None, src=0):
        import torch
        import torch.distributed as dist
        self.dist, self.group, self.src = dist, group, src
        self.engine, self.rows, self.K, self.N, self.dtype = engine, rows, K, N, np.dtype(dtype)
        self.device, self.stream = device, stream
        s = self.dtype.itemsize
        if panel_cols is None:
            panel_cols = 128 // s                      # 128-byte B rows: one lane group of 8 lanes
        self.pw = min(panel_cols, N)
        self.P = (N + self.pw - 1) // self.pw
        self.widths = [min(self.pw, N - p * self.pw) for p in range(self.P)]
        self.lds = [(w + 7) // 8 * 8 for w in self.widths]
        td = torch.float64 if self.dtype == np.float64 else torch.float32
        self.comm = torch.cuda.Stream(device=device)
        with torch.cuda.stream(stream):
            self.dB = [torch.zeros(K * ld, dtype=td, device=device) for ld in self.lds]
            self.dCin = [torch.zeros(rows * ld, dtype=td, device=device) for ld in self.lds]
            self.dCout = [torch.zeros(rows * ld, dtype=td, device=device) for ld in self.lds]
        self.landed = [torch.cuda.Event() for _ in range(self.P)]
        self.freed = [torch.cuda.Event() for _ in range(self.P)]
        self._first = True
        self.td = td

    def load(self, B_colmajor, Cin_block_colmajor):
        """Stage the operands: B (on the source rank; None elsewhere) and this rank's C_in block."""
        import torch
        e, K, rows = self.engine, self.K, self.rows
        with torch.cuda.stream(self.stream):
            c0 = 0
            for p, (w, ld) in enumerate(zip(self.widths, self.lds)):
                if B_colmajor is not None:
                    src = torch.from_numpy(np.ascontiguousarray(B_colmajor[c0 * K:(c0 + w) * K])).to(self.device)
                    e.colmajor_to_rowmajor(K, w, src, self.dB[p], ld)
                src = torch.from_numpy(np.ascontiguousarray(Cin_block_colmajor[c0 * rows:(c0 + w) * rows])).to(self.device)
                e.colmajor_to_rowmajor(rows, w, src, self.dCin[p], ld)
                c0 += w
        self.stream.synchronize()

    def step(self, alpha, beta):
        import torch
        for p in range(self.P):
            with torch.cuda.stream(self.comm):
                if self._first:
                    self.comm.wait_stream(self.stream)              # the staging of B has finished
                else:
                    self.comm.wait_event(self.freed[p])             # the previous SpMM on this panel has finished reading it
                self.dist.broadcast(self.dB[p], src=self.src, group=self.group)
                self.landed[p].record(self.comm)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.landed[p])
                self.engine.spmm_device(self.widths[p], alpha, self.dB[p], self.lds[p], beta, self.dCin[p], self.dCout[p], self.lds[p])
                self.freed[p].record(self.stream)
            self._first = False

    def result(self):
        import torch
        out = np.empty(self.rows * self.N, dtype=self.dtype)
        c0 = 0
        with torch.cuda.stream(self.stream):
            parts = []
            for p, (w, ld) in enumerate(zip(self.widths, self.lds)):
                t = torch.empty(self.rows * w, dtype=self.td, device=self.device)
                self.engine.rowmajor_to_colmajor(self.rows, w, self.dCout[p], ld, t)
                parts.append((c0, w, t))
                c0 += w
        self.stream.synchronize()
        for c0, w, t in parts:
            out[c0 * self.rows:(c0 + w) * self.rows] = t.cpu().numpy()
        return out
