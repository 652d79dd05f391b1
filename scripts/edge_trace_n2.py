#!/usr/bin/env python
"""Phase timeline of the edge-list kernel on every rank of an N-GPU run with the push exchange
(trace variant of the library; run under torchrun with SX_LIBRARY_PATH set)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import sextans_b200 as sx
    import bench
    from sextans_b200.rowblock import PushExchange
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = sx.lib()
    L.sx_debug_edge_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    w = bench.build_workload("nasa4704")
    M, K, N, dtype = w["M"], w["K"], w["N"], w["dtype"]
    stream = torch.cuda.Stream(device=dev)
    ld, steps, copies = 16, 20, 20
    engs, ops = [], []
    with torch.cuda.stream(stream):
        dB_cm, dC_cm = torch.from_numpy(w["B"]).to(dev), torch.from_numpy(w["Cin"]).to(dev)
    for _ in range(copies):
        e = sx.Engine(local)
        e.set_stream(stream.cuda_stream)
        e.upload_csr(M, K, w["rowptr"], w["colidx"], w["val"])
        with torch.cuda.stream(stream):
            dB = e.device_B(N)[0]
            dCi, dCo = (torch.zeros(M * ld, dtype=torch.float64, device=dev) for _ in range(2))
            if rank == 0:
                e.colmajor_to_rowmajor(K, N, dB_cm, dB, ld)
            e.colmajor_to_rowmajor(M, N, dC_cm, dCi, ld)
        engs.append(e); ops.append((dB, dCi, dCo))
    stream.synchronize()
    px = PushExchange(engs, N)

    def step(i):
        j = i % copies
        px.before_step(i)
        engs[j].spmm_device(N, bench.ALPHA, ops[j][0], ld, bench.BETA, ops[j][1], ops[j][2], ld)
    with torch.cuda.stream(stream):
        for i in range(copies):
            step(i)
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(steps):
            step(i)
    buf = np.zeros((1 << 16, 8), dtype=np.uint64)
    n = C.c_int()
    for _ in range(2):
        with torch.cuda.stream(stream):
            g.replay()
    torch.cuda.synchronize(); dist.barrier()
    L.sx_debug_edge_trace(engs[0]._ctx, buf.ctypes.data, buf.shape[0], C.byref(n))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        g.replay()
        e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    L.sx_debug_edge_trace(engs[0]._ctx, buf.ctypes.data, buf.shape[0], C.byref(n))
    rows = buf[:n.value].astype(np.int64)
    nblk = n.value // steps
    rows = rows[np.argsort(rows[:, 2], kind="stable")]
    t_ref = rows[:, 0].min()
    out = [f"rank {rank}: events say {e0.elapsed_time(e1) * 1e3 / steps:.2f} us per step; {nblk} blocks per launch",
           " launch | wait passed (first) | flag seen (med after wait) | staged (med after flag) | block done (med after staged) | last done | period"]
    firsts = []
    for k in range(steps):
        r = rows[k * nblk:(k + 1) * nblk] - t_ref
        firsts.append(r[:, 2].min())
        per = "" if k == 0 else f"{(firsts[k] - firsts[k - 1]) / 1e3:6.2f}"
        out.append(f" {k:5d}  | {r[:, 2].min() / 1e3:8.2f}          | {np.median(r[:, 4] - r[:, 2]) / 1e3:8.2f}                | {np.median(r[:, 3] - r[:, 4]) / 1e3:8.2f}"
                   f"              | {np.median(r[:, 5] - r[:, 3]) / 1e3:8.2f}                  | {r[:, 5].max() / 1e3:8.2f} | {per}")
    for r_ in range(world):
        if r_ == rank:
            print("\n".join(out), flush=True)
        dist.barrier()
    px.close()
    for e in engs:
        e.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
