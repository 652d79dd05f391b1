#!/usr/bin/env python
"""Phase timeline of spmm_edgelist_kernel from the trace variant of the library
(scripts/build_variant.sh trace "-DSX_EDGE_TRACE"; run with SX_LIBRARY_PATH pointing at it):
K steps replayed as one CUDA graph, every block's %globaltimer stamps read back and summarised
per launch -- where a ~3.5 us step goes."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nasa4704")
    ap.add_argument("--ncols", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--pdl", type=int, default=-1)
    ap.add_argument("--prefetch", type=int, default=-1)
    ap.add_argument("--copies", type=int, default=24)
    a = ap.parse_args()
    import torch
    import sextans_b200 as sx
    import bench
    w = bench.build_workload(a.workload, a.ncols)
    L = sx.lib()
    if not hasattr(L, "sx_debug_edge_trace"):
        raise SystemExit("not a trace build: scripts/build_variant.sh trace -DSX_EDGE_TRACE, then SX_LIBRARY_PATH=...")
    L.sx_debug_edge_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    M, K, N, dtype = w["M"], w["K"], w["N"], w["dtype"]
    td = torch.float64 if dtype == np.float64 else torch.float32
    ld = (N + 7) // 8 * 8
    engs, ops = [], []
    with torch.cuda.stream(stream):
        dB_cm, dC_cm = torch.from_numpy(w["B"]).to(dev), torch.from_numpy(w["Cin"]).to(dev)
    for _ in range(a.copies):
        e = sx.Engine(0)
        e.set_stream(stream.cuda_stream)
        e.set_option(sx.OPT_PDL, a.pdl)
        e.set_option(sx.OPT_PREFETCH, a.prefetch)
        e.upload_csr(M, K, w["rowptr"], w["colidx"], w["val"])
        with torch.cuda.stream(stream):
            dB, dCi, dCo = (torch.zeros(r * ld, dtype=td, device=dev) for r in (K, M, M))
            e.colmajor_to_rowmajor(K, N, dB_cm, dB, ld)
            e.colmajor_to_rowmajor(M, N, dC_cm, dCi, ld)
        engs.append(e); ops.append((dB, dCi, dCo))

    def step(i):
        j = i % a.copies
        engs[j].spmm_device(N, bench.ALPHA, ops[j][0], ld, bench.BETA, ops[j][1], ops[j][2], ld)
    with torch.cuda.stream(stream):
        for i in range(a.copies):
            step(i)
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(a.steps):
            step(i)
    buf = np.zeros((1 << 16, 8), dtype=np.uint64)
    n = C.c_int()
    with torch.cuda.stream(stream):
        g.replay()
    stream.synchronize()
    L.sx_debug_edge_trace(engs[0]._ctx, buf.ctypes.data, buf.shape[0], C.byref(n))     # discard: warm-up + first replay
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        g.replay()
        e1.record(stream)
    stream.synchronize()
    L.sx_debug_edge_trace(engs[0]._ctx, buf.ctypes.data, buf.shape[0], C.byref(n))
    rows = buf[:n.value].astype(np.int64)
    nblk = n.value // a.steps
    print(f"{a.workload} N={N} {dtype.name}: {a.steps} steps in one graph, {nblk} blocks per launch, events say {e0.elapsed_time(e1) * 1e3 / a.steps:.2f} us per step")
    # a launch = nblk consecutive rows in order of entry time is not guaranteed with PDL overlap: group by order of the
    # block-done stamp instead (launch k completes before k+1 passes its wait)
    order = np.argsort(rows[:, 2], kind="stable")          # by the moment the dependent-launch wait was passed
    rows = rows[order]
    t_ref = rows[:, 0].min()
    prev_done = None
    print(" launch  first entry  last entry | prologue(med)  held at wait(med) | wait passed: first..last | staged(med after wait)  block done(med)  last done | step period  gap done->next wait")
    firsts = []
    for k in range(a.steps):
        r = rows[k * nblk:(k + 1) * nblk]
        ent, pre, post, staged, t0done, done = (r[:, i] - t_ref for i in range(6))
        firsts.append(post.min())
        gap = "" if prev_done is None else f"{(post.min() - prev_done) / 1e3:6.2f}"
        period = "" if k == 0 else f"{(firsts[k] - firsts[k - 1]) / 1e3:6.2f}"
        print(f" {k:5d}  {ent.min() / 1e3:10.2f}  {ent.max() / 1e3:10.2f} | {np.median(pre - ent) / 1e3:8.2f}       {np.median(post - pre) / 1e3:8.2f}        |"
              f" {post.min() / 1e3:8.2f}..{post.max() / 1e3:8.2f}    | {np.median(staged - post) / 1e3:8.2f}               {np.median(done - post) / 1e3:8.2f}        {done.max() / 1e3:8.2f}  |"
              f" {period:>7}      {gap:>7}")
        prev_done = done.max()
    per = np.diff(firsts)[2:]
    print(f"median step period {np.median(per) / 1e3:.2f} us (globaltimer resolution: {np.min(np.diff(np.unique(rows[:, :6].ravel())))} ns)")
    for e in engs:
        e.close()


if __name__ == "__main__":
    main()
