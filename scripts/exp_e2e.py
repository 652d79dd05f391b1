"""Where does the host-facing call spend its time?  nasa4704 N=16 f64 through sx_spmm_f64."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sextans_b200 as sx
from sextans_b200 import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "nasa4704"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dtype = np.float64 if (len(sys.argv) <= 3 or sys.argv[3] == "f64") else np.float32
M, K, nnz, rp, ci, v = sx.load_mtx(wl.suitesparse_path(name), dtype)
B, Cin = wl.host_dense(M, K, N, dtype)
eng = sx.Engine(0)
eng.upload_csr(M, K, rp, ci, v)
hB, hC = sx.pinned_empty(K * N, dtype), sx.pinned_empty(M * N, dtype)
hB[:] = B

def loop(label, Bbuf, Cbuf, reps=200):
    for _ in range(5):
        Cbuf[:] = Cin
        eng.spmm(N, 0.85, Bbuf, -2.06, Cbuf)
    ts, ks = [], []
    for _ in range(reps):
        Cbuf[:] = Cin
        t0 = time.perf_counter()
        ns = eng.spmm(N, 0.85, Bbuf, -2.06, Cbuf)
        ts.append(time.perf_counter() - t0)
        ks.append(ns)
    print(f"{label:34s} wall median {np.median(ts)*1e6:8.1f} us  min {np.min(ts)*1e6:8.1f} us   kernel-only {np.median(ks)/1e3:6.1f} us   path {eng.info(sx.INFO_HOST_PATH)}")

loop("pinned, zero-copy kernels", hB, hC)
eng.set_option(sx.OPT_ZEROCOPY_BYTES, 0)
loop("pinned, cudaMemcpyAsync", hB, hC)
loop("pageable, cudaMemcpyAsync", B.copy(), Cin.copy())
eng.set_option(sx.OPT_ZEROCOPY_BYTES, 3 << 19)
