"""Experiment: C4 (uniform, N=128 fp32) as N/np passes over an np-column panel of B that
fits in L2, against the single pass.  Uses only the public device-resident call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sextans_b200 as sx
from sextans_b200 import workloads as wl

M = K = 1_000_000
N = 128
rp, ci, v = wl.uniform_csr(M, K, 20, 12345, np.float32)
dev = torch.device("cuda:0")
eng = sx.Engine(0)
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
eng.set_option(sx.OPT_KERNEL, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
eng.upload_csr(M, K, rp, ci, v)
with torch.cuda.stream(st):
    B = torch.rand(K, N, device=dev) * 2 - 1
    Cin = torch.rand(M, N, device=dev) * 2 - 1
    Cout = torch.empty_like(Cin)
    Cref = torch.empty_like(Cin)
    flush = torch.empty(512 << 18, dtype=torch.int32, device=dev)

def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(st):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); fn(); b.record(st)
        st.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.mean(ts))

with torch.cuda.stream(st):
    eng.spmm_device(N, 0.85, B, N, -2.06, Cin, Cref, N)
print("single pass       ", timed(lambda: eng.spmm_device(N, 0.85, B, N, -2.06, Cin, Cref, N)))
for npan in (64, 32, 16, 8):
    with torch.cuda.stream(st):
        Bp = B.view(K, N // npan, npan).permute(1, 0, 2).contiguous()      # [N/np][K][np]
    def run():
        for p in range(N // npan):
            off = p * npan * 4
            eng.spmm_device(npan, 0.85, Bp[p].data_ptr(), npan, -2.06, Cin.data_ptr() + off, Cout.data_ptr() + off, N)
    with torch.cuda.stream(st):
        run()
    st.synchronize()
    ok = torch.equal(Cout, Cref)
    print(f"panels of {npan:3d} cols", timed(run), "bit-equal" if ok else "MISMATCH")
