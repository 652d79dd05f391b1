"""Three batched launches (sx_spmm_device_batch_*, 20 operand triples each) of nasa4704 N=16 fp64 and nothing else on
the edge-list kernel: the target of `ncu --set full -k regex:spmm_edgelist_kernel` for the batched form."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sextans_b200 as sx  # noqa: E402
from sextans_b200 import workloads as wl  # noqa: E402

M, K, nnz, rp, ci, v = sx.load_mtx(wl.suitesparse_path("nasa4704"), np.float64)
N, nb, ld = 16, 20, 16
dev = torch.device("cuda", 0)
e = sx.Engine(0)
s = torch.cuda.Stream()
e.set_stream(s.cuda_stream)
e.upload_csr(M, K, rp, ci, v)
T = 180
dB = torch.ones(T * K * ld, dtype=torch.float64, device=dev)
dCin = torch.ones(T * M * ld, dtype=torch.float64, device=dev)
dCout = torch.zeros(T * M * ld, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for i in range(3):
    o = i * nb
    e.spmm_device_batch(N, nb, 0.85, dB[o * K * ld:], ld, K * ld, -2.06, dCin[o * M * ld:], dCout[o * M * ld:], ld, M * ld)
s.synchronize()
print("batched launches done; last kernel", e.info(sx.INFO_LAST_KERNEL))
