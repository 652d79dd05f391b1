#!/usr/bin/env python
"""GPU probe (no torch): power-law C5-like matrix, fp64 N=16, kernel time of one SpMM with
and without column windows, through the staged C-ABI calls (sx_stage_* / sx_launch_*).
Prints one JSON line per setting, flushed as it goes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sextans_b200 as sx  # noqa: E402
from sextans_b200 import workloads as wl  # noqa: E402

M = K = int(float(os.environ.get("PROBE_M", 1e6)))
NNZ = int(float(os.environ.get("PROBE_NNZ", 1e8)))
N = int(os.environ.get("PROBE_N", 16))
WINDOWS = [int(x) for x in os.environ.get("PROBE_W", "0,524288,262144,131072").split(",")]
REPS = int(os.environ.get("PROBE_REPS", 5))

t0 = time.time()
rp, ci, v = wl.powerlaw_csr(M, K, NNZ)
B, Cin = wl.random_dense(M, K, N, 12345, np.float64)
print(json.dumps({"gen_s": round(time.time() - t0, 1), "nnz": int(ci.size)}), flush=True)
ref = None
with sx.Engine(0) as eng:
    for W in WINDOWS:
        t0 = time.time()
        eng.set_option(sx.OPT_COL_WINDOW_ROWS, W)
        eng.upload_csr(M, K, rp, ci, v)
        up = time.time() - t0
        eng.stage_B(N, B)
        eng.stage_C(N, Cin)
        eng.launch(0.85, -2.06, 2)                      # warm-up (plans, attributes)
        ns = eng.launch(0.85, -2.06, REPS)
        out = np.empty(M * N, dtype=np.float64)
        eng.fetch_C(out)
        if ref is None:
            ref = out
        print(json.dumps({"W": W, "windows": eng.info(sx.INFO_COL_WINDOWS), "ms_per_spmm": ns / REPS / 1e6,
                          "gflops": 2.0 * ci.size * N / (ns / REPS), "upload_s": round(up, 2),
                          "last_kernel": eng.info(sx.INFO_LAST_KERNEL),
                          "bitwise_equal_to_unwindowed": bool(np.array_equal(out.view(np.uint64), ref.view(np.uint64))),
                          "max_abs_diff": float(np.max(np.abs(out - ref)))}), flush=True)
