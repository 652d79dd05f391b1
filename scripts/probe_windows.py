#!/usr/bin/env python
"""GPU probe (no torch): kernel time of one SpMM on a C5-like power-law matrix (fp64, N=16)
or a C4-like uniform one (fp32, N=128) for a list of (column-window rows, prefetch)
settings, through the staged C-ABI calls (sx_stage_* / sx_launch_*).  One JSON line per
setting, flushed as it goes.

  PROBE_KIND=powerlaw|uniform  PROBE_M  PROBE_NNZ  PROBE_N  PROBE_REPS
  PROBE_SET="W:prefetch,..."   e.g. "0:0,0:1,262144:1"
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sextans_b200 as sx  # noqa: E402
from sextans_b200 import workloads as wl  # noqa: E402

KIND = os.environ.get("PROBE_KIND", "powerlaw")
M = K = int(float(os.environ.get("PROBE_M", 1e6)))
NNZ = int(float(os.environ.get("PROBE_NNZ", 1e8 if KIND == "powerlaw" else 2e7)))
N = int(os.environ.get("PROBE_N", 16 if KIND == "powerlaw" else 128))
SETTINGS = [tuple(int(x) for x in s.split(":")) for s in os.environ.get("PROBE_SET", "0:0,0:1,0:-1").split(",")]
REPS = int(os.environ.get("PROBE_REPS", 10))
dtype = np.float64 if KIND == "powerlaw" else np.float32

t0 = time.time()
if KIND == "powerlaw":
    rp, ci, v = wl.powerlaw_csr(M, K, NNZ)
else:
    rp, ci, v = wl.uniform_csr(M, K, NNZ // M)
B, Cin = wl.random_dense(M, K, N, 12345, dtype)
print(json.dumps({"kind": KIND, "M": M, "N": N, "dtype": np.dtype(dtype).name, "gen_s": round(time.time() - t0, 1),
                  "nnz": int(ci.size)}), flush=True)
ref = None
a, b = dtype(np.float32(0.85)), dtype(np.float32(-2.06))
with sx.Engine(0) as eng:
    for W, pf in SETTINGS:
        t0 = time.time()
        eng.set_option(sx.OPT_COL_WINDOW_ROWS, W)
        eng.set_option(sx.OPT_PREFETCH, pf)
        eng.upload_csr(M, K, rp, ci, v)
        up = time.time() - t0
        eng.stage_B(N, B)
        eng.stage_C(N, Cin)
        eng.launch(a, b, 2)                      # warm-up (plans, attributes)
        ns = eng.launch(a, b, REPS)
        out = np.empty(M * N, dtype=dtype)
        eng.fetch_C(out)
        if ref is None:
            ref = out
        print(json.dumps({"W": W, "prefetch": pf, "windows": eng.info(sx.INFO_COL_WINDOWS),
                          "ms_per_spmm": ns / REPS / 1e6, "gflops": 2.0 * ci.size * N / (ns / REPS),
                          "upload_s": round(up, 2), "last_kernel": eng.info(sx.INFO_LAST_KERNEL),
                          "bitwise_equal_to_first": bool(np.array_equal(out.view(np.uint8), ref.view(np.uint8))),
                          "max_abs_diff": float(np.max(np.abs(out - ref)))}), flush=True)
