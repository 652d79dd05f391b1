#!/usr/bin/env python
"""bench.py -- SpMM GFLOP/s (2*nnz*N) and achieved HBM GB/s against the roofline.

A "step" is one SpMM  C = alpha*A*B + beta*C  over the named workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs all|none|a,b,..]

Headline (value / ms_per_step / roofline / e2e): BASELINE.json configs[1], nasa4704.mtx N=16 fp64.
`configs` (same JSON line, compact, emitted before the prose): every other BASELINE config --
pcrystk02 N=8/16/32/64 fp32, the uniform synthetic C4, the power-law synthetic C5 -- each with
{ms, gflops, frac, traffic_ratio, kernel, parity}; parity = the GPU result of one step against
the CPU oracle over EVERY row (all host threads; rows are independent, so the threaded oracle
is bitwise the single-threaded one).  At --gpus N>1 `configs` carries the STRONG scaling of
the one C4 and the one C5 matrix over nnz-balanced row blocks (sx_partition_rows, one per GPU):
kernel-only and including the broadcast of B, per-rank nnz imbalance, parity on every rank.

Own arm, per rank (one process per GPU):
  value : steps between CUDA events on the launching stream, replayed as a CUDA graph.  The
          operands exist in R independent device copies (R x bytes > 2.2 x the 126 MB L2) and
          step i uses copy i mod R, across replays too, so every step finds its operands in
          HBM ("inputs larger than L2"; no flush kernel in the timed region).  One graph holds
          every distinct K-step window of the rotation back to back (S = W*K steps, a whole
          number of passes over the copies): the steady state of a long chain of dependent
          SpMMs.  The replay is repeated until the timed region is >= 50 ms; ms_per_step is the
          MEDIAN repetition / S (max over ranks per repetition).  `run.k_step_graphs` holds
          the same steps replayed as separate K-step graphs (every K steps then pay a graph
          launch), what earlier records of this bench reported.
  batched: the headline workload through sx_spmm_device_batch_* (20 operand triples per launch).
  e2e   : the same step through the host-facing C-ABI call sx_spmm_* with pinned host B and C
          (column-major, as the host program holds them), H2D + kernels + D2H inside the timed
          region; at N>1 through ShardedSpMM (B on rank 0's host, exchange, C blocks back).
  N>1 headline: weak scaling -- N stacked copies of the matrix, one row block per GPU, the
          exchange of B from rank 0 inside every step.
Reference arm (--impl reference): the reference's CPU path cpu_spmm_CSR on the host cores
(oracle/_ref = the reference's own header for fp32, the line-for-line port for fp64).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    #            kind         dtype        N
    "nasa4704": ("suitesparse", np.float64, 16),
    "pcrystk02": ("suitesparse", np.float32, 16),
    "uniform": ("uniform", np.float32, 128),
    "powerlaw": ("powerlaw", np.float64, 16),
    "powerlaw_blocked": ("powerlaw_blocked", np.float64, 16),
    "fem": ("fem", np.float64, 16),
}
# the `configs` entries of the default run: (key, workload, N)
CONFIGS_1GPU = [("pcrystk02_n8", "pcrystk02", 8), ("pcrystk02_n16", "pcrystk02", 16),
                ("pcrystk02_n32", "pcrystk02", 32), ("pcrystk02_n64", "pcrystk02", 64),
                ("uniform_c4", "uniform", 128), ("powerlaw_c5", "powerlaw", 16)]
CONFIGS_EXTRA = [("powerlaw_blocked", "powerlaw_blocked", 16)]      # on request: --configs powerlaw_blocked
CONFIGS_NGPU = [("uniform_c4", "uniform", 128), ("powerlaw_c5", "powerlaw", 16)]
ALPHA, BETA = float(np.float32(0.85)), float(np.float32(-2.06))   # host.cpp:29-30
L2_BYTES = 126 * 1024 * 1024
KERNEL_NAMES = {1: "spmm_rows_kernel", 2: "spmm_staged_kernel", 3: "spmm_window_kernel", 4: "spmm_panels_dmma_kernel",
                5: "spmm_staged_kernel<WIN>", 7: "spmm_slide_kernel", 8: "spmm_edgelist_kernel", 9: "spmm_edgelist_kernel<HOSTC>",
                10: "spmm_edgelist_host_kernel"}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--workload", default="nasa4704", choices=sorted(WORKLOADS))
    p.add_argument("--configs", default="all", help="all | none | comma-separated keys of the `configs` entries to run")
    p.add_argument("--ncols", type=int, default=0, help="override the workload's N")
    p.add_argument("--dtype", default="", choices=["", "f32", "f64"])
    p.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic workloads (testing only; the line says so)")
    p.add_argument("--arith", default="strict", choices=["strict", "fast"])
    p.add_argument("--kernel", type=int, default=0, help="SX_OPT_KERNEL (0 auto, 1 row per lane group, 2 TMA-staged items, 3 TMA-staged B window, 4 sliding B window (needs --slide), 5 edge lists)")
    p.add_argument("--item-nnz", type=int, default=0, help="SX_OPT_ITEM_NNZ (0 auto)")
    p.add_argument("--split", type=int, default=-1, help="SX_OPT_SPLIT_ROW_NNZ (-1 default)")
    p.add_argument("--band", type=int, default=2000, help="fem workload: couplings reach +-band nodes")
    p.add_argument("--tiles", type=int, default=0, help="SX_OPT_TILE_MIN_ROWS (fp64 dense-tile tensor-core variant; 0 off)")
    p.add_argument("--col-window-rows", type=int, default=0, help="SX_OPT_COL_WINDOW_ROWS (0 off, -1 = 32 MiB of B per window)")
    p.add_argument("--slide", type=int, default=0, help="SX_OPT_SLIDE: chains per SM of the sliding-window kernel; use with --kernel 4")
    p.add_argument("--pdl", type=int, default=-1, choices=[-1, 0, 1], help="SX_OPT_PDL: programmatic dependent launch (-1 auto: on for the edge-list kernel)")
    p.add_argument("--prefetch", type=int, default=-1, choices=[-1, 0, 1], help="SX_OPT_PREFETCH: L2 prefetch hints (-1 auto)")
    p.add_argument("--batch", type=int, default=20, help="also time sx_spmm_device_batch_* with this many (B, C) pairs per launch on the headline workload (0/1: skip)")
    p.add_argument("--host-groups", type=int, default=0, help="SX_OPT_HOST_GROUPS: column groups of the fused host-facing call (0 auto, 1 no pipeline)")
    p.add_argument("--panel-cols", type=int, default=0, help="SX_OPT_PANEL_COLS: columns of B and C per pass (0 auto)")
    p.add_argument("--host-fused", type=int, default=-1, choices=[-1, 0, 1], help="SX_OPT_HOST_FUSED: e2e calls let the SpMM kernel carry C across PCIe (-1 auto: on)")
    p.add_argument("--ref-threads", type=int, default=-1, help="--impl reference: threads of the CPU path (-1 = all cores, row-parallel; 1 = as the reference runs it)")
    p.add_argument("--peer-bytes", type=int, default=8 << 20, help="N>1: B images up to this size travel through peer memory instead of NCCL")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-pipelined-e2e", action="store_true", help="skip the two-calls-in-flight form of the e2e measurement")
    p.add_argument("--no-flush", action="store_true", help="one device copy: leave L2 warm between steps")
    p.add_argument("--no-graph", action="store_true", help="launch the timed steps one by one instead of replaying CUDA graphs")
    p.add_argument("--min-region-ms", type=float, default=50.0, help="repeat the K-step replay until the timed region is this long")
    return p.parse_args()


# ---- workloads ------------------------------------------------------------------------------
def build_workload(name, ncols=0, dtype_override="", scale=1.0, band=2000, loader="product"):
    """-> dict(name, desc, M, K, nnz, N, dtype, rowptr, colidx, val, B, Cin); B, Cin column-major 1-D.
    loader: "product" (sx_load_mtx_*) for the own arm, "oracle" for the reference arm, which must
    not touch the product library."""
    from sextans_b200 import workloads as wl
    kind, dtype, N = WORKLOADS[name]
    if dtype_override:
        dtype = np.float32 if dtype_override == "f32" else np.float64
    if ncols:
        N = ncols
    extra = {}
    tname = np.dtype(dtype).name
    if kind == "suitesparse":
        if loader == "oracle":
            import oracle
            M, K, nnz, rp, ci, v = oracle.load_mtx(wl.suitesparse_path(name), dtype)[:6]
        else:
            import sextans_b200 as sx
            M, K, nnz, rp, ci, v = sx.load_mtx(wl.suitesparse_path(name), dtype)
        B, Cin = wl.host_dense(M, K, N, dtype)
        desc = f"{name}.mtx M=K={M} nnz={nnz} N={N} {tname}, B=1, C_in=(m+1)(n+1)/M/N (host.cpp:100-111)"
    elif kind == "uniform":
        M = K = max(1000, int(1_000_000 * scale))
        rp, ci, v = wl.uniform_csr(M, K, 20, 12345, dtype)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        desc = f"synthetic uniform CSR M=K={M} nnz={nnz} (20/row) N={N} {tname}, seed 12345"
    elif kind == "fem":
        nodes = max(250, int(250_000 * scale))
        M = K = nodes * 4
        rp, ci, v = wl.fem_like_csr(nodes, 4, 23, 12345, dtype, band=band)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        desc = f"synthetic FEM-like CSR (4 dof/node, dense 4x4 couplings within +-{band} nodes) M=K={M} nnz={nnz} N={N} {tname}, seed 12345"
    elif kind == "powerlaw_blocked":
        M = K = max(1000, int(1_000_000 * scale))
        rp, ci, v, planted = wl.powerlaw_blocked_csr(M, K, int(100_000_000 * scale), 12345, dtype)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        extra["planted_nnz"] = planted
        desc = f"synthetic power-law CSR with planted dense 16x16 blocks ({planted / nnz:.0%} of nnz) M=K={M} nnz={nnz} N={N} {tname}, seed 12345"
    else:
        M = K = max(1000, int(1_000_000 * scale))
        rp, ci, v = wl.powerlaw_csr(M, K, int(100_000_000 * scale), 12345, dtype)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        desc = f"synthetic power-law CSR M=K={M} nnz={nnz} N={N} {tname}, seed 12345"
    if scale != 1.0 and kind != "suitesparse":
        desc += f" [SCALED x{scale}: not the BASELINE size]"
    return dict(name=name, desc=desc, M=M, K=K, nnz=nnz, N=N, dtype=np.dtype(dtype),
                rowptr=rp, colidx=ci, val=v, B=B, Cin=Cin, **extra)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name, N, dtype):
    """dram bytes per launch of the SpMM kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{name}_n{N}_{dtype}")
    except Exception:
        return None


def max_rel_err(x, y):
    """|x-y| / max(|y|, 1e-30), max over elements (SURVEY.md 8(c) definition)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-30))) if x.size else 0.0


def scaled_err(x, y):
    """max |x-y| / max|y| -- insensitive to cancellation in individual entries."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.max(np.abs(x - y)) / max(float(np.max(np.abs(y))), 1e-30)) if x.size else 0.0


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU arms (checker code, timed as a reported baseline) -----------------------------------
def bounded_sample(w, seconds, threads):
    """A row prefix of the workload sized so that one CPU pass takes about `seconds` (full B)."""
    M, N, nnz = w["M"], w["N"], w["nnz"]
    est = 2.0 * nnz * N / (1.5e9 * max(1, threads))
    if est <= seconds:
        return w["M"], w["nnz"], w["rowptr"], w["colidx"], w["val"], w["Cin"], "the whole workload"
    Ms = max(1, int(M * seconds / est))
    rp = np.ascontiguousarray(w["rowptr"][:Ms + 1])
    nz = int(rp[-1])
    Cin = np.ascontiguousarray(w["Cin"].reshape(N, M)[:, :Ms]).ravel()
    return Ms, nz, rp, w["colidx"][:nz], w["val"][:nz], Cin, f"first {Ms} rows ({nz} nnz) of the workload, full B"


def cpu_pass(w, M, rp, ci, v, C, threads):
    import oracle
    if w["dtype"] == np.float32 and threads == 1 and oracle.ref() is not None:
        oracle.ref_spmm_csr(M, w["N"], w["K"], rp, ci, v, ALPHA, w["B"], BETA, C)
        return "reference"
    oracle.spmm_csr(M, w["N"], w["K"], rp, ci, v, w["dtype"].type(ALPHA), w["B"], w["dtype"].type(BETA), C, threads=threads)
    return "port"


def cpu_baseline(w, threads, budget_s=10.0):
    M, nnz, rp, ci, v, Cin, sample = bounded_sample(w, budget_s / 3, threads)
    times, kind = [], "port"
    t_all = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_all < budget_s / 2 and len(times) < 20):
        C = Cin.copy()
        t0 = time.perf_counter()
        kind = cpu_pass(w, M, rp, ci, v, C, threads)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s * 1.5:
            break
    best = min(times)
    return {"value": 2.0 * nnz * w["N"] / best / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": kind,
            "sample": f"{sample}; best of {len(times)} runs, {best * 1e3:.3f} ms"}


def run_reference(args):
    """The reference arm: cpu_spmm_CSR on the host cores.  The function itself is single-threaded
    (src/sparse_helper.h:262-290); with all the host threads it can use, its rows -- which are
    independent -- are dealt to OpenMP threads by the oracle port, bitwise the same result.  The
    line also carries the 1-thread figure: the reference's own compiled header when the run is
    fp32, the port otherwise."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    w = build_workload(args.workload, args.ncols, args.dtype, args.scale, args.band, loader="oracle")
    all_threads = max(1, oracle.lib().sx_oracle_max_threads())
    threads = all_threads if args.ref_threads < 0 else max(1, args.ref_threads)
    M, nnz, rp, ci, v, Cin, sample = bounded_sample(w, 2.0, threads)
    warm = max(3, args.warmup)
    for _ in range(warm):
        cpu_pass(w, M, rp, ci, v, Cin.copy(), threads)
    total, kind = 0.0, "port"
    for _ in range(args.steps):
        C = Cin.copy()
        t0 = time.perf_counter()
        kind = cpu_pass(w, M, rp, ci, v, C, threads)
        total += time.perf_counter() - t0
    val = 2.0 * nnz * w["N"] * args.steps / total / 1e9
    one = cpu_baseline(w, 1, budget_s=6.0) if threads != 1 else None
    emit({
        "impl": "reference", "metric": "SpMM GFLOP/s (2*nnz*N)", "value": val, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if w["dtype"] == np.float64 else "f32",
        "data": data_label(w), "config": config_of(w),
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": kind, "sample": sample + " per step",
                         "note": "cpu_spmm_CSR (src/sparse_helper.h:262-290), rows dealt to OpenMP threads by the oracle port" if threads > 1
                                 else "cpu_spmm_CSR as the reference runs it: one thread",
                         "one_thread": one},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def data_label(w):
    return "synthetic" if w["name"] not in ("nasa4704", "pcrystk02") else "SuiteSparse fixture shipped with the reference, host program's B/C (synthetic dense operands)"


def config_of(w):
    """The workload's identity: the same dict on both arms."""
    return {"workload": w["desc"], "alpha": ALPHA, "beta": BETA}


# ---- own arm ---------------------------------------------------------------------------------
class Case:
    """One workload (or one rank's row block of it) resident on the device in R rotating copies.
    own_B: the engines' own B images are the operands (what a push exchange writes into)."""

    def __init__(self, w, args, dev, stream, copies=None, fill_B=True, own_B=False):
        import torch
        import sextans_b200 as sx
        from sextans_b200 import workloads as wl
        self.w, self.args, self.dev, self.stream = w, args, dev, stream
        M, K, N, nnz, dtype = w["M"], w["K"], w["N"], w["nnz"], w["dtype"]
        self.s = dtype.itemsize
        self.tdtype = torch.float64 if dtype == np.float64 else torch.float32
        self.alg_bytes = wl.algorithmic_bytes(M, K, nnz, N, self.s)
        if copies is None:
            copies = 1 if (args.no_flush or self.alg_bytes >= 2 * L2_BYTES) else int(math.ceil(2.2 * L2_BYTES / self.alg_bytes))
        self.R = copies
        self.ld = (N + 7) // 8 * 8
        self.engines, self.dB, self.dCin, self.dCout = [], [], [], []
        with torch.cuda.stream(stream):
            dB_cm = torch.from_numpy(w["B"]).to(dev)
            dC_cm = torch.from_numpy(w["Cin"]).to(dev)
        for _ in range(self.R):
            e = sx.Engine(dev.index, arith=sx.STRICT if args.arith == "strict" else sx.FAST)
            e.set_stream(stream.cuda_stream)
            e.set_option(sx.OPT_KERNEL, args.kernel)
            e.set_option(sx.OPT_ITEM_NNZ, args.item_nnz)
            e.set_option(sx.OPT_TILE_MIN_ROWS, args.tiles if dtype == np.float64 else 0)
            cw = args.col_window_rows
            if cw < 0:
                cw = max(1, (32 << 20) // (self.ld * self.s))
            e.set_option(sx.OPT_COL_WINDOW_ROWS, cw)
            e.set_option(sx.OPT_PDL, args.pdl)
            e.set_option(sx.OPT_PREFETCH, args.prefetch)
            e.set_option(sx.OPT_PANEL_COLS, args.panel_cols)
            e.set_option(sx.OPT_HOST_GROUPS, args.host_groups)
            e.set_option(sx.OPT_HOST_FUSED, args.host_fused)
            e.set_option(sx.OPT_SLIDE, args.slide)
            if args.split >= 0:
                e.set_option(sx.OPT_SPLIT_ROW_NNZ, args.split)
            e.upload_csr(M, K, w["rowptr"], w["colidx"], w["val"])
            with torch.cuda.stream(stream):
                dB = e.device_B(N)[0] if own_B else torch.zeros(K * self.ld, dtype=self.tdtype, device=dev)   # device_B: zero-filled
                dCin = torch.zeros(M * self.ld, dtype=self.tdtype, device=dev)
                dCout = torch.zeros(M * self.ld, dtype=self.tdtype, device=dev)
                if fill_B:
                    e.colmajor_to_rowmajor(K, N, dB_cm, dB, self.ld)
                e.colmajor_to_rowmajor(M, N, dC_cm, dCin, self.ld)
            self.engines.append(e); self.dB.append(dB); self.dCin.append(dCin); self.dCout.append(dCout)
        del dB_cm, dC_cm
        stream.synchronize()

    def step(self, i):
        j = i % self.R
        self.engines[j].spmm_device(self.w["N"], ALPHA, self.dB[j], self.ld, BETA, self.dCin[j], self.dCout[j], self.ld)

    def launches(self):
        return sum(e.launches for e in self.engines)

    def kernel_name(self):
        import sextans_b200 as sx
        e = self.engines[0]
        lk = e.info(sx.INFO_LAST_KERNEL)
        fam = lk // 10000
        if fam == 5:
            return f"spmm_staged_kernel<WIN>, {lk % 10000} column-window passes"
        name = KERNEL_NAMES.get(fam, "?") + f"<G={lk % 10000 // 100},{'fast' if lk % 10 else 'strict'}>"
        if fam == 2:
            name += f" {e.info(sx.INFO_ITEMS)} items<={e.info(sx.INFO_ITEM_NNZ)}nnz, {e.info(sx.INFO_SPLIT_ROWS)} split rows"
        if fam == 8:
            name += f" {e.info(sx.INFO_EDGE_BLOCKS)} blocks staging {e.info(sx.INFO_EDGE_COLS)} B rows"
        if e.info(sx.INFO_TILE_NNZ) > 0:
            name += (f"; dense tiles {e.info(sx.INFO_TILE_NNZ)} nnz / {e.info(sx.INFO_TILE_SLOTS)} slots on "
                     f"spmm_panels_dmma_kernel, {e.info(sx.INFO_REST_NNZ)} nnz left to CSR")
        return name

    def result_colmajor(self, j=0):
        """C_out of copy j as the host holds it (column-major 1-D numpy)."""
        import torch
        M, N = self.w["M"], self.w["N"]
        with torch.cuda.stream(self.stream):
            out = torch.empty(M * N, dtype=self.tdtype, device=self.dev)
            self.engines[j].rowmajor_to_colmajor(M, N, self.dCout[j], self.ld, out)
        self.stream.synchronize()
        return out.cpu().numpy()

    def close(self):
        for e in self.engines:
            e.close()
        self.engines, self.dB, self.dCin, self.dCout = [], [], [], []


def time_steps(step, R, K, warmup, stream, min_region_ms, use_graph, world, dev, max_reps=4000, extra_streams=(), chain=True, tail=None):
    """Warm up, then repeat a replay until the timed region is >= min_region_ms; every repetition
    has its own event pair on `stream`.  Steps are numbered consecutively over warm-up and all
    repetitions (copy = step mod R), so the rotation through the R copies never restarts.
    Graph mode, chain=True: ONE graph holds every distinct K-step window of the rotation back to
    back (W windows = W*K steps, a whole number of passes over the R copies) and a repetition is one
    replay of it -- the steady state of a long chain of dependent SpMMs.  chain=False: one graph per
    K-step window, a repetition is one K-step replay (every K steps pay a graph launch and lose the
    overlap of the next launch's prologue across the graph boundary).
    `tail` (optional) is enqueued once behind every run of steps -- at the end of every captured graph, of
    the warm-up and of every eager repetition (the push exchange's flush of a deferred publication).
    -> dict(ms_median, ms_min, ms_mean (per step, max over ranks per repetition), reps, region_ms, graphs,
            steps_per_rep, nwarm)"""
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nwarm = max(3, warmup, R)        # every copy runs once outside a capture (its first SpMM builds its plan)
    with torch.cuda.stream(stream):
        for i in range(nwarm):
            step(i)
        if tail is not None:
            tail()
    barrier()
    base = nwarm
    graphs = None
    S = K                                            # steps per repetition
    if use_graph:
        nwin = min(R // math.gcd(R, K) if R > 1 else 1, 64)
        per_graph = nwin if chain else 1             # K-step windows per graph
        graphs = []
        for g in range(nwin // per_graph):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                for x in extra_streams:
                    x.wait_stream(stream)            # fork
                for i in range(per_graph * K):
                    step(base + g * per_graph * K + i)
                if tail is not None:
                    tail()
                for x in extra_streams:
                    stream.wait_stream(x)            # join
            graphs.append(gr)
        with torch.cuda.stream(stream):
            for gr in graphs:                        # one untimed pass over the whole rotation
                gr.replay()
        barrier()
        base += nwin * K
        S = per_graph * K

    def one_rep(r):
        if graphs is not None:
            graphs[r % len(graphs)].replay()
        else:
            for i in range(K):
                step(base + r * K + i)
            if tail is not None:
                tail()

    # pilot: three repetitions to size the run
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for r in range(3):
            one_rep(r)
        ev[1].record(stream)
    barrier()
    t = torch.tensor([max(ev[0].elapsed_time(ev[1]) / 3, 1e-4)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    reps = int(min(max_reps, max(3, math.ceil(min_region_ms / float(t.item())))))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    barrier()
    with torch.cuda.stream(stream):
        for r in range(reps):
            ev[r].record(stream)
            one_rep(3 + r)
        ev[reps].record(stream)
    barrier()
    per_rep = torch.tensor([ev[r].elapsed_time(ev[r + 1]) for r in range(reps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(per_rep, op=dist.ReduceOp.MAX)
    per_rep = per_rep.cpu().numpy()
    return {"ms_median": float(np.median(per_rep)) / S, "ms_min": float(per_rep.min()) / S,
            "ms_mean": float(per_rep.mean()) / S, "reps": reps, "region_ms": float(per_rep.sum()),
            "graphs": 0 if graphs is None else len(graphs), "steps_per_rep": S, "nwarm": nwarm}


def parity(case, w, rows0=0, rows1=None, threads=None):
    """One step on copy 0 against the CPU oracle over every row of this case (rows0:rows1 of the
    full workload `w` when the case is a row block).  Untimed.  -> dict"""
    import oracle
    M, K, N = w["M"], w["K"], w["N"]
    rows1 = M if rows1 is None else rows1
    if threads is None:
        threads = max(1, oracle.lib().sx_oracle_max_threads())
    with _on(case.stream):
        case.step(0)
    got = case.result_colmajor(0)
    rp = w["rowptr"]
    j0, j1 = int(rp[rows0]), int(rp[rows1])
    m = rows1 - rows0
    brp = (rp[rows0:rows1 + 1] - rp[rows0]).astype(np.int32)
    Cin = np.ascontiguousarray(w["Cin"].reshape(N, M)[:, rows0:rows1]).ravel()
    ref = oracle.spmm_csr(m, N, K, brp, w["colidx"][j0:j1], w["val"][j0:j1], w["dtype"].type(ALPHA), w["B"],
                          w["dtype"].type(BETA), Cin, threads=threads)
    return {"max_rel_err": max_rel_err(got, ref), "scaled_err": scaled_err(got, ref),
            "bit_exact": bool(np.array_equal(got.view(np.uint8), ref.view(np.uint8))),
            "rows": m, "checksum": float(np.asarray(got, dtype=np.float64).sum()),
            "checksum_oracle": float(np.asarray(ref, dtype=np.float64).sum())}


def run_batched(w, args, dev, stream, peak, nb):
    """`nb` SpMMs of the headline workload in ONE launch (sx_spmm_device_batch_*):
    nb distinct (B, C_in, C_out) triples with the same A.  Cold operands: T >= nb triples, T * bytes >
    2.2 x L2, call i takes triples [i*nb mod T, ...); A itself is the one matrix every triple shares.
    Every triple of the last call is compared with the oracle's result (they hold the same B and C_in)."""
    import torch
    import sextans_b200 as sx
    import oracle
    M, K, N, nnz, dtype = w["M"], w["K"], w["N"], w["nnz"], w["dtype"]
    s = dtype.itemsize
    td = torch.float64 if dtype == np.float64 else torch.float32
    ld = (N + 7) // 8 * 8
    sB, sC = K * ld, M * ld
    per = (sB + 2 * sC) * s
    T = max(nb, int(math.ceil(2.2 * L2_BYTES / per)))
    T = (T + nb - 1) // nb * nb
    e = sx.Engine(dev.index, arith=sx.STRICT if args.arith == "strict" else sx.FAST)
    e.set_stream(stream.cuda_stream)
    e.set_option(sx.OPT_PDL, args.pdl)
    e.upload_csr(M, K, w["rowptr"], w["colidx"], w["val"])
    with torch.cuda.stream(stream):
        dB = torch.zeros(T * sB, dtype=td, device=dev)
        dCin = torch.zeros(T * sC, dtype=td, device=dev)
        dCout = torch.zeros(T * sC, dtype=td, device=dev)
        b_cm = torch.from_numpy(w["B"]).to(dev)
        c_cm = torch.from_numpy(w["Cin"]).to(dev)
        e.colmajor_to_rowmajor(K, N, b_cm, dB, ld)
        e.colmajor_to_rowmajor(M, N, c_cm, dCin, ld)
        dB.view(T, sB)[1:] = dB.view(T, sB)[0]
        dCin.view(T, sC)[1:] = dCin.view(T, sC)[0]
    stream.synchronize()
    ncalls = T // nb

    def call(i):
        o = (i % ncalls) * nb
        e.spmm_device_batch(N, nb, ALPHA, dB[o * sB:], ld, sB, BETA, dCin[o * sC:], dCout[o * sC:], ld, sC)
    t = time_steps(call, ncalls, ncalls, args.warmup, stream, min(args.min_region_ms, 30.0), not args.no_graph, 1, dev)
    l0 = e.info(sx.INFO_LAUNCHES)
    with torch.cuda.stream(stream):
        call(0)
    stream.synchronize()
    launches = e.info(sx.INFO_LAUNCHES) - l0
    kernel = e.info(sx.INFO_LAST_KERNEL)
    ref = oracle.spmm_csr(M, N, K, w["rowptr"], w["colidx"], w["val"], dtype.type(ALPHA), w["B"], dtype.type(BETA), w["Cin"].copy(),
                          threads=max(1, oracle.lib().sx_oracle_max_threads()))
    ref_rm = np.zeros((M, ld), dtype)
    ref_rm[:, :N] = ref.reshape(N, M).T
    got = dCout.view(T, M, ld).cpu().numpy()
    bit_exact = bool(all(np.array_equal(got[b].view(np.uint8), ref_rm.view(np.uint8)) for b in range(T)))
    e.close()
    ms = t["ms_median"] / nb
    alg = per + nnz * (4 + s) / nb                      # per SpMM: its own B, C_in, C_out; A once per batch
    return {"nb": nb, "ms_per_spmm": round(ms, 6), "gflops": round(2.0 * nnz * N / (ms * 1e-3) / 1e9, 1),
            "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4), "launches_per_batch": int(launches),
            "kernel_family": int(kernel // 10000), "triples": T, "bit_exact_every_triple": bit_exact,
            "kernel": KERNEL_NAMES.get(int(kernel // 10000), "?"),
            "note": f"{nb} distinct (B, C_in, C_out) triples per launch, same A; cold: {T} triples = {T * per / 1e6:.0f} MB rotate; "
                    "frac counts A once per batch"}


def _on(stream):
    import torch
    return torch.cuda.stream(stream)


def run_config(key, wname, N, args, dev, stream, world, rank, peak):
    """One `configs` entry.  world == 1: the whole matrix on this GPU.  world > 1: strong scaling,
    this rank's nnz-balanced row block, B broadcast from rank 0 inside every step."""
    import torch
    import torch.distributed as dist
    from sextans_b200 import workloads as wl
    t_start = time.perf_counter()
    w = build_workload(wname, N, "", args.scale, args.band)
    s = w["dtype"].itemsize
    alg_total = wl.algorithmic_bytes(w["M"], w["K"], w["nnz"], w["N"], s)
    flops = 2.0 * w["nnz"] * w["N"]
    traffic = ncu_traffic(w["name"], w["N"], "f64" if s == 8 else "f32")
    K = max(3, min(args.steps, 20))
    cargs = argparse.Namespace(**vars(args))
    if wname == "powerlaw_blocked" and not cargs.tiles:
        cargs.tiles = 4
    if world == 1:
        case = Case(w, cargs, dev, stream)
        t = time_steps(case.step, case.R, K, args.warmup, stream, min(args.min_region_ms, 30.0), not args.no_graph, 1, dev)
        par = parity(case, w)
        out = {"ms": round(t["ms_median"], 6), "gflops": round(flops / (t["ms_median"] * 1e-3) / 1e9, 1),
               "frac": round(alg_total / (t["ms_median"] * 1e-3) / 1e9 / peak, 4),
               "traffic_ratio": None if traffic is None else round(traffic / alg_total, 2),
               "kernel": case.kernel_name(), "parity": par["max_rel_err"], "bit_exact": par["bit_exact"],
               "checksum": par["checksum"], "reps": t["reps"], "copies": case.R}
        if w["name"] == "uniform":
            # the gather-aware bound next to the algorithmic one: every nonzero
            # pulls one N*s-byte B row, and with B >> L2 and uniform columns they come from HBM
            gather = w["nnz"] * w["N"] * s
            out["frac_gather_bound"] = round(gather / (t["ms_median"] * 1e-3) / 1e9 / peak, 4)
        case.close()
        if wname == "powerlaw_blocked" and cargs.tiles:
            # the same matrix through the CSR path alone, beside the dense-tile variant
            plain = argparse.Namespace(**vars(args))
            plain.tiles = 0
            case = Case(w, plain, dev, stream)
            t2 = time_steps(case.step, case.R, K, args.warmup, stream, min(args.min_region_ms, 30.0), not args.no_graph, 1, dev)
            par2 = parity(case, w)
            out["csr_path"] = {"ms": round(t2["ms_median"], 6), "frac": round(alg_total / (t2["ms_median"] * 1e-3) / 1e9 / peak, 4),
                               "kernel": case.kernel_name(), "parity": par2["max_rel_err"]}
            out["note"] = ("dense-tile FP64 tensor-core variant (SX_OPT_TILE_MIN_ROWS=4, mma.sync m8n8k4 f64) + CSR remainder; parity is "
                           "tolerance-level (the MMA's summation order), an explicit zero times a non-finite B entry would give NaN")
            case.close()
        out["wall_s"] = round(time.perf_counter() - t_start, 1)
        return out
    # ---- strong scaling over row blocks ----
    from sextans_b200.rowblock import RowBlock
    blk = RowBlock(w["M"], w["K"], w["rowptr"], w["colidx"], w["val"], world, rank)
    wb = dict(w, M=blk.rows, nnz=blk.nnz, rowptr=blk.rowptr, colidx=blk.colidx, val=blk.val, Cin=blk.take_C(w["Cin"], w["N"]))
    case = Case(wb, cargs, dev, stream, copies=1, fill_B=(rank == 0))
    with torch.cuda.stream(stream):
        dist.broadcast(case.dB[0], src=0)              # kernel-only timing needs B everywhere: one untimed broadcast
    stream.synchronize()
    tk = time_steps(case.step, 1, K, args.warmup, stream, min(args.min_region_ms, 20.0), False, world, dev)

    def step_x(i):
        dist.broadcast(case.dB[0], src=0)              # NCCL over NVLink, on `stream`
        case.step(i)
    tx = time_steps(step_x, 1, K, args.warmup, stream, min(args.min_region_ms, 20.0), False, world, dev)
    par = parity(case, w, blk.r0, blk.r1, threads=max(1, (os.cpu_count() or 8) // world))
    # a large B also travels in column panels, panel p+1 on the wire while panel p is multiplied
    piped = None
    if w["K"] * case.ld * s >= (192 << 20) and w["N"] * s > 128:
        from sextans_b200.rowblock import PanelPipeline
        import oracle
        pp = PanelPipeline(case.engines[0], blk.rows, w["K"], w["N"], w["dtype"], dev, stream)
        pp.load(w["B"] if rank == 0 else None, wb["Cin"])
        tp = time_steps(lambda i: pp.step(ALPHA, BETA), 1, K, args.warmup, stream, min(args.min_region_ms, 20.0), False, world, dev)
        got = pp.result()
        ref = oracle.spmm_csr(blk.rows, w["N"], w["K"], blk.rowptr, blk.colidx, blk.val, w["dtype"].type(ALPHA), w["B"],
                              w["dtype"].type(BETA), wb["Cin"].copy(), threads=max(1, (os.cpu_count() or 8) // world))
        piped = {"ms_step": round(tp["ms_median"], 6), "gflops": round(flops / (tp["ms_median"] * 1e-3) / 1e9, 1), "panels": pp.P,
                 "panel_cols": pp.pw, "max_rel_err": max_rel_err(got, ref),
                 "bit_exact": bool(np.array_equal(got.view(np.uint8), ref.view(np.uint8)))}
        par["max_rel_err"] = max(par["max_rel_err"], piped["max_rel_err"])
        par["bit_exact"] = par["bit_exact"] and piped["bit_exact"]
        del pp
    errs = torch.tensor([par["max_rel_err"], 0.0 if par["bit_exact"] else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    nnzs = torch.tensor([float(blk.nnz)], dtype=torch.float64, device=dev)
    allnnz = [torch.zeros_like(nnzs) for _ in range(world)]
    dist.all_gather(allnnz, nnzs)
    allnnz = [float(x.item()) for x in allnnz]
    per_gpu_alg = wl.algorithmic_bytes(blk.rows, w["K"], blk.nnz, w["N"], s)
    out = {"scaling": "strong", "ms_kernel": round(tk["ms_median"], 6), "ms_step": round(tx["ms_median"], 6),
           "gflops_kernel": round(flops / (tk["ms_median"] * 1e-3) / 1e9, 1),
           "gflops": round(flops / (tx["ms_median"] * 1e-3) / 1e9, 1),
           "frac_kernel_per_gpu": round(per_gpu_alg / (tk["ms_median"] * 1e-3) / 1e9 / peak, 4),
           "nnz_imbalance": round(max(allnnz) / (sum(allnnz) / world), 4),
           "exchange": f"ncclBroadcast of B ({w['K'] * case.ld * s / 1e6:.0f} MB) from rank 0 inside every step",
           "pipelined": None if piped is None else dict(piped, note="B broadcast and multiplied in column panels: panel p+1 on the wire (NCCL, side stream) while panel p is multiplied"),
           "kernel": case.kernel_name(), "parity_all_ranks": float(errs[0].item()),
           "bit_exact_all_ranks": bool(errs[1].item() == 0.0), "wall_s": round(time.perf_counter() - t_start, 1)}
    case.close()
    return out


def run_native(args):
    import torch
    import torch.distributed as dist
    import sextans_b200 as sx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    peak, peak_src = measured_peak()
    K = args.steps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline ----------------------------------------------------------------------------
    w = build_workload(args.workload, args.ncols, args.dtype, args.scale, args.band)
    M, Kc, N, nnz, dtype = w["M"], w["K"], w["N"], w["nnz"], w["dtype"]
    s = dtype.itemsize
    ld = (N + 7) // 8 * 8
    # N > 1 (weak scaling: every rank owns one stacked copy of the matrix): how B reaches the
    # other ranks inside every step.  Small B: pushed through peer memory; large B: one NCCL broadcast.
    push = world > 1 and Kc * ld * s <= args.peer_bytes
    case = Case(w, args, dev, stream, fill_B=(world == 1 or rank == 0), own_B=push)
    R = case.R
    alg_bytes = case.alg_bytes
    xch = None
    exchange = "none (1 GPU)"
    extra_streams = ()
    if world > 1:
        exchange = "one ncclBroadcast of B from rank 0 inside every step"
        if push:
            from sextans_b200.rowblock import PushExchange
            xch = PushExchange(case.engines, N)
            exchange = xch.describe()

    def step(i):
        if xch is not None:
            xch.before_step(i)
        elif world > 1:
            dist.broadcast(case.dB[i % R], src=0)
        case.step(i)

    sampler = ClockSampler(local)
    sampler.start()
    use_graph = not args.no_graph and (world == 1 or xch is not None)   # NCCL collectives are launched eagerly
    l0 = case.launches()
    tail = xch.flush if xch is not None else None
    t = time_steps(step, R, K, args.warmup, stream, args.min_region_ms, use_graph, world, dev, extra_streams=extra_streams, tail=tail)
    # launches in the timed region: a captured graph holds the launches of its K steps and is replayed once per repetition
    S = t["steps_per_rep"]
    enq_steps = t["nwarm"] + (t["graphs"] * S if use_graph else (3 + t["reps"]) * K)
    per_step = (case.launches() - l0) / max(1, enq_steps)
    launches_dev = int(round(per_step * S * t["reps"]))
    kern_ms = t["ms_median"]
    # the same steps replayed as separate K-step graphs (what round 1 and the first half of round 2 reported)
    t_k = time_steps(step, R, K, 0, stream, min(args.min_region_ms, 20.0), True, world, dev, extra_streams=extra_streams, chain=False, tail=tail) if use_graph and S > K else None
    if t_k is not None:
        launches_dev += int(round(per_step * K * t_k["reps"]))
    timeouts = case.engines[0].info(sx.INFO_EXCHANGE_TIMEOUTS) if world > 1 else 0

    # the same on ONE copy (L2 warm when the working set fits), for comparison
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nwarm_steps = 200 if alg_bytes < L2_BYTES else 5
    warm_ms = None
    if world == 1:
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(nwarm_steps):
                case.engines[0].spmm_device(N, ALPHA, case.dB[0], ld, BETA, case.dCin[0], case.dCout[0], ld)
            e1.record(stream)
        barrier()
        warm_ms = e0.elapsed_time(e1) / nwarm_steps

    # parity of the headline step on every rank (each owns one copy of the matrix; after the timed
    # region every rank's image 0 holds the B that rank 0 sent)
    barrier()
    par = parity(case, w, threads=max(1, (os.cpu_count() or 8) // world))
    perr = torch.tensor([par["max_rel_err"], 0.0 if par["bit_exact"] else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(perr, op=dist.ReduceOp.MAX)

    kernel_name = case.kernel_name()
    # ---- e2e: host-facing call, pinned host buffers ------------------------------------------
    hB = sx.pinned_empty(Kc * N, dtype)
    hC = sx.pinned_empty(M * N, dtype)
    hB[:] = w["B"]
    e2e_note = "host wall clock around the blocking sx_spmm_* call (H2D of B and C_in, kernels, D2H of C)"
    sharded = None
    eng = case.engines[0]
    if world > 1:
        from sextans_b200.rowblock import ShardedSpMM
        if xch is not None:
            xch.close()
            xch = None
        # the stacked matrix: rank r's nnz-balanced row block IS copy r
        sharded = ShardedSpMM.around(eng, w["val"].dtype, M, Kc, stream, dev.index, peer_bytes=args.peer_bytes)
        e2e_note = "host wall clock around ShardedSpMM.spmm: B on rank 0's host -> staged, exchanged, every rank's C block in and out over PCIe"

    def e2e_call():
        if sharded is not None:
            sharded.spmm(N, ALPHA, hB if rank == 0 else None, BETA, hC, want_ns=False)
        else:
            eng.spmm(N, ALPHA, hB, BETA, hC, want_ns=False)     # kernel_ns = NULL: nobody needs the kernel-only time here
    for _ in range(3):
        hC[:] = w["Cin"]
        e2e_call()
    checksum = float(np.asarray(hC, dtype=np.float64).sum())
    barrier()
    l2 = case.launches()
    e2e_times = []
    e2e_steps = max(K, 50)
    for _ in range(e2e_steps):
        hC[:] = w["Cin"]                       # restore the in/out operand (untimed)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_call()                             # returns synchronised
        e2e_times.append(time.perf_counter() - t0)
    launches_e2e = case.launches() - l2
    barrier()
    # ---- the same call with TWO calls in flight (1 GPU): two contexts on two streams, double-buffered page-locked
    # operands, sx_spmm_enqueue_* + sx_synchronize -- call i's results leave over PCIe while call i+1's operands arrive.
    # Every step still moves its own B and C_in in and its C out; C evolves in place (no host restore between steps),
    # and buffer 0 is checked against the oracle applied as many times as it was used.
    pipelined = None
    if world == 1 and not args.no_pipelined_e2e:
        try:
            import oracle
            e2 = sx.Engine(dev.index, arith=sx.STRICT if args.arith == "strict" else sx.FAST)
            e2.set_option(sx.OPT_HOST_FUSED, args.host_fused)
            e2.set_option(sx.OPT_HOST_GROUPS, args.host_groups)
            e2.upload_csr(M, Kc, w["rowptr"], w["colidx"], w["val"])
            eng.set_stream(0)                                   # its own stream again (the timed steps ran on torch's)
            pe = [eng, e2]
            pB = [hB, sx.pinned_empty(Kc * N, dtype)]
            pC = [hC, sx.pinned_empty(M * N, dtype)]
            pB[1][:] = w["B"]

            def run(nsteps):
                for j in range(2):
                    pC[j][:] = w["Cin"]
                t0 = time.perf_counter()
                for i in range(nsteps):
                    j = i & 1
                    if i >= 2:
                        pe[j].synchronize()                     # buffer j's previous result is in host memory
                    pe[j].spmm_enqueue(N, ALPHA, pB[j], BETA, pC[j])
                pe[0].synchronize()
                pe[1].synchronize()
                return time.perf_counter() - t0
            run(6)
            nst = 40
            times = sorted(run(nst) for _ in range(7))
            ref = w["Cin"].copy()
            for _ in range(nst // 2):
                ref = oracle.spmm_csr(M, N, Kc, w["rowptr"], w["colidx"], w["val"], dtype.type(ALPHA), w["B"], dtype.type(BETA), ref,
                                      threads=max(1, oracle.lib().sx_oracle_max_threads()))
            ok = bool(np.array_equal(np.asarray(pC[0]).view(np.uint8), ref.view(np.uint8)))
            pms = times[len(times) // 2] / nst * 1e3
            pipelined = {"value": 2.0 * nnz * N / (pms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": pms, "calls_in_flight": 2,
                         "steps": nst, "bit_exact_after_chain": ok, "host_path": e2.info(sx.INFO_HOST_PATH),
                         "note": "two contexts on two streams, double-buffered page-locked B and C, sx_spmm_enqueue_* + sx_synchronize per buffer; "
                                 "every step moves its own operands in and its result out; median of 7 runs of 40 steps"}
            launches_e2e += 7 * nst + 6
            e2.close()
        except Exception as ex:
            pipelined = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    clocks = sampler.result()
    e2e_t = torch.tensor(e2e_times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.median().item()) * 1e3
    host_path = {1: "zero-copy kernels over PCIe (no memcpy): B/C_in staging, SpMM, C out", 2: "zero-copy, two launches: B staging, then the SpMM kernel reads C_in from and writes C to the caller's array itself",
                 3: "zero-copy, ONE launch: blocks fetch their share of B and their C_in tile from the caller's arrays (cp.async over PCIe), exchange B through L2 behind a counter, and write C back, pipelined over column groups so that both directions of the link are busy"}.get(
        eng.info(sx.INFO_HOST_PATH), "cudaMemcpyAsync + layout kernels")
    if sharded is not None:
        host_path = f"ShardedSpMM ({sharded.last_exchange} exchange of B)"
        sharded.close(keep_engine=True)
    case.close()
    del hB, hC

    # ---- the other BASELINE configs ------------------------------------------------------------
    configs = {}
    pool = (CONFIGS_1GPU + CONFIGS_EXTRA) if world == 1 else CONFIGS_NGPU
    if args.configs == "none":
        wanted = []
    elif args.configs == "all":
        wanted = CONFIGS_1GPU if world == 1 else CONFIGS_NGPU
    else:
        keys = set(args.configs.split(","))
        wanted = [c for c in pool if c[0] in keys]
    for key, wname, n in wanted:
        try:
            configs[key] = run_config(key, wname, n, args, dev, stream, world, rank, peak)
        except Exception as ex:                                  # one config must not take the line down
            if world > 1:
                raise
            configs[key] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    batched = None
    if world == 1 and args.batch > 1:
        try:
            batched = run_batched(w, args, dev, stream, peak, args.batch)
        except Exception as ex:
            batched = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    flops_step = 2.0 * nnz * N * world
    value = flops_step / (kern_ms * 1e-3) / 1e9
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9        # per GPU: every rank moves its own copy's bytes
    if rank == 0:
        line = {
            "metric": "SpMM GFLOP/s (2*nnz*N)", "value": value, "unit": "GFLOP/s",
            "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if dtype == np.float64 else "f32",
            "data": data_label(w),
            "configs": configs,
            "batched": batched,
            "parity": {"max_rel_err_all_ranks": float(perr[0].item()), "bit_exact_all_ranks": bool(perr[1].item() == 0.0),
                       "rows_checked_per_rank": M, "against": "oracle.spmm_csr (cpu_spmm_CSR restated), every row",
                       "exchange_timeouts": int(timeouts)},
            "config": config_of(w),
            "run": {"arith": args.arith,
                    "l2": "warm (--no-flush)" if args.no_flush else (
                        f"cold: {R} device copies of A/B/C ({R * alg_bytes / 1e6:.0f} MB > 2 x 126 MB L2), step i uses copy i mod {R} across all replays"
                        if R > 1 else f"cold: one copy is {alg_bytes / 1e6:.0f} MB > 2 x L2"),
                    "timed": f"{t['reps']} repetitions x {S} steps = {t['region_ms']:.1f} ms; ms_per_step = median repetition / {S} (min {t['ms_min'] * 1e3:.3f} us, mean {t['ms_mean'] * 1e3:.3f} us)",
                    "launch": (f"one CUDA graph of {S} steps = {S // K} windows of K = {K} steps back to back (a whole number of passes over the copies); one replay per repetition"
                               if use_graph and S > K else f"{t['graphs']} CUDA graphs of {K} steps" if use_graph else "one by one"),
                    "k_step_graphs": None if t_k is None else {
                        "ms_per_step": t_k["ms_median"], "reps": t_k["reps"],
                        "note": f"the same steps as separate graphs of K = {K} steps, one replay per repetition: every K steps pay a graph launch and the first step of a graph cannot overlap its prologue with the previous kernel"},
                    "partition": "1 row block" if world == 1 else f"{world} stacked copies, one row block per GPU",
                    "exchange": exchange},
            "gflops_ref_formula": 2.0 * (nnz + M) * N * world / (kern_ms * 1e-3) / 1e9,
            "single_copy_back_to_back": None if warm_ms is None else {
                "ms_per_step": warm_ms, "gflops": 2.0 * nnz * N / (warm_ms * 1e-3) / 1e9, "note": "one copy, launched one by one (L2-warm when it fits)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(w["name"], N, "f64" if s == 8 else "f32"),
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src, "kernel": kernel_name},
            "e2e": {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": (Kc * N + M * N * world) * s, "d2h_bytes_per_step": M * N * s * world,
                    "timer": e2e_note, "path": host_path, "steps": e2e_steps, "stat": "median, max over ranks per step",
                    "pipelined": pipelined},
            "gpu_launches": int(launches_dev + launches_e2e),
            "gpu_launches_detail": {"timed_device_steps": launches_dev, "e2e_steps": int(launches_e2e)},
            "clocks": clocks, "checksum_C": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            line["cpu_baseline"] = cpu_baseline(w, max(1, oracle.lib().sx_oracle_max_threads()))
            line["cpu_baseline"]["one_thread"] = cpu_baseline(w, 1, budget_s=5.0)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library writes to
    file descriptor 1 (NCCL prints its version banner there) has been sent to stderr."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
