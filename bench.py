#!/usr/bin/env python
"""bench.py -- SpMM GFLOP/s (2*nnz*N) and achieved HBM GB/s against the roofline.

A "step" is one SpMM  C = alpha*A*B + beta*C  over the named workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Workloads (BASELINE.json configs):
  nasa4704   nasa4704.mtx  N=16 fp64   (configs[1]; the default)
  pcrystk02  pcrystk02.mtx N=16 fp32   (configs[2]; --ncols 8|16|32|64)
  uniform    synthetic M=K=1e6, 20 nnz/row, N=128 fp32      (configs[3])
  powerlaw   synthetic power-law M=K=1e6, nnz~1e8, N=16 fp64 (configs[4])
  fem        synthetic block-structured (4 dof/node) M=K=1e6, nnz~9.5e7, N=16 fp64: the
             input of the dense-tile tensor-core variant (configs[4] tail; --tiles 4)

Own arm, per rank (one process per GPU): A's row block resident on the device; at N>1
the matrix is N row blocks of the workload stacked (weak scaling), every rank owns one,
and each step starts with the NCCL broadcast of B from rank 0 over NVLink.
  value : steps timed with CUDA events on the launching stream, inputs resident in
          HBM, L2 flushed (a 512 MB buffer is overwritten) before every step
  e2e   : the same step through the host-facing C-ABI call sx_spmm_* with pinned host
          B and C (column-major, as the host program holds them): H2D copies, layout
          change, kernel, layout change, D2H copy -- all inside the timed region
Reference arm (--impl reference): the CPU path (the reference's cpu_spmm_CSR when the
run is fp32 and oracle/_ref is built, else the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
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
    "fem": ("fem", np.float64, 16),
}
ALPHA, BETA = float(np.float32(0.85)), float(np.float32(-2.06))   # host.cpp:29-30


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--workload", default="nasa4704", choices=sorted(WORKLOADS))
    p.add_argument("--ncols", type=int, default=0, help="override the workload's N")
    p.add_argument("--dtype", default="", choices=["", "f32", "f64"])
    p.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic workloads (testing)")
    p.add_argument("--arith", default="strict", choices=["strict", "fast"])
    p.add_argument("--kernel", type=int, default=0, help="SX_OPT_KERNEL (0 auto, 1 row per lane group, 2 TMA-staged items, 3 TMA-staged B window, 4 sliding B window: experimental, needs --slide)")
    p.add_argument("--item-nnz", type=int, default=0, help="SX_OPT_ITEM_NNZ (0 auto)")
    p.add_argument("--split", type=int, default=-1, help="SX_OPT_SPLIT_ROW_NNZ (-1 default)")
    p.add_argument("--band", type=int, default=2000, help="fem workload: couplings reach +-band nodes")
    p.add_argument("--tiles", type=int, default=0, help="SX_OPT_TILE_MIN_ROWS (fp64 dense-tile tensor-core variant; 0 off)")
    p.add_argument("--col-window-rows", type=int, default=0, help="SX_OPT_COL_WINDOW_ROWS (column-window passes keeping a window of B in L2; 0 off, -1 = 32 MiB of B per window)")
    p.add_argument("--autotune", action="store_true", help="SX_OPT_AUTOTUNE (experimental): time the applicable variants on the first call and keep the fastest")
    p.add_argument("--slide", type=int, default=0, help="SX_OPT_SLIDE (experimental): chains per SM of the sliding-window kernel; use with --kernel 4")
    p.add_argument("--window-rows", type=int, default=0, choices=[0, 32, 64, 128], help="SX_OPT_WINDOW_ROWS (experimental): rows per block of variant 3")
    p.add_argument("--pdl", type=int, default=-1, choices=[-1, 0, 1], help="SX_OPT_PDL: programmatic dependent launch (-1 auto: on for the edge-list kernel)")
    p.add_argument("--prefetch", type=int, default=-1, choices=[-1, 0, 1], help="SX_OPT_PREFETCH: L2 prefetch hints (-1 auto)")
    p.add_argument("--host-fused", action="store_true", help="SX_OPT_HOST_FUSED (experimental): e2e calls pass kernel_ns=NULL and the SpMM kernel carries C across PCIe")
    p.add_argument("--ref-threads", type=int, default=1, help="--impl reference: threads of the CPU path (1 = as the reference runs it; -1 = all cores, OpenMP port)")
    p.add_argument("--peer-bytes", type=int, default=8 << 20, help="N>1: B images up to this size travel by peer copy instead of NCCL")
    p.add_argument("--peer-mode", default="fused", choices=["fused", "memops"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-flush", action="store_true", help="leave L2 warm between steps")
    p.add_argument("--no-graph", action="store_true", help="launch the timed steps one by one instead of replaying a CUDA graph")
    return p.parse_args()


def build_workload(args):
    """-> dict(name, M, K, nnz, N, dtype, rowptr, colidx, val, B, Cin) for ONE row block."""
    import sextans_b200 as sx
    from sextans_b200 import workloads as wl
    kind, dtype, N = WORKLOADS[args.workload]
    if args.dtype:
        dtype = np.float32 if args.dtype == "f32" else np.float64
    if args.ncols:
        N = args.ncols
    if kind == "suitesparse":
        M, K, nnz, rp, ci, v = sx.load_mtx(wl.suitesparse_path(args.workload), dtype)
        B, Cin = wl.host_dense(M, K, N, dtype)
        desc = f"{args.workload}.mtx M=K={M} nnz={nnz} N={N} {np.dtype(dtype).name}, B=1, C_in=(m+1)(n+1)/M/N (host.cpp:100-111)"
    elif kind == "uniform":
        M = K = max(1000, int(1_000_000 * args.scale))
        rp, ci, v = wl.uniform_csr(M, K, 20, 12345, dtype)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        desc = f"synthetic uniform CSR M=K={M} nnz={nnz} (20/row) N={N} {np.dtype(dtype).name}, seed 12345"
    elif kind == "fem":
        nodes = max(250, int(250_000 * args.scale))
        M = K = nodes * 4
        rp, ci, v = wl.fem_like_csr(nodes, 4, 23, 12345, dtype, band=args.band)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        desc = f"synthetic FEM-like CSR (4 dof/node, dense 4x4 couplings within +-{args.band} nodes) M=K={M} nnz={nnz} N={N} {np.dtype(dtype).name}, seed 12345"
    else:
        M = K = max(1000, int(1_000_000 * args.scale))
        rp, ci, v = wl.powerlaw_csr(M, K, int(100_000_000 * args.scale), 12345, dtype)
        nnz = int(ci.size)
        B, Cin = wl.random_dense(M, K, N, 12345, dtype)
        desc = f"synthetic power-law CSR M=K={M} nnz={nnz} N={N} {np.dtype(dtype).name}, seed 12345"
    return dict(name=args.workload, desc=desc, M=M, K=K, nnz=nnz, N=N, dtype=np.dtype(dtype),
                rowptr=rp, colidx=ci, val=v, B=B, Cin=Cin)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name, N, dtype):
    """dram bytes per launch of the SpMM kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{name}_n{N}_{dtype}")
    except Exception:
        return None


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
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(w, threads, budget_s=12.0, min_runs=3):
    """Time the CPU path on this workload (checker code, timed as a reported baseline).
    fp32 + oracle/_ref present -> the reference's own cpu_spmm_CSR (1 thread, as written);
    otherwise the oracle port, `threads` OpenMP threads over rows."""
    import oracle
    M, K, N, nnz = w["M"], w["K"], w["N"], w["nnz"]
    rows_sample = None
    # bound the work: if one full SpMM would take more than the budget (~2 GFLOP/s per
    # thread), time a contiguous prefix of the rows with the same B
    est = 2.0 * nnz * N / (1.5e9 * max(1, threads))
    rp, ci, v = w["rowptr"], w["colidx"], w["val"]
    Cin = w["Cin"]
    if est > budget_s:
        frac = budget_s / est
        Ms = max(1, int(M * frac))
        rows_sample = Ms
        rp = np.ascontiguousarray(rp[:Ms + 1])
        ci, v = ci[:rp[-1]], v[:rp[-1]]
        Cin = np.ascontiguousarray(Cin.reshape(N, M)[:, :Ms]).ravel()
        M = Ms
        nnz = int(rp[-1])
    use_ref = (w["dtype"] == np.float32 and threads == 1 and oracle.ref() is not None)
    times = []
    t_all = time.perf_counter()
    while len(times) < min_runs or (time.perf_counter() - t_all < budget_s / 2 and len(times) < 20):
        C = Cin.copy()
        t0 = time.perf_counter()
        if use_ref:
            oracle.ref_spmm_csr(M, N, K, rp, ci, v, ALPHA, w["B"], BETA, C)
        else:
            oracle.spmm_csr(M, N, K, rp, ci, v, w["dtype"].type(ALPHA), w["B"], w["dtype"].type(BETA), C, threads=threads)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s * 2:
            break
    best = min(times)
    sample = "the whole workload" if rows_sample is None else f"first {rows_sample} rows ({nnz} nnz) of the workload, full B"
    return {"value": 2.0 * nnz * N / best / 1e9, "unit": "GFLOP/s", "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": f"{sample}; best of {len(times)} runs, {best * 1e3:.3f} ms",
            "ms": best * 1e3, "mean_ms": float(np.mean(times)) * 1e3}


def run_reference(args):
    """The reference arm: cpu_spmm_CSR as the reference runs it -- ONE thread, the function
    has no threading (src/sparse_helper.h:262-290) -- through oracle/_ref (the reference's
    own header, compiled unmodified) for fp32 and the line-for-line double port for fp64.
    --ref-threads N times the port's row-parallel OpenMP variant instead; the default line
    carries that all-core figure as extra information."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    w = build_workload(args)
    all_threads = max(1, oracle.lib().sx_oracle_max_threads())
    threads = all_threads if args.ref_threads < 0 else max(1, args.ref_threads)
    M, K, N, nnz = w["M"], w["K"], w["N"], w["nnz"]
    est = 2.0 * nnz * N / (1.5e9 * threads)
    rp, ci, v, Cin = w["rowptr"], w["colidx"], w["val"], w["Cin"]
    sample = "the whole workload per step"
    if est > 2.0:   # bounded sample: a row prefix sized to ~2 s per step
        Ms = max(1, int(M * 2.0 / est))
        rp = np.ascontiguousarray(rp[:Ms + 1])
        ci, v = ci[:rp[-1]], v[:rp[-1]]
        Cin = np.ascontiguousarray(Cin.reshape(N, M)[:, :Ms]).ravel()
        M, nnz = Ms, int(rp[-1])
        sample = f"first {Ms} rows ({nnz} nnz) of the workload per step, full B"
    a, b = w["dtype"].type(ALPHA), w["dtype"].type(BETA)
    use_ref = w["dtype"] == np.float32 and threads == 1 and oracle.ref() is not None

    def one(C):
        if use_ref:
            oracle.ref_spmm_csr(M, N, K, rp, ci, v, ALPHA, w["B"], BETA, C)
        else:
            oracle.spmm_csr(M, N, K, rp, ci, v, a, w["B"], b, C, threads=threads)

    steps = args.steps
    if est * steps > 120:   # keep the whole run within a few minutes
        steps = max(3, int(120 / est))
    for _ in range(min(args.warmup, 3)):
        one(Cin.copy())
    total = 0.0
    for _ in range(steps):
        C = Cin.copy()
        t0 = time.perf_counter()
        one(C)
        total += time.perf_counter() - t0
    val = 2.0 * nnz * N * steps / total / 1e9
    extra = None
    if threads == 1 and all_threads > 1:
        extra = cpu_baseline(dict(w, M=M, nnz=nnz, rowptr=rp, colidx=ci, val=v, Cin=Cin), all_threads, budget_s=6.0)
        extra["note"] = "row-parallel OpenMP over the oracle port (bitwise the same result); NOT the reference, which is single-threaded"
    line = {
        "impl": "reference", "metric": "SpMM GFLOP/s (2*nnz*N)", "value": val, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
        "ms_per_step": total / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if w["dtype"] == np.float64 else "f32",
        "data": "synthetic" if w["name"] in ("uniform", "powerlaw") else "SuiteSparse fixture shipped with the reference, host program's B/C",
        "config": {"workload": w["desc"], "alpha": ALPHA, "beta": BETA},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "reference" if use_ref else "port",
                         "sample": sample,
                         "note": "cpu_spmm_CSR as the reference runs it: one thread (src/sparse_helper.h:262-290 has no threading)"
                                 if threads == 1 else "row-parallel OpenMP over the oracle port of cpu_spmm_CSR",
                         "all_cores_openmp_port": extra},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_native(args):
    import torch
    import torch.distributed as dist
    import sextans_b200 as sx
    from sextans_b200 import workloads as wl

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

    w = build_workload(args)
    M, K, N, nnz, dtype = w["M"], w["K"], w["N"], w["nnz"], w["dtype"]
    tdtype = torch.float64 if dtype == np.float64 else torch.float32
    s = dtype.itemsize

    # ---- replicas: "inputs larger than L2" ------------------------------------------------
    # The timed steps run back to back with no flush kernel between them; instead the
    # operands (A, B, C_in, C_out) exist in R independent device copies whose total size
    # is > 2x the 126 MB L2, and step i uses copy i mod R, so every step finds its data
    # in HBM, not in L2.  R = 1 when one copy alone is that large.
    alg_bytes = wl.algorithmic_bytes(M, K, nnz, N, s)
    L2_BYTES = 126 * 1024 * 1024
    R = 1 if (args.no_flush or alg_bytes >= 2 * L2_BYTES) else int(np.ceil(2.2 * L2_BYTES / alg_bytes))
    stream = torch.cuda.Stream(device=dev)
    ld = (N + 7) // 8 * 8
    engines, dBs, dCins, dCouts = [], [], [], []
    with torch.cuda.stream(stream):
        dB_cm = torch.from_numpy(w["B"]).to(dev)
        dC_cm = torch.from_numpy(w["Cin"]).to(dev)
    for i in range(R):
        e = sx.Engine(local, arith=sx.STRICT if args.arith == "strict" else sx.FAST)
        e.set_stream(stream.cuda_stream)
        e.set_option(sx.OPT_KERNEL, args.kernel)
        e.set_option(sx.OPT_ITEM_NNZ, args.item_nnz)
        e.set_option(sx.OPT_TILE_MIN_ROWS, args.tiles)
        cw = args.col_window_rows
        if cw < 0:  # a window of B of ~32 MiB: well inside one L2 partition
            cw = max(1, (32 << 20) // (ld * s))
        e.set_option(sx.OPT_COL_WINDOW_ROWS, cw)
        e.set_option(sx.OPT_HOST_FUSED, 1 if args.host_fused else 0)
        e.set_option(sx.OPT_PDL, args.pdl)
        e.set_option(sx.OPT_PREFETCH, args.prefetch)
        e.set_option(sx.OPT_WINDOW_ROWS, args.window_rows)
        e.set_option(sx.OPT_SLIDE, args.slide)
        e.set_option(sx.OPT_AUTOTUNE, 1 if args.autotune else 0)
        if args.split >= 0:
            e.set_option(sx.OPT_SPLIT_ROW_NNZ, args.split)
        e.upload_csr(M, K, w["rowptr"], w["colidx"], w["val"])
        with torch.cuda.stream(stream):
            dB = torch.zeros(K * ld, dtype=tdtype, device=dev)
            dCin = torch.zeros(M * ld, dtype=tdtype, device=dev)
            dCout = torch.zeros(M * ld, dtype=tdtype, device=dev)
            e.colmajor_to_rowmajor(K, N, dB_cm, dB, ld)
            e.colmajor_to_rowmajor(M, N, dC_cm, dCin, ld)
            if world > 1 and rank != 0:
                dB.zero_()          # non-root ranks receive B through the broadcast
        engines.append(e); dBs.append(dB); dCins.append(dCin); dCouts.append(dCout)
    eng = engines[0]
    stream.synchronize()

    # N > 1: how B reaches the other ranks.  Small B: every rank pulls the root's image with a
    # peer copy over NVLink ordered by device-side step counters (PeerBroadcast); large B
    # (or if CUDA IPC is not available): one NCCL broadcast.
    peer = None
    exchange = "none"
    if world > 1:
        exchange = "one NCCL broadcast of B from rank 0 inside every step"
        if K * ld * s <= args.peer_bytes:
            try:
                from sextans_b200.rowblock import PeerBroadcast
                for j in range(R):                      # the engines' own B images are the operands here
                    ptr, _ = engines[j].device_B(N)          # zero-filled; only the root holds B
                    if rank == 0:
                        engines[j].colmajor_to_rowmajor(K, N, dB_cm, ptr, ld)
                    dBs[j] = ptr
                stream.synchronize()
                peer = PeerBroadcast(engines, N, fused=(args.peer_mode == "fused"))
                exchange = ("B pulled from rank 0 over NVLink by one fused kernel per step (spin on the step flag, copy, acknowledge)" if args.peer_mode == "fused" else "peer copy of B from rank 0 over NVLink (copy engine) ordered by stream memory operations") + ", inside every step"
            except Exception as ex:                     # no IPC in this sandbox: keep NCCL
                peer = None
                exchange += f" (peer path unavailable: {type(ex).__name__})"
    step_no = [0]
    copy_stream = torch.cuda.Stream(device=dev) if peer is not None else None

    def step_device(i):
        j = i % R
        if peer is not None:
            # the pull of step k runs on its own stream, so it overlaps the SpMM of step k-1
            step_no[0] += 1
            assert (step_no[0] - 1) % R == j
            if rank == 0:
                # B is resident and constant here, so publishing step k does not depend on the
                # SpMM stream: it runs beside it (in a pipeline it would follow B's producer)
                engines[j].set_stream(copy_stream.cuda_stream)
                peer.publish(step_no[0])
                engines[j].set_stream(stream.cuda_stream)
            else:
                engines[j].set_stream(copy_stream.cuda_stream)
                peer.pull(step_no[0])
                engines[j].set_stream(stream.cuda_stream)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                stream.wait_event(ev)
        elif world > 1:
            dist.broadcast(dBs[j], src=0)
        engines[j].spmm_device(N, ALPHA, dBs[j], ld, BETA, dCins[j], dCouts[j], ld)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        return sum(e.launches for e in engines)

    nwarm = max(3, args.warmup, R)     # every copy is touched (and its plan built) before timing
    with torch.cuda.stream(stream):
        for i in range(nwarm):
            step_device(i)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed: exactly K steps between two events on the launching stream ------------
    # The K steps are captured once into a CUDA graph (launch-bound loop: a nasa4704 SpMM
    # is ~10 us of device work) and the graph is replayed inside the timed region;
    # --no-graph launches them one by one instead.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launches()
    graph = None
    use_graph = not args.no_graph and (world == 1 or peer is not None)   # NCCL collectives are launched eagerly
    if use_graph:
        def capture():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                if copy_stream is not None:
                    copy_stream.wait_stream(stream)          # fork
                for i in range(args.steps):
                    step_device(next_i[0])
                    next_i[0] += 1
                if copy_stream is not None:
                    stream.wait_stream(copy_stream)          # join
            return g
        next_i = [nwarm]
        warm_graph = capture()
        launches_dev = launches() - l0       # kernels recorded into one graph = launched per replay
        # step counters only move forward: the timed replay is a second graph over the NEXT K steps
        graph = capture() if peer is not None else warm_graph
        with torch.cuda.stream(stream):
            warm_graph.replay()              # one untimed replay
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        if graph is not None:
            graph.replay()
        else:
            for i in range(args.steps):
                step_device(nwarm + i)
        e1.record(stream)
    barrier()
    if graph is None:
        launches_dev = launches() - l0
    total_ms = float(e0.elapsed_time(e1))

    # ---- the same on ONE copy (L2 warm when the working set fits), for comparison ------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            eng.spmm_device(N, ALPHA, dBs[0], ld, BETA, dCins[0], dCouts[0], ld)
        e1.record(stream)
    barrier()
    warm_ms = e0.elapsed_time(e1) / args.steps

    # ---- one isolated cold launch: flush L2 by overwriting 512 MB, then a single step ---
    with torch.cuda.stream(stream):
        flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
        iso = []
        for i in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.spmm_device(N, ALPHA, dBs[0], ld, BETA, dCins[0], dCouts[0], ld)
            b.record(stream)
            iso.append((a, b))
    barrier()
    isolated_ms = float(np.median([a.elapsed_time(b) for a, b in iso]))
    del flush

    # ---- e2e: host-facing call, pinned host buffers -----------------------------------
    hB = sx.pinned_empty(K * N, dtype)
    hC = sx.pinned_empty(M * N, dtype)
    hB[:] = w["B"]
    for _ in range(3):
        hC[:] = w["Cin"]
        eng.spmm(N, ALPHA, hB, BETA, hC, want_ns=not args.host_fused)
    checksum = float(np.asarray(hC, dtype=np.float64).sum())
    barrier()
    l1 = launches()
    e2e_s = 0.0
    for _ in range(args.steps):
        hC[:] = w["Cin"]                       # restore the in/out operand (untimed)
        t0 = time.perf_counter()
        eng.spmm(N, ALPHA, hB, BETA, hC, want_ns=not args.host_fused)   # H2D B, H2D C, kernels, D2H C; returns synchronised
        e2e_s += time.perf_counter() - t0
    launches_e2e = launches() - l1
    barrier()
    clocks = sampler.result()

    # ---- reduce over ranks: max time ---------------------------------------------------
    t = torch.tensor([total_ms, e2e_s, warm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, warm_ms = t.tolist()

    flops_step = 2.0 * nnz * N * world
    value = flops_step * args.steps / (total_ms * 1e-3) / 1e9
    e2e = flops_step * args.steps / e2e_s / 1e9
    kern_ms = total_ms / args.steps       # N=1: the step is exactly one SpMM kernel launch
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    lk = eng.info(sx.INFO_LAST_KERNEL)
    if eng.info(sx.INFO_TILE_NNZ) > 0:
        tiles_note = (f"; dense tiles: {eng.info(sx.INFO_TILE_NNZ)} nnz in {eng.info(sx.INFO_TILE_SLOTS)} slots "
                      f"(fill {eng.info(sx.INFO_TILE_NNZ) / max(1, eng.info(sx.INFO_TILE_SLOTS)):.2f}) on spmm_panels_dmma_kernel, "
                      f"{eng.info(sx.INFO_REST_NNZ)} nnz left to CSR")
    else:
        tiles_note = ""
    kernel_name = ({1: "spmm_rows_kernel (+segments/finalize)", 2: "spmm_staged_kernel", 3: "spmm_window_kernel", 4: "spmm_panels_dmma_kernel", 6: "spmm_window_hostc_kernel", 7: "spmm_slide_kernel", 8: "spmm_edgelist_kernel"}.get(lk // 10000, "?")
                   + f" <G={lk % 10000 // 100}, VPL={lk % 100 // 10}, {'fast' if lk % 10 else 'strict'}>"
                   if lk // 10000 != 5 else f"spmm_staged_kernel<WIN>, {lk % 10000} column-window passes of {cw} columns each") + (
                   (f", {eng.info(sx.INFO_ITEMS)} items of <= {eng.info(sx.INFO_ITEM_NNZ)} nnz, {eng.info(sx.INFO_SPLIT_ROWS)} split rows" if lk // 10000 == 2 else "") + tiles_note)
    host_path = {1: "zero-copy kernels over PCIe (no memcpy)", 2: "zero-copy, C carried by the SpMM kernel (SX_OPT_HOST_FUSED)"}.get(eng.info(sx.INFO_HOST_PATH), "cudaMemcpyAsync + layout kernels")
    if rank == 0:
        line = {
            "metric": "SpMM GFLOP/s (2*nnz*N)", "value": value, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if dtype == np.float64 else "f32",
            "data": "synthetic" if w["name"] in ("uniform", "powerlaw") else "SuiteSparse fixture shipped with the reference, host program's B/C",
            "config": {"workload": w["desc"], "alpha": ALPHA, "beta": BETA, "arith": args.arith,
                       "l2": "warm (--no-flush)" if args.no_flush else (f"inputs larger than L2: {R} independent device copies of A/B/C ({R * alg_bytes / 1e6:.0f} MB > 2 x 126 MB L2), step i uses copy i mod {R}; no flush kernel in the timed region" if R > 1 else f"inputs larger than L2: one copy is {alg_bytes / 1e6:.0f} MB; steps run back to back"),
                       "launch": f"the {args.steps} steps are one CUDA graph replay" if use_graph else "one by one",
                       "partition": "1 row block" if world == 1 else f"{world} stacked row blocks, one per GPU; {exchange}"},
            "gflops_ref_formula": 2.0 * (nnz + M) * N * world * args.steps / (total_ms * 1e-3) / 1e9,
            "single_copy_back_to_back": {"ms_per_step": warm_ms, "value": flops_step / (warm_ms * 1e-3) / 1e9,
                                         "gbs": alg_bytes / (warm_ms * 1e-3) / 1e9, "note": "L2-warm when one copy fits in L2"},
            "isolated_cold_launch": {"ms": isolated_ms, "note": "512 MB overwritten, then ONE step between two events (includes ~5 us event-to-event launch floor)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(w["name"], N, "f64" if s == 8 else "f32"),
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                         "kernel": kernel_name},
            "e2e": {"value": e2e, "unit": "GFLOP/s", "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": (K * N + M * N) * s, "d2h_bytes_per_step": M * N * s,
                    "timer": "host wall clock around the blocking sx_spmm_* call", "path": host_path},
            "gpu_launches": int(launches_dev + launches_e2e),
            "gpu_launches_detail": {"device_steps": int(launches_dev), "e2e_steps": int(launches_e2e)},
            "clocks": clocks, "checksum_C": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w, 1)
        emit(line)
    for e in engines:
        e.close()
    if world > 1:
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
