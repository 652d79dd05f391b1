"""N>1 on real GPUs (needs >= 2 devices; skipped on a 1-GPU box): row blocks over NCCL
through the Python host mirror, and the `sextans --gpus 2` host program."""
import os
import subprocess
import sys

import pytest

from helpers import mtx_path

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("nproc", [2, 4])
def test_row_blocks_over_nccl_match_oracle_bitwise(nproc):
    """2 ranks: holder -> receiver.  4 ranks: the push tree has a rank that receives AND forwards
    (profiles/r02_multi_gpu_worker_4gpu.txt is the kept output of the 4-GPU run)."""
    if _gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29515 + nproc), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=560)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK ") == 7, r.stdout


@pytest.mark.timeout(300)
def test_sextans_program_two_gpus():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "sextans_b200", "sextans")
    r = subprocess.run([exe, mtx_path("nasa4704"), "16", "10", "--gpus", "2", "--dtype", "f64", "--json"],
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Success!" in r.stdout and "num_mismatch = 0" in r.stdout and "max_rel_err = 0.000e+00" in r.stdout
    assert "B broadcast over NCCL" in r.stdout


@pytest.mark.timeout(300)
def test_sextans_program_reference_call_surface():
    """The positional forms of src/sextans-host.cpp:33-48 and its stdout lines."""
    exe = os.path.join(ROOT, "sextans_b200", "sextans")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage:" in r.stdout and "[matrix A file] [N] [rp_time] [alpha] [beta]" in r.stdout
    r = subprocess.run([exe, mtx_path("nasa4704"), "13", "3", "1.5", "0.25"], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    for line in ("start host", "N = 16", "alpha = 1.5", "beta = 0.25", "A: sparse matrix, 4704 x 4704. NNZ = 104756",
                 "B: dense matrix, 4704 x 16", "launch kernel", "Kernel time is", "GFLOPS:", "Success!",
                 "num_mismatch = 0, percent = 0.00%"):
        assert line in out, (line, out)
    r = subprocess.run([exe, "/nonexistent.mtx", "8"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open" in r.stdout
