"""The block-level kernels of spmm_kernels.cuh, executed on the CPU.

tests/emu/ holds a small emulation of the CUDA block execution model (one OS thread per CUDA
thread, real __syncthreads, TMA bulk copies as memcpy completing on an emulated mbarrier).
This test builds the PRODUCT's kernel source against it -- textually the same file, with
only the PTX helper section (cache-policy loads/stores, mbarrier, cp.async.bulk) replaced by
tests/emu/emu_helpers.h and the dynamic shared-memory declaration pointed at the emulator's
buffer -- and runs variant 3 (32/64/128-row blocks, with and without the PDL code path), the
host-boundary fusion kernel, and variant 2 (the TMA-staged lane-group kernel with its finalize
kernel: every lane-group shape, split rows, the prefetch code path, column-window passes)
and variant 4 (the sliding-window kernel, planned by the product's own sx_plan_slide)
against the cpu_spmm_CSR loop, bit for bit.  It is how the
kernels that were written after the round's GPU time was spent had their index arithmetic,
staging and barrier structure checked; it does not replace the GPU parity tests."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
# the host-side planners (sx_plan_slide) come from the product library; it loads without a GPU
LIBDIR = os.path.join(ROOT, "sextans_b200")
LINK = [f"-L{LIBDIR}", "-lsextans_b200", f"-Wl,-rpath,{LIBDIR}"]
# SX_EMU_DEFS="-DSX_STAGED_UMAX=16" checks an alternative build-time configuration of the kernels
LINK += os.environ.get("SX_EMU_DEFS", "").split()


def emulated_header():
    src = open(os.path.join(ROOT, "sextans_b200", "csrc", "spmm_kernels.cuh")).read()
    a = src.index("// ---- cache-policy loads/stores")
    b = src.index("// ---- main kernel, staged")
    assert 0 < a < b
    src = src[:a] + '#include "emu_helpers.h"\n\n' + src[b:]
    decl = "extern __shared__ __align__(128) unsigned char smem_raw[];"
    assert src.count(decl) >= 3
    src = src.replace(decl, "unsigned char *smem_raw = sx_emu::dyn_smem();")
    src = src.replace("uint64_t pol_b;", "uint64_t pol_b = 0;")
    assert "asm volatile" in src          # what is left is fences / policies / griddepcontrol: no-ops here
    return src


@pytest.mark.skipif(CXX is None, reason="no host C++ compiler")
def test_block_level_kernels_on_the_cpu_emulation(tmp_path):
    (tmp_path / "spmm_kernels_emu.cuh").write_text(emulated_header())
    exe = tmp_path / "emu_kernels"
    cmd = [CXX, "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-w", f"-I{tmp_path}", f"-I{EMU}",
           f"-I{os.path.join(EMU, 'include')}", os.path.join(EMU, "emu_kernels.cpp"), "-o", str(exe)] + LINK
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "EMULATION: all bit-exact" in r.stdout and "MISMATCH" not in r.stdout
    ran = re.findall(r"^(window RB=\d+(?: PDL)?) +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    kinds = {k.strip() for k in ran}
    assert {"window RB=32", "window RB=32 PDL"} <= kinds
    staged = re.findall(r"^(staged.*?) +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    assert {"staged", "staged + prefetch path", "staged as column-window passes"} <= {k.strip() for k in staged}
    assert len(staged) >= 60
    rows = re.findall(r"^rows \+ segments \(variant 1\) +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    assert len(rows) >= 14
    slide = re.findall(r"^slide \(variant 4\) +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    assert len(slide) >= 14
    edge = re.findall(r"^edge lists \(variant 5\).*? +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    assert len(edge) >= 18 and "PLAN INVARIANT MISMATCH" not in r.stdout
    hostc = re.findall(r"^edge lists HOSTC .*? +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    assert len(hostc) >= 18
    host1 = re.findall(r"^edge lists HOST1 .*? +f(?:32|64) M=.*bit-exact$", r.stdout, flags=re.M)
    assert len(host1) >= 9


@pytest.mark.skipif(CXX is None or os.environ.get("SX_EMU_ASAN") != "1",
                    reason="address-sanitizer pass of the emulation: set SX_EMU_ASAN=1 (about two minutes)")
def test_emulated_kernels_stay_inside_their_buffers(tmp_path):
    """The same run under AddressSanitizer.  Device buffers are allocated at exactly the sizes
    (and pads) the product allocates, dynamic shared memory at exactly the launch's size, so an
    out-of-bounds access of a kernel is reported instead of going unnoticed."""
    (tmp_path / "spmm_kernels_emu.cuh").write_text(emulated_header())
    exe = tmp_path / "emu_kernels_asan"
    cmd = [CXX, "-std=c++20", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-ffp-contract=off",
           "-pthread", "-w", f"-I{tmp_path}", f"-I{EMU}", f"-I{os.path.join(EMU, 'include')}",
           os.path.join(EMU, "emu_kernels.cpp"), "-o", str(exe)] + LINK
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, r.stderr[-4000:]
    assert "EMULATION: all bit-exact" in r.stdout


@pytest.mark.skipif(CXX is None or os.environ.get("SX_EMU_TSAN") != "1",
                    reason="thread-sanitizer pass of the emulation: set SX_EMU_TSAN=1 (about two minutes)")
def test_emulated_kernels_have_no_shared_memory_races(tmp_path):
    """The same run under ThreadSanitizer: the emulation's CUDA threads are OS threads, its
    barriers / mbarriers are real synchronisation, so a missing __syncthreads, a TMA copy into
    shared memory that other threads still read, or two threads writing one location shows up
    as a data race.  (Removing one barrier from the host-boundary fusion kernel produces 64
    reports.)  A CPU-side stand-in for `compute-sanitizer --tool racecheck`."""
    (tmp_path / "spmm_kernels_emu.cuh").write_text(emulated_header())
    exe = tmp_path / "emu_kernels_tsan"
    cmd = [CXX, "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-pthread", "-w",
           f"-I{tmp_path}", f"-I{EMU}", f"-I{os.path.join(EMU, 'include')}", os.path.join(EMU, "emu_kernels.cpp"),
           "-o", str(exe)] + LINK
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1800)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and "EMULATION: all bit-exact" in r.stdout

