"""CPU tests of the product's host side: the C-ABI library loads here (no GPU),
exports every symbol include/sextans_b200.h declares, fails loudly without a device,
and its host-only entry points (loader, partitioner) agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import SMALL_MTX, SUITESPARSE, mtx_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sextans_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sx_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 30 and "sx_spmm_f64" in names and "sx_load_mtx_f32" in names
    L = ctypes.CDLL(sx.library_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sx.lib().sx_abi_version() == 1


def test_python_mirror_binds_every_declared_symbol():
    L = sx.lib()
    for name in declared_symbols():
        assert getattr(L, name).argtypes is not None, name


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sx.SextansError) as e:
        sx.Engine(0)
    assert "NO_DEVICE" in str(e.value)


def test_product_does_not_touch_the_oracle():
    # the product path must not import, link, include or execute anything under oracle/
    banned = re.compile(r"import\s+oracle|from\s+oracle|liboracle|sextans_ref|oracle/|\boracle\.")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sextans_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not banned.search(text), (dirpath, f, banned.search(text).group(0))
    out = os.popen(f"ldd {sx.library_path()}").read()
    assert "oracle" not in out and "sextans_ref" not in out


@pytest.mark.parametrize("name", SMALL_MTX + SUITESPARSE)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_loader_equals_oracle_loader(name, dtype):
    mine = sx.load_mtx(mtx_path(name), dtype)
    ref = oracle.load_mtx(mtx_path(name), dtype)
    assert mine[:3] == ref[:3]
    for a, b in zip(mine[3:], ref[3:6]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert np.array_equal(mine[5].view(np.uint8), ref[5].view(np.uint8))  # -0.0 survives


def test_loader_free_form_whitespace(tmp_path):
    # the reference reads entries with fscanf, i.e. token by token: entries may share
    # or straddle lines
    p = tmp_path / "ws.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% c\n3 3 4\n1 1 1.0 2 2\n2.0\n\n3 1   -4e0\t3 3 5\n")
    M, K, nnz, rp, ci, v = sx.load_mtx(str(p))
    o = oracle.load_mtx(str(p))
    assert (M, K, nnz) == o[:3] == (3, 3, 4)
    assert np.array_equal(rp, o[3]) and np.array_equal(ci, o[4]) and np.array_equal(v, o[5])


def test_loader_errors(tmp_path):
    def write(name, text):
        p = tmp_path / name
        p.write_text(text)
        return str(p)
    with pytest.raises(sx.SextansError, match="IO"):
        sx.load_mtx(str(tmp_path / "missing.mtx"))
    with pytest.raises(sx.SextansError, match="FORMAT"):
        sx.load_mtx(write("a.mtx", "%%NotMM matrix coordinate real general\n1 1 1\n1 1 1\n"))
    with pytest.raises(sx.SextansError, match="coordinate"):
        sx.load_mtx(write("b.mtx", "%%MatrixMarket matrix array real general\n1 1 1\n1.0\n"))
    with pytest.raises(sx.SextansError, match="complex"):
        sx.load_mtx(write("c.mtx", "%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 2\n"))
    with pytest.raises(sx.SextansError, match="below 1"):
        sx.load_mtx(write("d.mtx", "%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 3.0\n"))
    # the reference would write out of bounds / reuse stale values here; we refuse
    with pytest.raises(sx.SextansError, match="beyond"):
        sx.load_mtx(write("e.mtx", "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 3.0\n"))
    with pytest.raises(sx.SextansError, match="missing"):
        sx.load_mtx(write("f.mtx", "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 3.0\n"))


def test_partition_rows():
    rp = sx.load_mtx(mtx_path("nasa4704"))[3]
    for parts in (1, 2, 3, 4, 8, 64):
        b = sx.partition_rows(rp, parts)
        assert b[0] == 0 and b[-1] == rp.size - 1 and np.all(np.diff(b) >= 0)
        share = np.diff(rp[b])
        assert share.sum() == rp[-1]
        assert share.max() - share.min() <= 2 * np.diff(rp).max()
    # skewed: one heavy row
    rp = np.array([0, 1, 2, 1000, 1001, 1002], dtype=np.int32)
    assert sx.partition_rows(rp, 2).tolist() == [0, 3, 5] or sx.partition_rows(rp, 2).tolist() == [0, 2, 5]
    # more parts than rows, and an empty matrix
    assert sx.partition_rows(np.array([0, 5], dtype=np.int32), 4)[-1] == 1
    assert sx.partition_rows(np.zeros(5, dtype=np.int32), 2).tolist() == [0, 2, 4]


def _write_big_mtx(path, M, K, nz, seed, field="real", symmetry="general", irregular=False):
    rng = np.random.default_rng(seed)
    r = rng.integers(1, M + 1, size=nz)
    c = rng.integers(1, K + 1, size=nz)
    if symmetry == "symmetric":
        r, c = np.maximum(r, c), np.minimum(r, c)
    v = rng.uniform(-1, 1, size=nz)
    v[::97] = 0.0            # explicit +0 entries are dropped
    v[1::97] = -0.0          # -0 is kept
    r[5::211], c[5::211] = r[4::211][:r[5::211].size], c[4::211][:c[5::211].size]   # duplicates
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {symmetry}\n% generated\n{M} {K} {nz}\n")
        if field == "pattern":
            lines = [f"{a} {b}" for a, b in zip(r, c)]
        else:
            lines = [f"{a} {b} {float(x)!r}" for a, b, x in zip(r, c, v)]
        if irregular:         # two entries on one line here and there: legal for fscanf
            lines[10] = lines[10] + " " + lines.pop(11)
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("field,symmetry", [("real", "general"), ("real", "symmetric"), ("pattern", "symmetric")])
def test_parallel_loader_equals_oracle_for_any_thread_count(tmp_path, monkeypatch, field, symmetry):
    p = str(tmp_path / "big.mtx")
    K = 3000 if symmetry == "symmetric" else 2500                      # symmetric files are square
    _write_big_mtx(p, 3000, K, 160_000 if field == "pattern" else 120_000, 5, field, symmetry)   # > 1 MB: the parallel route
    assert os.path.getsize(p) > (1 << 20)
    ref = oracle.load_mtx(p, np.float64)
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("SX_LOADER_THREADS", threads)
        mine = sx.load_mtx(p, np.float64)
        assert mine[:3] == ref[:3]
        for a, b in zip(mine[3:], ref[3:6]):
            assert np.array_equal(a, b)
        assert np.array_equal(mine[5].view(np.uint64), ref[5].view(np.uint64))
    mine32, ref32 = sx.load_mtx(p, np.float32), oracle.load_mtx(p, np.float32)
    assert all(np.array_equal(a, b) for a, b in zip(mine32[3:], ref32[3:6]))


def test_parallel_loader_falls_back_on_irregular_lines(tmp_path):
    p = str(tmp_path / "irr.mtx")
    _write_big_mtx(p, 3000, 2500, 120_000, 6, irregular=True)
    mine, ref = sx.load_mtx(p, np.float64), oracle.load_mtx(p, np.float64)
    assert mine[:3] == ref[:3] and all(np.array_equal(a, b) for a, b in zip(mine[3:], ref[3:6]))
    # fewer entries than declared: still an error, whatever the route
    with open(p) as f:
        text = f.read().splitlines()
    with open(p, "w") as f:
        f.write("\n".join(text[:-50]) + "\n")
    with pytest.raises(sx.SextansError, match="missing"):
        sx.load_mtx(p, np.float64)


def test_host_program_call_surface_without_gpu():
    """The argv contract of src/sextans-host.cpp:33-48 needs no device; without a GPU the
    program must stop with an error after loading A -- never compute on the CPU instead."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "sextans_b200", "sextans")
    assert os.path.exists(exe), "build with make -C sextans_b200/csrc"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and r.stdout.startswith("start host\n")
    assert "Usage: " in r.stdout and "[matrix A file] [N] [rp_time] [alpha] [beta]" in r.stdout
    r = subprocess.run([exe, "/nonexistent.mtx", "8"], capture_output=True, text=True)
    assert r.returncode == 1 and "N = 8" in r.stdout and "Could not open" in r.stdout
    if not torch.cuda.is_available():
        r = subprocess.run([exe, mtx_path("nasa4704"), "13"], capture_output=True, text=True)
        assert r.returncode != 0
        assert "N = 16" in r.stdout and "A: sparse matrix, 4704 x 4704. NNZ = 104756" in r.stdout
        assert "NO_DEVICE" in r.stderr and "Success!" not in r.stdout


def test_python_constants_match_the_header_enums():
    text = open(os.path.join(ROOT, "include", "sextans_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    pairs = re.findall(r"\b(SX_(?:OPT|INFO)_[A-Z_]+)\s*=\s*(\d+)", text)
    assert len(pairs) >= 25
    for name, value in pairs:
        assert getattr(sx, name[3:]) == int(value), name
