"""Shared helpers for the test-suite (fixtures, seeded inputs, error metrics)."""
from __future__ import annotations

import hashlib
import lzma
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
SMALL_MTX = ["general_real", "symmetric_real", "integer_general", "skew",
             "pattern_general", "pattern_symmetric"]
SUITESPARSE = ["nasa4704", "pcrystk02"]

_CACHE = os.path.join(tempfile.gettempdir(), "sextans_b200_fixtures")


def mtx_path(name: str) -> str:
    """Path of a fixture .mtx.  The two SuiteSparse matrices are stored
    xz-compressed (tests/golden/matrices) and unpacked once per machine."""
    small = os.path.join(GOLDEN, "mtx", name + ".mtx")
    if os.path.exists(small):
        return small
    packed = os.path.join(GOLDEN, "matrices", name + ".mtx.xz")
    if not os.path.exists(packed):
        raise FileNotFoundError(name)
    os.makedirs(_CACHE, exist_ok=True)
    out = os.path.join(_CACHE, name + ".mtx")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(packed):
        with lzma.open(packed, "rb") as f:
            data = f.read()
        tmp = out + f".{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, out)
    return out


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def perturbed_inputs(M, K, N, nnz, dtype):
    """Value-sensitive inputs for the pattern matrices:
    val[j] = 1 + 0.001*(j mod 97); B[k,n] = (1 + k mod 13) + 0.1*(1+n) (the
    alternative the reference host has commented out, sextans-host.cpp:102);
    C_in[m,n] = ((m*7 + n*3) mod 19 - 9) / 8."""
    j = np.arange(nnz, dtype=np.int64)
    val = (1.0 + 0.001 * (j % 97)).astype(dtype)
    k = np.arange(K, dtype=np.int64)[:, None]
    n = np.arange(N, dtype=np.int64)[None, :]
    B = ((1.0 + k % 13) + 0.1 * (1.0 + n)).astype(dtype)          # [K, N]
    m = np.arange(M, dtype=np.int64)[:, None]
    Cin = ((((m * 7 + n * 3) % 19) - 9) / 8.0).astype(dtype)      # [M, N]
    # column-major 1-D, as the host driver stores dense operands
    return val, np.ascontiguousarray(B.T).ravel(), np.ascontiguousarray(Cin.T).ravel()


def random_csr(M, K, avg, seed, dtype, empty_frac=0.1, long_row=None, sort=True):
    """Seeded random CSR with some empty rows, optional one very long row,
    duplicate-free ascending columns per row."""
    rng = np.random.default_rng(seed)
    lens = rng.poisson(avg, size=M).astype(np.int64)
    lens[rng.random(M) < empty_frac] = 0
    lens = np.minimum(lens, K)
    if long_row is not None and M > 0:
        lens[rng.integers(0, M)] = min(long_row, K)
    rowptr = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(lens, out=rowptr[1:])
    cols = []
    for r in range(M):
        c = rng.choice(K, size=int(lens[r]), replace=False)
        cols.append(np.sort(c) if sort else c)
    colidx = (np.concatenate(cols) if cols else np.zeros(0)).astype(np.int32)
    val = rng.uniform(-1.0, 1.0, size=colidx.size).astype(dtype)
    return rowptr, colidx, val


def random_dense(M, K, N, seed, dtype):
    rng = np.random.default_rng(seed + 1000003)
    B = rng.uniform(-1.0, 1.0, size=K * N).astype(dtype)
    Cin = rng.uniform(-1.0, 1.0, size=M * N).astype(dtype)
    return B, Cin


def max_rel_err(x, y):
    """|x-y| / max(|y|, 1e-30), max over elements (SURVEY.md 8(c) definition)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.size == 0:
        return 0.0
    return float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-30)))


def scaled_err(x, y):
    """max |x-y| / max|y| -- insensitive to cancellation in individual entries."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.size == 0:
        return 0.0
    return float(np.max(np.abs(x - y)) / max(float(np.max(np.abs(y))), 1e-30))
