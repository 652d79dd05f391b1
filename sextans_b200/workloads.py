"""Named inputs of the SpMM path (BASELINE.json configs; SURVEY.md 8(d) "Synthetic inputs").

Host-side input construction only (numpy): the two SuiteSparse matrices that ship with
the reference (matrices/nasa4704, matrices/pcrystk02 -- kept xz-compressed under
tests/golden/matrices) and the two synthetic CSR families.  Generators are seeded and
deterministic, so every rank of a multi-GPU run builds identical data without a file.
"""
from __future__ import annotations

import lzma
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FIXTURES = os.path.join(os.path.dirname(_HERE), "tests", "golden", "matrices")
_CACHE = os.path.join(tempfile.gettempdir(), "sextans_b200_fixtures")

SUITESPARSE = ("nasa4704", "pcrystk02")


def suitesparse_path(name: str) -> str:
    """Unpack tests/golden/matrices/<name>.mtx.xz once per machine and return the .mtx path."""
    packed = os.path.join(_FIXTURES, name + ".mtx.xz")
    if not os.path.exists(packed):
        raise FileNotFoundError(packed)
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


def host_dense(M, K, N, dtype):
    """The host program's operands (src/sextans-host.cpp:100-111): B = 1 and
    C_in[m,n] = float((m+1)(n+1)/M/N), column-major 1-D."""
    B = np.ones(K * N, dtype=dtype)
    m = np.arange(1, M + 1, dtype=np.float64)[None, :]
    n = np.arange(1, N + 1, dtype=np.float64)[:, None]
    C = (1.0 * m * n / M / N).astype(np.float32).astype(dtype)   # [N, M] == column-major
    return B, np.ascontiguousarray(C).ravel()


def random_dense(M, K, N, seed, dtype):
    """uniform(-1,1) B and C_in, column-major 1-D."""
    rng = np.random.default_rng(seed + 1000003)
    B = rng.random(K * N, dtype=np.float32 if np.dtype(dtype) == np.float32 else np.float64)
    C = rng.random(M * N, dtype=B.dtype)
    B *= 2; B -= 1; C *= 2; C -= 1
    return B, C


def uniform_csr(M, K, per_row, seed=12345, dtype=np.float32):
    """Every row has exactly ``per_row`` nonzeros, columns distinct, ascending, uniform in
    [0, K); values uniform(-1, 1).  (C4: M = K = 1e6, per_row = 20.)"""
    rng = np.random.default_rng(seed)
    # distinct ascending columns per row without a per-row loop: sorted uniform draws from
    # [0, K - per_row] plus 0..per_row-1 are strictly increasing and stay below K
    cols = rng.integers(0, K - per_row + 1, size=(M, per_row), dtype=np.int32)
    cols.sort(axis=1)
    cols += np.arange(per_row, dtype=np.int32)[None, :]
    rowptr = (np.arange(M + 1, dtype=np.int64) * per_row).astype(np.int32)
    val = rng.random(M * per_row, dtype=np.float32).astype(dtype)
    val *= 2; val -= 1
    return rowptr, cols.ravel(), val


def powerlaw_csr(M, K, nnz, seed=12345, dtype=np.float64, exponent=2.0, col_skew=2.0):
    """Row lengths ~ Pareto(exponent) (min 1, cap K/4) rescaled towards ``nnz`` in total;
    columns follow a power law over a fixed random permutation of [0, K) (hub columns),
    de-duplicated and ascending inside a row; values uniform(-1, 1).  The returned nnz is
    what is left after de-duplication (a few per cent below the request).
    (C5: M = K = 1e6, nnz = 1e8.)"""
    rng = np.random.default_rng(seed)
    raw = (1.0 - rng.random(M)) ** (-1.0 / (exponent - 1.0))        # Pareto, x_min = 1
    cap = max(1, K // 4)
    lens = np.minimum(raw, cap)
    for _ in range(6):                                               # rescale under the cap
        lens = np.minimum(np.maximum(1.0, lens * (nnz / lens.sum())), cap)
    lens = np.floor(lens).astype(np.int64)
    short = nnz - int(lens.sum())
    if short > 0:                                                    # spread the remainder
        idx = rng.integers(0, M, size=short)
        np.add.at(lens, idx, 1)
    lens = np.minimum(lens, cap)
    total = int(lens.sum())
    rows = np.repeat(np.arange(M, dtype=np.int64), lens)
    u = rng.random(total)
    perm = rng.permutation(K).astype(np.int64)
    cols = perm[np.minimum((K * u ** col_skew).astype(np.int64), K - 1)]
    key = rows * K + cols
    del rows, cols, u
    key.sort()
    keep = np.empty(total, dtype=bool)
    keep[0:1] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    key = key[keep]
    rows = key // K
    colidx = (key - rows * K).astype(np.int32)
    rowptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=rowptr[1:])
    val = rng.random(colidx.size).astype(dtype)
    val *= 2; val -= 1
    return rowptr.astype(np.int32), colidx, val


def fem_like_csr(nodes, dof, nbrs, seed=12345, dtype=np.float64, band=2000, noise_per_row=0):
    """Block-structured CSR of the kind finite-element codes produce (what the shipped
    nasa4704 / pcrystk02 are): `nodes` mesh nodes with `dof` unknowns each; a node is
    coupled to itself and to ~`nbrs` other nodes within `band` node numbers, and every
    coupling is a dense dof x dof block.  M = K = nodes*dof, nnz ~ nodes*(1+nbrs)*dof^2.
    `noise_per_row` extra uniformly random entries per row break the pure block structure.
    Rows keep ascending, duplicate-free columns."""
    rng = np.random.default_rng(seed)
    off = rng.integers(-band, band + 1, size=(nodes, nbrs), dtype=np.int64)
    nb = np.clip(np.arange(nodes, dtype=np.int64)[:, None] + off, 0, nodes - 1)
    nb = np.concatenate([nb, np.arange(nodes, dtype=np.int64)[:, None]], axis=1)
    nb.sort(axis=1)
    keep = np.ones_like(nb, dtype=bool)
    keep[:, 1:] = nb[:, 1:] != nb[:, :-1]
    lens_n = keep.sum(axis=1)                         # couplings per node
    bcol = nb[keep]                                   # block columns, node-major, ascending
    bptr = np.zeros(nodes + 1, dtype=np.int64)
    np.cumsum(lens_n, out=bptr[1:])
    # element columns of one node's rows: every block column expanded to dof columns
    pat = (bcol[:, None] * dof + np.arange(dof, dtype=np.int64)[None, :]).ravel()
    L = lens_n * dof                                  # row length of each of the node's dof rows
    row_node = np.repeat(np.arange(nodes, dtype=np.int64), dof)
    row_len = L[row_node]
    rowptr = np.zeros(nodes * dof + 1, dtype=np.int64)
    np.cumsum(row_len, out=rowptr[1:])
    total = int(rowptr[-1])
    start = (bptr[:-1] * dof)[row_node]               # where the node's pattern starts in `pat`
    idx = np.repeat(start - rowptr[:-1], row_len) + np.arange(total, dtype=np.int64)
    colidx = pat[idx].astype(np.int32)
    del idx, pat
    M = K = nodes * dof
    if noise_per_row > 0:
        extra = rng.integers(0, K, size=(M, noise_per_row), dtype=np.int64)
        rows = np.concatenate([np.repeat(np.arange(M, dtype=np.int64), row_len),
                               np.repeat(np.arange(M, dtype=np.int64), noise_per_row)])
        cols = np.concatenate([colidx.astype(np.int64), extra.ravel()])
        key = np.unique(rows * K + cols)
        rows = key // K
        colidx = (key - rows * K).astype(np.int32)
        rowptr = np.zeros(M + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=M), out=rowptr[1:])
    val = rng.random(colidx.size).astype(dtype)
    val *= 2; val -= 1
    return rowptr.astype(np.int32), colidx, val


def algorithmic_bytes(M, K, nnz, N, itemsize, beta_nonzero=True):
    """CSR A once (32-bit indices), B once, C read once and written once."""
    c_passes = 2 if beta_nonzero else 1
    return nnz * (4 + itemsize) + 4 * (M + 1) + K * N * itemsize + c_passes * M * N * itemsize
