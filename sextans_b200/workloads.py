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


def _powerlaw_keys(rng, M, K, nnz, exponent, col_skew, perm):
    """Sorted unique keys row*K + col of a power-law matrix with about ``nnz`` entries before
    de-duplication: row lengths ~ Pareto(exponent) (min 1, cap K/4) rescaled to ``nnz``, columns
    ~ u**col_skew over the permutation (hub columns)."""
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
    cols = perm[np.minimum((K * u ** col_skew).astype(np.int64), K - 1)]
    key = rows * K + cols
    del rows, cols, u
    key.sort()
    keep = np.empty(total, dtype=bool)
    keep[0:1] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    return key[keep]


def _trim_keys(rng, key, K, target, protect=None):
    """Drop random entries until exactly ``target`` are left; never a row's first entry (so no
    row becomes empty) and never an entry flagged in ``protect``."""
    surplus = key.size - target
    if surplus <= 0:
        return key
    free = np.empty(key.size, dtype=bool)
    free[0:1] = False
    np.not_equal(key[1:] // K, key[:-1] // K, out=free[1:])
    np.logical_not(free, out=free)                                   # free = not the first entry of its row
    if protect is not None:
        free &= ~protect
    cand = np.flatnonzero(free)
    if cand.size < surplus:
        raise ValueError("cannot trim to the requested number of nonzeros")
    drop = rng.choice(cand, size=surplus, replace=False)
    keep = np.ones(key.size, dtype=bool)
    keep[drop] = False
    return key[keep]


def _keys_to_csr(rng, key, M, K, dtype):
    rows = key // K
    colidx = (key - rows * K).astype(np.int32)
    rowptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=rowptr[1:])
    val = rng.random(colidx.size).astype(dtype)
    val *= 2; val -= 1
    return rowptr.astype(np.int32), colidx, val


def powerlaw_csr(M, K, nnz, seed=12345, dtype=np.float64, exponent=2.0, col_skew=2.0, oversample=1.06):
    """EXACTLY ``nnz`` nonzeros ("rescaled to hit nnz exactly"): row lengths ~
    Pareto(exponent) (min 1, cap K/4); columns follow a power law over a fixed random
    permutation of [0, K) (hub columns), de-duplicated and ascending inside a row; values
    uniform(-1, 1).  The draw asks for ``oversample`` x nnz entries (de-duplication of the hub
    columns removes ~4 % at C5's size), more if that is not enough, and random surplus entries
    are then dropped.  (C5: M = K = 1e6, nnz = 1e8.)"""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(K).astype(np.int64)
    nnz = int(min(nnz, M * (K // 4) if K >= 4 else M))
    ask = oversample
    key = _powerlaw_keys(rng, M, K, int(nnz * ask), exponent, col_skew, perm)
    while key.size < nnz:                                            # very dense requests only
        ask *= 1.25
        key = _powerlaw_keys(rng, M, K, int(nnz * ask), exponent, col_skew, perm)
    key = _trim_keys(rng, key, K, nnz)
    return _keys_to_csr(rng, key, M, K, dtype)


def powerlaw_blocked_csr(M, K, nnz, seed=12345, dtype=np.float64, block=16, block_frac=0.25,
                         exponent=2.0, col_skew=2.0):
    """C5's blocked sub-case: the power-law matrix with dense ``block`` x
    ``block`` sub-blocks planted so that they hold ``block_frac`` of the EXACTLY ``nnz``
    nonzeros -- what a blocked-ELL / dense-tile variant can find.  Blocks start at rows that are
    multiples of ``block`` and at arbitrary columns (a block covers ``block`` consecutive
    columns); the rest is powerlaw_csr's pattern.  Returns (rowptr, colidx, val, planted_nnz)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(K).astype(np.int64)
    nblocks = int(nnz * block_frac) // (block * block)
    br = rng.integers(0, max(1, M // block), size=nblocks, dtype=np.int64) * block
    bc = rng.integers(0, max(1, K - block + 1), size=nblocks, dtype=np.int64)
    rr = (br[:, None, None] + np.arange(block, dtype=np.int64)[None, :, None])
    cc = (bc[:, None, None] + np.arange(block, dtype=np.int64)[None, None, :])
    planted = np.unique((rr * K + cc).ravel())
    del rr, cc
    rest = nnz - planted.size
    ask = 1.10
    while True:
        key = _powerlaw_keys(rng, M, K, int(rest * ask), exponent, col_skew, perm)
        key = np.concatenate([key, planted])
        key.sort()
        keep = np.empty(key.size, dtype=bool)
        keep[0:1] = True
        np.not_equal(key[1:], key[:-1], out=keep[1:])
        key = key[keep]
        if key.size >= nnz:
            break
        ask *= 1.25
    at = np.minimum(np.searchsorted(planted, key), planted.size - 1)
    protect = planted[at] == key
    key = _trim_keys(rng, key, K, nnz, protect=protect)
    rp, ci, v = _keys_to_csr(rng, key, M, K, dtype)
    return rp, ci, v, int(planted.size)


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
