"""sextans_b200 -- B200-native SpMM engine behind the Sextans host call surface.

``C = alpha * A @ B + beta * C`` for a CSR / Matrix-Market sparse ``A`` and a tall
dense ``B``.  The product is ``libsextans_b200.so`` (hand-written sm_100a CUDA kernels
behind the C ABI of ``include/sextans_b200.h``) and the ``sextans`` host program;
this package is the thin Python mirror of that ABI used by the tests and the bench.

There is no CPU path here: importing works anywhere (so that the ABI can be
inspected), but creating an :class:`Engine` without the built library or without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["Engine", "SextansError", "lib", "library_path", "load_mtx", "partition_rows",
           "split_col_windows", "plan_slide", "plan_edge_lists",
           "pinned_empty", "STRICT", "FAST", "images_decode_A", "images_decode_B",
           "images_decode_C", "images_encode_C"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SX_F32, SX_F64 = 0, 1
STRICT, FAST = 0, 1
(OPT_ARITH, OPT_SPLIT_ROW_NNZ, OPT_KERNEL, OPT_ITEM_NNZ, OPT_ZEROCOPY_BYTES, OPT_TILE_MIN_ROWS,
 OPT_COL_WINDOW_ROWS, OPT_PREFETCH, OPT_HOST_FUSED, OPT_PDL, OPT_WINDOW_ROWS, OPT_SLIDE,
 OPT_AUTOTUNE, OPT_PANEL_COLS, OPT_HOST_GROUPS) = range(15)
(INFO_LAUNCHES, INFO_M, INFO_K, INFO_NNZ, INFO_DTYPE, INFO_SPLIT_ROWS, INFO_LAST_KERNEL, INFO_LD,
 INFO_ITEMS, INFO_ITEM_NNZ, INFO_HOST_PATH, INFO_TILE_NNZ, INFO_TILE_SLOTS, INFO_REST_NNZ,
 INFO_UPLOAD_SERIAL, INFO_COL_WINDOWS, INFO_TUNED_KERNEL, INFO_EXCHANGE_TIMEOUTS, INFO_EDGE_BLOCKS,
 INFO_EDGE_COLS, INFO_PUSH_PENDING) = range(21)

_PI32 = C.POINTER(C.c_int32)
_PF = C.POINTER(C.c_float)
_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)
_A8 = _PU64 * 8   # const uint64_t *const edge_list_ch[8]
_F4 = _PF * 4     # const float *const mat_B_ch[4]
_F8 = _PF * 8     # (const) float *const mat_C_ch[8]


class SextansError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"{text} [{status}]")
        self.status = status


def library_path() -> str:
    """The C-ABI library.  SX_LIBRARY_PATH selects an alternative build of it (a tuning
    variant made by scripts/build_variant.sh); it is still this product's CUDA library."""
    return os.environ.get("SX_LIBRARY_PATH") or os.path.join(_HERE, "libsextans_b200.so")


def lib():
    """The C ABI.  Raises if the CUDA library has not been built -- never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C sextans_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    vp, i, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    sig = {
        "sx_abi_version": ([], i),
        "sx_last_error": ([], C.c_char_p),
        "sx_status_name": ([i], C.c_char_p),
        "sx_device_count": ([C.POINTER(i)], i),
        "sx_create": ([i, C.POINTER(vp)], i),
        "sx_destroy": ([vp], i),
        "sx_set_stream": ([vp, vp], i),
        "sx_set_option": ([vp, i, i64], i),
        "sx_get_info": ([vp, i, C.POINTER(i64)], i),
        "sx_synchronize": ([vp], i),
        "sx_upload_csr_f32": ([vp, i, i, i64, _PI32, _PI32, _PF], i),
        "sx_upload_csr_f64": ([vp, i, i, i64, _PI32, _PI32, _PD], i),
        "sx_spmm_f32": ([vp, i, C.c_float, vp, C.c_float, vp, i, _PD], i),
        "sx_spmm_f64": ([vp, i, C.c_double, vp, C.c_double, vp, i, _PD], i),
        "sx_stage_B_f32": ([vp, i, vp], i), "sx_stage_B_f64": ([vp, i, vp], i),
        "sx_stage_C_f32": ([vp, i, vp], i), "sx_stage_C_f64": ([vp, i, vp], i),
        "sx_spmm_enqueue_f32": ([vp, i, C.c_float, vp, C.c_float, vp], i),
        "sx_spmm_enqueue_f64": ([vp, i, C.c_double, vp, C.c_double, vp], i),
        "sx_spmm_staged_B_f32": ([vp, i, C.c_float, C.c_float, vp], i),
        "sx_spmm_staged_B_f64": ([vp, i, C.c_double, C.c_double, vp], i),
        "sx_launch_f32": ([vp, C.c_float, C.c_float, i, _PD], i),
        "sx_launch_f64": ([vp, C.c_double, C.c_double, i, _PD], i),
        "sx_fetch_C_f32": ([vp, vp], i), "sx_fetch_C_f64": ([vp, vp], i),
        "sx_device_B": ([vp, i, C.POINTER(vp), C.POINTER(sz)], i),
        "sx_spmm_device_f32": ([vp, i, C.c_float, vp, i64, C.c_float, vp, vp, i64], i),
        "sx_spmm_device_f64": ([vp, i, C.c_double, vp, i64, C.c_double, vp, vp, i64], i),
        "sx_spmm_device_batch_f32": ([vp, i, i, C.c_float, vp, i64, i64, C.c_float, vp, vp, i64, i64], i),
        "sx_spmm_device_batch_f64": ([vp, i, i, C.c_double, vp, i64, i64, C.c_double, vp, vp, i64, i64], i),
        "sx_colmajor_to_rowmajor": ([vp, i, i64, i, vp, vp, i64], i),
        "sx_rowmajor_to_colmajor": ([vp, i, i64, i, vp, i64, vp], i),
        "sx_device_alloc": ([vp, sz, C.POINTER(vp)], i),
        "sx_device_free": ([vp, vp], i),
        "sx_ipc_export": ([vp, vp, C.c_char_p], i),
        "sx_ipc_offset": ([vp, vp, C.POINTER(sz)], i),
        "sx_ipc_import": ([vp, C.c_char_p, C.POINTER(vp)], i),
        "sx_ipc_close": ([vp, vp], i),
        "sx_push_B": ([vp, vp, sz, C.POINTER(vp), C.POINTER(vp), i, vp, vp], i),
        "sx_spmm_expect_push": ([vp, vp, vp, vp], i),
        "sx_spmm_fuse_push": ([vp, C.POINTER(vp), C.POINTER(vp), i, vp, vp], i),
        "sx_spmm_fuse_push_deferred": ([vp, C.POINTER(vp), C.POINTER(vp), i, vp, vp], i),
        "sx_spmm_fuse_publish": ([vp, C.POINTER(vp), i, vp], i),
        "sx_push_publish": ([vp, C.POINTER(vp), i, vp], i),
        "sx_host_alloc": ([sz, C.POINTER(vp)], i),
        "sx_host_free": ([vp], i),
        "sx_partition_rows": ([i, _PI32, i, _PI32], i),
        "sx_load_mtx_f32": ([C.c_char_p, C.POINTER(i), C.POINTER(i), C.POINTER(i64),
                             C.POINTER(_PI32), C.POINTER(_PI32), C.POINTER(_PF)], i),
        "sx_load_mtx_f64": ([C.c_char_p, C.POINTER(i), C.POINTER(i), C.POINTER(i64),
                             C.POINTER(_PI32), C.POINTER(_PI32), C.POINTER(_PD)], i),
        "sx_split_col_windows": ([i, i, _PI32, _PI32, i, C.POINTER(i), C.POINTER(_PI32),
                                  C.POINTER(C.POINTER(i64)), C.POINTER(_PI32), C.POINTER(i)], i),
        "sx_plan_slide": ([i, _PI32, _PI32, i, C.POINTER(i), C.POINTER(_PI32), C.POINTER(i), C.POINTER(_PI32),
                           C.POINTER(i), C.POINTER(i)], i),
        "sx_plan_edge_lists": ([i, i, _PI32, _PI32, i, i, i, i64, i, C.POINTER(i), C.POINTER(_PI32), C.POINTER(i64),
                                C.POINTER(_PI32), C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(i64), C.POINTER(i),
                                C.POINTER(_PI32)], i),
        "sx_free": ([vp], None),
        "sx_sextans_invoke": ([vp, _PI32, _A8, _F4, _F8, _F8, i, i, i, i, i, i, i, _PD], i),
        "sx_sextans_last_kernel_ns": ([], C.c_double),
        "sx_images_A_words": ([i], i64),
        "sx_images_B_floats": ([i, i], i64),
        "sx_images_C_floats": ([i, i], i64),
        "sx_images_decode_A": ([_PI32, _A8, i, i, i, i, C.POINTER(i64), C.POINTER(_PI32),
                                C.POINTER(_PI32), C.POINTER(_PF)], i),
        "sx_images_decode_B": ([_F4, i, i, _PF], i),
        "sx_images_decode_C": ([_F8, i, i, _PF], i),
        "sx_images_encode_C": ([_PF, i, i, C.c_float, C.c_float, _F8, _F8], i),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _LIB = L
    return L


def _check(status):
    if status != 0:
        L = lib()
        raise SextansError(L.sx_status_name(status).decode(), L.sx_last_error().decode())


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float, SX_F32
    if dtype == np.float64:
        return "f64", C.c_double, SX_F64
    raise TypeError(f"float32 or float64 expected, got {dtype}")


def _host_ptr(a):
    return C.c_void_p(a.ctypes.data)


def _dev_ptr(t):
    # torch tensor / anything with data_ptr(), or a raw integer address
    return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))


def load_mtx(path, dtype=np.float32):
    """Matrix Market -> (M, K, nnz, rowptr, colidx, val) with the reference loader's
    semantics (sx_load_mtx_*; src/sparse_helper.h:169-259 + 475-509)."""
    suf, ct, _ = _suffix(dtype)
    M, K, nnz = C.c_int(), C.c_int(), C.c_int64()
    rp, ci, v = _PI32(), _PI32(), C.POINTER(ct)()
    L = lib()
    _check(getattr(L, f"sx_load_mtx_{suf}")(os.fsencode(path), C.byref(M), C.byref(K),
                                             C.byref(nnz), C.byref(rp), C.byref(ci), C.byref(v)))
    try:
        n = nnz.value
        rowptr = np.ctypeslib.as_array(rp, shape=(M.value + 1,)).copy()
        colidx = np.ctypeslib.as_array(ci, shape=(max(n, 1),))[:n].copy()
        val = np.ctypeslib.as_array(v, shape=(max(n, 1),))[:n].copy()
    finally:
        L.sx_free(rp), L.sx_free(ci), L.sx_free(v)
    return M.value, K.value, n, rowptr, colidx, val


def partition_rows(rowptr, parts):
    """nnz-balanced contiguous row blocks -> bounds[parts+1] (sx_partition_rows)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    bounds = np.empty(parts + 1, dtype=np.int32)
    _check(lib().sx_partition_rows(rowptr.size - 1, rowptr.ctypes.data_as(_PI32), parts,
                                   bounds.ctypes.data_as(_PI32)))
    return bounds


def plan_slide(M, rowptr, colidx, nchains):
    """Plan of the sliding-window kernel (sx_plan_slide) ->
    (steps [nsteps, 4], chains [nchains, 2], ring_rows, max_step_entries)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    ns, nc, ring, ent = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    st, ch = _PI32(), _PI32()
    L = lib()
    _check(L.sx_plan_slide(M, rowptr.ctypes.data_as(_PI32), colidx.ctypes.data_as(_PI32), nchains, C.byref(ns),
                           C.byref(st), C.byref(nc), C.byref(ch), C.byref(ring), C.byref(ent)))
    try:
        steps = np.ctypeslib.as_array(st, shape=(max(ns.value, 1) * 4,))[:ns.value * 4].reshape(-1, 4).copy()
        chains = np.ctypeslib.as_array(ch, shape=(max(nc.value, 1) * 2,))[:nc.value * 2].reshape(-1, 2).copy()
    finally:
        L.sx_free(st), L.sx_free(ch)
    return steps, chains, ring.value, ent.value


def plan_edge_lists(M, K, rowptr, colidx, row_bytes, elem_bytes, smem_budget, max_rows=32, nnz_target=0):
    """Plan of the edge-list kernel (sx_plan_edge_lists) ->
    (blocks [nblocks, 8], cols [ncols] int32, lcol [prow[M]] uint16, total_cols, max_smem, prow [M+1] int32);
    the streams are row-aligned (row r's entries at prow[r] + k, rows padded to multiples of 8 entries, pad
    entries 0; blocks[:, 2:4] in padded coordinates); nblocks == 0 if some row does not fit ``smem_budget``."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    nb, ms, tot, nc = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
    bl, co, lc, pr = _PI32(), _PI32(), C.POINTER(C.c_uint16)(), _PI32()
    L = lib()
    _check(L.sx_plan_edge_lists(M, K, rowptr.ctypes.data_as(_PI32), colidx.ctypes.data_as(_PI32), row_bytes, elem_bytes,
                                max_rows, nnz_target, smem_budget, C.byref(nb), C.byref(bl), C.byref(nc), C.byref(co),
                                C.byref(lc), C.byref(tot), C.byref(ms), C.byref(pr)))
    if nb.value == 0:
        L.sx_free(pr)
        return np.zeros((0, 8), np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint16), 0, 0, np.zeros(M + 1, np.int32)
    try:
        prow = np.ctypeslib.as_array(pr, shape=(M + 1,)).copy()
        n = int(prow[M])
        blocks = np.ctypeslib.as_array(bl, shape=(nb.value * 8,)).reshape(-1, 8).copy()
        cols = np.ctypeslib.as_array(co, shape=(max(nc.value, 1),))[:nc.value].copy()
        lcol = np.ctypeslib.as_array(lc, shape=(max(n, 1),))[:n].copy()
    finally:
        L.sx_free(bl), L.sx_free(co), L.sx_free(lc), L.sx_free(pr)
    return blocks, cols, lcol, tot.value, ms.value, prow


def split_col_windows(M, K, rowptr, colidx, window_rows):
    """Column windows of ``window_rows`` columns (sx_split_col_windows) ->
    (win_rowptr [nwin, M+1], win_base [nwin+1], order [nnz], ascending)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    nwin, asc = C.c_int(), C.c_int()
    wrp, order, base = _PI32(), _PI32(), C.POINTER(C.c_int64)()
    L = lib()
    _check(L.sx_split_col_windows(M, K, rowptr.ctypes.data_as(_PI32), colidx.ctypes.data_as(_PI32),
                                  window_rows, C.byref(nwin), C.byref(wrp), C.byref(base),
                                  C.byref(order), C.byref(asc)))
    try:
        n = int(rowptr[M])
        w = nwin.value
        win_rowptr = np.ctypeslib.as_array(wrp, shape=(w * (M + 1),)).reshape(w, M + 1).copy()
        win_base = np.ctypeslib.as_array(base, shape=(w + 1,)).copy()
        ordr = np.ctypeslib.as_array(order, shape=(max(n, 1),))[:n].copy()
    finally:
        L.sx_free(wrp), L.sx_free(base), L.sx_free(order)
    return win_rowptr, win_base, ordr, bool(asc.value)


# ---- FPGA channel images (the literal Sextans(...) argument list) ----------------------
def _chan(arrs, n, ctype, dtype):
    """n contiguous channel arrays -> (ctypes pointer array, keep-alive list)."""
    if len(arrs) != n:
        raise ValueError(f"{n} channel images expected, got {len(arrs)}")
    keep = []
    for a in arrs:
        if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous):
            raise ValueError(f"channel images must be contiguous {np.dtype(dtype).name} arrays")
        keep.append(a)
    return (C.POINTER(ctype) * n)(*[a.ctypes.data_as(C.POINTER(ctype)) for a in keep]), keep


def images_decode_A(ptr, images, M, K):
    """Edge-list images (src/sparse_helper.h:345-473) -> CSR (rowptr, colidx, val)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    num_ite, num_a_len = ptr.size - 1, int(ptr[-1])
    a, keep = _chan(images, 8, C.c_uint64, np.uint64)
    for im in keep:
        if im.size < 8 * num_a_len:
            raise ValueError("A channel image shorter than 8 * NUM_A_LEN words")
    nnz = C.c_int64()
    rp, ci, v = _PI32(), _PI32(), _PF()
    L = lib()
    _check(L.sx_images_decode_A(ptr.ctypes.data_as(_PI32), a, num_ite, num_a_len, M, K, C.byref(nnz),
                                C.byref(rp), C.byref(ci), C.byref(v)))
    try:
        n = nnz.value
        rowptr = np.ctypeslib.as_array(rp, shape=(M + 1,)).copy()
        colidx = np.ctypeslib.as_array(ci, shape=(max(n, 1),))[:n].copy()
        val = np.ctypeslib.as_array(v, shape=(max(n, 1),))[:n].copy()
    finally:
        L.sx_free(rp), L.sx_free(ci), L.sx_free(v)
    return rowptr, colidx, val


def _need(images, elems, what):
    for im in images:
        if im.size < elems:
            raise ValueError(f"{what} channel image shorter than {elems} floats")


def images_decode_B(images, K, N):
    """4 B channel images (src/sextans-host.cpp:158-171) -> column-major K x roundup(N,8)."""
    L = lib()
    b, keep = _chan(images, 4, C.c_float, np.float32)
    _need(keep, L.sx_images_B_floats(K, N), "B")
    out = np.empty(K * ((N + 7) // 8 * 8), dtype=np.float32)
    _check(L.sx_images_decode_B(b, K, N, out.ctypes.data_as(_PF)))
    return out


def images_decode_C(images, M, N):
    """8 C channel images (src/sextans-host.cpp:181-195) -> column-major M x roundup(N,8)."""
    L = lib()
    c, keep = _chan(images, 8, C.c_float, np.float32)
    _need(keep, L.sx_images_C_floats(M, N), "C")
    out = np.empty(M * ((N + 7) // 8 * 8), dtype=np.float32)
    _check(L.sx_images_decode_C(c, M, N, out.ctypes.data_as(_PF)))
    return out


def images_encode_C(Cm, M, N, alpha, beta, images_in, images_out):
    """Column-major C -> the 8 output images, pad rows as the FPGA writes them."""
    L = lib()
    Cm = np.ascontiguousarray(Cm, dtype=np.float32)
    assert Cm.size == M * ((N + 7) // 8 * 8)
    ci, keep_i = _chan(images_in, 8, C.c_float, np.float32)
    co, keep_o = _chan(images_out, 8, C.c_float, np.float32)
    _need(keep_i + keep_o, L.sx_images_C_floats(M, N), "C")
    _check(L.sx_images_encode_C(Cm.ctypes.data_as(_PF), M, N, alpha, beta, ci, co))


class _Pinned:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        _check(lib().sx_host_alloc(nbytes, C.byref(self.ptr)))

    def __del__(self):
        try:
            if self.ptr:
                lib().sx_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(n, dtype):
    """1-D page-locked host array (sx_host_alloc), for B and C at the boundary."""
    dtype = np.dtype(dtype)
    nbytes = max(1, n * dtype.itemsize)
    owner = _Pinned(nbytes)
    buf = (C.c_char * nbytes).from_address(owner.ptr.value)
    return _PinnedArray(np.frombuffer(buf, dtype=dtype, count=n), owner)


class _PinnedArray(np.ndarray):
    """ndarray view that keeps its page-locked allocation alive."""

    def __new__(cls, arr, owner):
        obj = arr.view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class Engine:
    """One GPU's SpMM context (``sx_ctx``).

    Mirrors the reference host's single device call: upload A once, then
    ``spmm(N, alpha, B, beta, C)`` with column-major host operands, exactly the
    arguments ``cpu_spmm_CSR`` takes (src/sparse_helper.h:262-272).
    """

    def __init__(self, device: int = 0, arith: int = STRICT):
        self._ctx = C.c_void_p()
        self._L = lib()
        _check(self._L.sx_create(device, C.byref(self._ctx)))
        self.device = device
        self.dtype = None
        self.M = self.K = self.nnz = 0
        if arith != STRICT:
            self.set_option(OPT_ARITH, arith)

    # -- lifecycle -----------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.sx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream):
        """Run on an existing CUDA stream (integer handle, e.g. torch's .cuda_stream)."""
        _check(self._L.sx_set_stream(self._ctx, C.c_void_p(cuda_stream or None)))

    def set_option(self, option, value):
        _check(self._L.sx_set_option(self._ctx, option, int(value)))

    def info(self, what) -> int:
        v = C.c_int64()
        _check(self._L.sx_get_info(self._ctx, what, C.byref(v)))
        return v.value

    @property
    def launches(self) -> int:
        return self.info(INFO_LAUNCHES)

    def synchronize(self):
        _check(self._L.sx_synchronize(self._ctx))

    # -- A ---------------------------------------------------------------------------
    def upload_csr(self, M, K, rowptr, colidx, val):
        suf, ct, _ = _suffix(val.dtype)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        val = np.ascontiguousarray(val)
        if rowptr.size != M + 1:
            raise ValueError("rowptr must have M+1 entries")
        _check(getattr(self._L, f"sx_upload_csr_{suf}")(
            self._ctx, M, K, int(colidx.size), rowptr.ctypes.data_as(_PI32),
            colidx.ctypes.data_as(_PI32), val.ctypes.data_as(C.POINTER(ct))))
        self.dtype, self.M, self.K, self.nnz = val.dtype, M, K, int(colidx.size)
        return self

    # -- the call ----------------------------------------------------------------------
    def spmm(self, N, alpha, B, beta, C_inout, rp_time=1, want_ns=True):
        """In place on ``C_inout`` (column-major 1-D, like the host program's vectors).
        Returns the kernel time in ns summed over the ``rp_time`` repeats, the value
        ``tapa::invoke`` returns in the reference (src/sextans-host.cpp:237); with
        ``want_ns=False`` the C call gets ``kernel_ns = NULL`` and returns ``None``."""
        suf, ct, _ = _suffix(self.dtype)
        if B.dtype != self.dtype or not B.flags.c_contiguous:
            B = np.ascontiguousarray(B, dtype=self.dtype)
        if C_inout.dtype != self.dtype or not C_inout.flags.c_contiguous:
            raise ValueError("C must be a contiguous array of the matrix dtype")
        if B.size != self.K * N or C_inout.size != self.M * N:
            raise ValueError("B must hold K*N and C must hold M*N elements")
        fn = self._L.sx_spmm_f64 if suf == "f64" else self._L.sx_spmm_f32
        if not want_ns:   # the short way through: nothing to allocate, nothing to read back
            status = fn(self._ctx, N, ct(alpha), B.ctypes.data, ct(beta), C_inout.ctypes.data, rp_time, None)
            if status != 0:
                _check(status)
            return None
        ns = C.c_double()
        _check(fn(self._ctx, N, ct(alpha), _host_ptr(B), ct(beta), _host_ptr(C_inout), rp_time, C.byref(ns)))
        return ns.value

    def spmm_enqueue(self, N, alpha, B, beta, C_inout):
        """The host-facing call without its final host sync (sx_spmm_enqueue_*): B and C_inout (page-locked,
        of the matrix dtype) must stay valid and untouched until ``synchronize()``."""
        suf, ct, _ = _suffix(self.dtype)
        if B.dtype != self.dtype or C_inout.dtype != self.dtype or not B.flags.c_contiguous or not C_inout.flags.c_contiguous:
            raise ValueError("B and C must be contiguous arrays of the matrix dtype")
        if B.size != self.K * N or C_inout.size != self.M * N:
            raise ValueError("B must hold K*N and C must hold M*N elements")
        fn = self._L.sx_spmm_enqueue_f64 if suf == "f64" else self._L.sx_spmm_enqueue_f32
        status = fn(self._ctx, N, ct(alpha), B.ctypes.data, ct(beta), C_inout.ctypes.data)
        if status != 0:
            _check(status)

    def spmm_staged_B(self, N, alpha, beta, C_inout):
        """One blocking SpMM on the B image the context already holds (stage_B, a peer's push, a
        broadcast into device_B) and the host C (column-major 1-D, in place)."""
        suf, ct, _ = _suffix(self.dtype)
        if C_inout.dtype != self.dtype or not C_inout.flags.c_contiguous or C_inout.size != self.M * N:
            raise ValueError("C must be a contiguous array of M*N elements of the matrix dtype")
        _check(getattr(self._L, f"sx_spmm_staged_B_{suf}")(self._ctx, N, ct(alpha), ct(beta), _host_ptr(C_inout)))

    def sextans_invoke(self, ptr, A_images, B_images, Cin_images, Cout_images, M, K, P_N,
                       alpha_u, beta_u):
        """The reference's device call with its own arguments (src/sextans.h:20-26):
        FPGA channel images in, C images out (written in place into ``Cout_images``).
        Returns the kernel time in ns like ``tapa::invoke``."""
        L = self._L
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        num_ite, num_a_len = ptr.size - 1, int(ptr[-1])
        N = P_N & 0xFFFF
        a, ka = _chan(A_images, 8, C.c_uint64, np.uint64)
        b, kb = _chan(B_images, 4, C.c_float, np.float32)
        ci, kci = _chan(Cin_images, 8, C.c_float, np.float32)
        co, kco = _chan(Cout_images, 8, C.c_float, np.float32)
        for im in ka:
            if im.size < L.sx_images_A_words(num_a_len):
                raise ValueError("A channel image shorter than 8 * NUM_A_LEN words")
        _need(kb, L.sx_images_B_floats(K, N), "B")
        _need(kci + kco, L.sx_images_C_floats(M, N), "C")
        ns = C.c_double()
        _check(L.sx_sextans_invoke(self._ctx, ptr.ctypes.data_as(_PI32), a, b, ci, co, num_ite,
                                   num_a_len, M, K, P_N, alpha_u, beta_u, C.byref(ns)))
        self.dtype, self.M, self.K = np.dtype(np.float32), M, K
        self.nnz = self.info(INFO_NNZ)
        return ns.value

    # -- staged ------------------------------------------------------------------------
    def stage_B(self, N, B):
        suf = _suffix(self.dtype)[0]
        B = np.ascontiguousarray(B, dtype=self.dtype)
        assert B.size == self.K * N
        _check(getattr(self._L, f"sx_stage_B_{suf}")(self._ctx, N, _host_ptr(B)))

    def stage_C(self, N, Cin):
        suf = _suffix(self.dtype)[0]
        Cin = np.ascontiguousarray(Cin, dtype=self.dtype)
        assert Cin.size == self.M * N
        _check(getattr(self._L, f"sx_stage_C_{suf}")(self._ctx, N, _host_ptr(Cin)))

    def launch(self, alpha, beta, rp_time=1):
        suf, ct, _ = _suffix(self.dtype)
        ns = C.c_double()
        _check(getattr(self._L, f"sx_launch_{suf}")(self._ctx, ct(alpha), ct(beta), rp_time, C.byref(ns)))
        return ns.value

    def fetch_C(self, out):
        suf = _suffix(self.dtype)[0]
        assert out.dtype == self.dtype and out.flags.c_contiguous
        _check(getattr(self._L, f"sx_fetch_C_{suf}")(self._ctx, _host_ptr(out)))
        return out

    def device_B(self, N):
        """(device address, bytes) of the context's row-major B image."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(self._L.sx_device_B(self._ctx, N, C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- peer memory -----------------------------------------------------------------------
    def device_alloc(self, nbytes) -> int:
        p = C.c_void_p()
        _check(self._L.sx_device_alloc(self._ctx, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr):
        _check(self._L.sx_device_free(self._ctx, C.c_void_p(ptr)))

    def ipc_export(self, ptr) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._L.sx_ipc_export(self._ctx, C.c_void_p(ptr), buf))
        return buf.raw

    def ipc_export_ref(self, ptr):
        """(handle of the allocation that contains ptr, offset of ptr inside it) -- what a peer needs."""
        off = C.c_size_t()
        _check(self._L.sx_ipc_offset(self._ctx, C.c_void_p(ptr), C.byref(off)))
        return self.ipc_export(ptr), off.value

    def ipc_import(self, handle: bytes) -> int:
        p = C.c_void_p()
        _check(self._L.sx_ipc_import(self._ctx, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        _check(self._L.sx_ipc_close(self._ctx, C.c_void_p(ptr)))

    def push_B(self, image_ptr, nbytes, peer_image_ptrs, peer_ready_ptrs, done_flags_ptr, pushes_ptr):
        """Copy a row-major B image into every peer's (one kernel); see sx_push_B."""
        n = len(peer_image_ptrs)
        imgs = (C.c_void_p * n)(*peer_image_ptrs)
        rdy = (C.c_void_p * n)(*peer_ready_ptrs)
        _check(self._L.sx_push_B(self._ctx, C.c_void_p(image_ptr), nbytes, imgs, rdy, n, C.c_void_p(done_flags_ptr),
                                 C.c_void_p(pushes_ptr)))

    def fuse_push(self, peer_image_ptrs, peer_ready_ptrs, done_flags_ptr, pushes_ptr):
        """The next SpMM launch also pushes the B image it reads to the peers (sx_spmm_fuse_push)."""
        n = len(peer_image_ptrs)
        imgs = (C.c_void_p * n)(*peer_image_ptrs)
        rdy = (C.c_void_p * n)(*peer_ready_ptrs)
        _check(self._L.sx_spmm_fuse_push(self._ctx, imgs, rdy, n, C.c_void_p(done_flags_ptr), C.c_void_p(pushes_ptr)))

    def fuse_push_deferred(self, peer_image_ptrs, peer_ready_ptrs, done_flags_ptr, pushes_ptr):
        """As fuse_push, the publication left to a later launch (sx_spmm_fuse_push_deferred)."""
        n = len(peer_image_ptrs)
        imgs = (C.c_void_p * n)(*peer_image_ptrs)
        rdy = (C.c_void_p * n)(*peer_ready_ptrs)
        _check(self._L.sx_spmm_fuse_push_deferred(self._ctx, imgs, rdy, n, C.c_void_p(done_flags_ptr), C.c_void_p(pushes_ptr)))

    def fuse_publish(self, peer_ready_ptrs, pushes_ptr):
        """The next SpMM launch publishes an earlier launch's push (sx_spmm_fuse_publish)."""
        n = len(peer_ready_ptrs)
        rdy = (C.c_void_p * n)(*peer_ready_ptrs)
        _check(self._L.sx_spmm_fuse_publish(self._ctx, rdy, n, C.c_void_p(pushes_ptr)))

    def push_publish(self, peer_ready_ptrs, pushes_ptr):
        """Publish an earlier launch's push with a one-warp kernel (sx_push_publish)."""
        n = len(peer_ready_ptrs)
        rdy = (C.c_void_p * n)(*peer_ready_ptrs)
        _check(self._L.sx_push_publish(self._ctx, rdy, n, C.c_void_p(pushes_ptr)))

    def expect_push(self, ready_ptr, epoch_ptr, done_ptr):
        """The next SpMM launch waits for the push into its B image and acknowledges it (sx_spmm_expect_push)."""
        _check(self._L.sx_spmm_expect_push(self._ctx, C.c_void_p(ready_ptr), C.c_void_p(epoch_ptr), C.c_void_p(done_ptr)))

    # -- device-resident -----------------------------------------------------------------
    def spmm_device(self, N, alpha, dB, ldb, beta, dCin, dCout, ldc):
        """Enqueue one SpMM on row-major device operands (torch tensors or addresses)."""
        suf, ct, _ = _suffix(self.dtype)
        _check(getattr(self._L, f"sx_spmm_device_{suf}")(
            self._ctx, N, ct(alpha), _dev_ptr(dB), ldb, ct(beta), _dev_ptr(dCin), _dev_ptr(dCout), ldc))

    def spmm_device_batch(self, N, nb, alpha, dB, ldb, strideB, beta, dCin, dCout, ldc, strideC):
        """Enqueue nb SpMMs with the same A: operand b at base + b * stride elements (one launch on
        the edge-list kernel)."""
        suf, ct, _ = _suffix(self.dtype)
        _check(getattr(self._L, f"sx_spmm_device_batch_{suf}")(
            self._ctx, N, nb, ct(alpha), _dev_ptr(dB), ldb, strideB, ct(beta), _dev_ptr(dCin), _dev_ptr(dCout), ldc, strideC))

    def colmajor_to_rowmajor(self, rows, cols, d_src, d_dst, ld_dst, dtype=None):
        code = _suffix(dtype or self.dtype)[2]
        _check(self._L.sx_colmajor_to_rowmajor(self._ctx, code, rows, cols, _dev_ptr(d_src),
                                               _dev_ptr(d_dst), ld_dst))

    def rowmajor_to_colmajor(self, rows, cols, d_src, ld_src, d_dst, dtype=None):
        code = _suffix(dtype or self.dtype)[2]
        _check(self._L.sx_rowmajor_to_colmajor(self._ctx, code, rows, cols, _dev_ptr(d_src),
                                               ld_src, _dev_ptr(d_dst)))
