"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``sextans_b200/`` does.  ``liboracle.so`` is the C restatement
(``oracle/spmm_oracle.c``); ``_ref/libsextans_ref.so`` is the reference's own
header compiled unmodified (``oracle/ref_shim.cpp``), present when it was built in
the authoring container (it travels to the GPU box as a prebuilt file).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_I = C.c_int
_PI = C.POINTER(C.c_int)
_PF = C.POINTER(C.c_float)
_PD = C.POINTER(C.c_double)

ERRORS = {1: "cannot open file", 2: "bad Matrix Market banner", 3: "bad size line",
          4: "not a coordinate file", 5: "complex matrices unsupported",
          6: "index < 1", 7: "out of memory"}


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and _ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        for suf, ct, pt in (("f32", C.c_float, _PF), ("f64", C.c_double, _PD)):
            f = getattr(L, f"sx_oracle_spmm_csr_{suf}")
            f.argtypes = [_I, _I, _I, _PI, _PI, pt, ct, pt, ct, pt]
            f.restype = _I
            f = getattr(L, f"sx_oracle_spmm_csr_mt_{suf}")
            f.argtypes = [_I, _I, _I, _PI, _PI, pt, ct, pt, ct, pt, _I]
            f.restype = _I
            f = getattr(L, f"sx_oracle_spmm_csr_rows_{suf}")
            f.argtypes = [_I, _I, _I, _PI, _PI, pt, ct, pt, ct, pt, _PI, _I, pt]
            f.restype = _I
            f = getattr(L, f"sx_oracle_init_dense_{suf}")
            f.argtypes = [_I, _I, _I, pt, pt]
            f.restype = None
            f = getattr(L, f"sx_oracle_load_mtx_{suf}")
            f.argtypes = [C.c_char_p, _PI, _PI, _PI, C.POINTER(_PI), C.POINTER(_PI),
                          C.POINTER(pt), C.c_char * 4]
            f.restype = _I
        L.sx_oracle_verify_f32.argtypes = [C.c_int64, _PF, _PF, _I, _I, _PF]
        L.sx_oracle_verify_f32.restype = C.c_int64
        L.sx_oracle_sextans_images.argtypes = [_PI, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(_PF),
                                               C.POINTER(_PF), C.POINTER(_PF), _I, _I, _I, _I, _I, _I]
        L.sx_oracle_sextans_images.restype = _I
        L.sx_oracle_free.argtypes = [C.c_void_p]
        L.sx_oracle_max_threads.restype = _I
        _LIB = L
    return _LIB


def ref():
    """The compiled reference header, or None when it is not available."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libsextans_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.sxref_load_csr.argtypes = [C.c_char_p, _PI, _PI, _PI, C.POINTER(_PI),
                                     C.POINTER(_PI), C.POINTER(_PF)]
        R.sxref_load_csr.restype = _I
        R.sxref_cpu_spmm_csr.argtypes = [_I, _I, _I, _I, C.c_float, _PI, _PI, _PF, _PF,
                                         C.c_float, _PF]
        R.sxref_cpu_spmm_csr.restype = C.c_double
        R.sxref_edge_list_slots.argtypes = [_I, _I, _I, _PI, _PI, _PF, _I, _I, _I]
        R.sxref_edge_list_slots.restype = _I
        R.sxref_free.argtypes = [C.c_void_p]
        if hasattr(R, "sxref_build_images"):
            R.sxref_build_images.argtypes = [_I, _I, _I, _PI, _PI, _PF, C.POINTER(_PI), _PI,
                                             C.POINTER(C.POINTER(C.c_ulong)), C.POINTER(C.c_long)]
            R.sxref_build_images.restype = _I
        _REF = R
    return _REF


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(f"oracle supports float32/float64, not {dtype}")


def _csr_args(rowptr, colidx, val, dtype):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=dtype)
    return rowptr, colidx, val


def spmm_csr(M, N, K, rowptr, colidx, val, alpha, B, beta, C_inout, threads=1):
    """C <- alpha*A*B + beta*C in place; B, C column-major 1-D arrays (ld K / M).

    threads == 1 is the reference's own single-threaded traversal; threads > 1
    runs rows in parallel (bitwise the same result).
    """
    suf, ct = _suffix(C_inout.dtype)
    rowptr, colidx, val = _csr_args(rowptr, colidx, val, C_inout.dtype)
    B = np.ascontiguousarray(B, dtype=C_inout.dtype)
    assert C_inout.flags.c_contiguous and C_inout.size == M * N and B.size == K * N
    args = [M, N, K, _ptr(rowptr, C.c_int), _ptr(colidx, C.c_int), _ptr(val, ct),
            ct(alpha), _ptr(B, ct), ct(beta), _ptr(C_inout, ct)]
    if threads == 1:
        rc = getattr(lib(), f"sx_oracle_spmm_csr_{suf}")(*args)
    else:
        rc = getattr(lib(), f"sx_oracle_spmm_csr_mt_{suf}")(*args, int(threads))
    if rc:
        raise RuntimeError(f"oracle spmm failed: {ERRORS.get(rc, rc)}")
    return C_inout


def spmm_csr_rows(M, N, K, rowptr, colidx, val, alpha, B, beta, C_in, rows):
    """Oracle result for the listed rows only -> array [len(rows), N]."""
    suf, ct = _suffix(C_in.dtype)
    rowptr, colidx, val = _csr_args(rowptr, colidx, val, C_in.dtype)
    B = np.ascontiguousarray(B, dtype=C_in.dtype)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.empty((rows.size, N), dtype=C_in.dtype)
    rc = getattr(lib(), f"sx_oracle_spmm_csr_rows_{suf}")(
        M, N, K, _ptr(rowptr, C.c_int), _ptr(colidx, C.c_int), _ptr(val, ct), ct(alpha),
        _ptr(B, ct), ct(beta), _ptr(C_in, ct), _ptr(rows, C.c_int), rows.size, _ptr(out, ct))
    if rc:
        raise RuntimeError(f"oracle spmm rows failed: {ERRORS.get(rc, rc)}")
    return out


def init_dense(M, K, N, dtype):
    """The host driver's B (all ones) and C_in ((m+1)(n+1)/M/N), column-major."""
    suf, ct = _suffix(dtype)
    B = np.empty(K * N, dtype=dtype)
    Cm = np.empty(M * N, dtype=dtype)
    getattr(lib(), f"sx_oracle_init_dense_{suf}")(M, K, N, _ptr(B, ct), _ptr(Cm, ct))
    return B, Cm


def verify_f32(cpu, dev, M, N):
    """Reference pass/fail criterion -> (mismatches, percent, passed)."""
    cpu = np.ascontiguousarray(cpu, dtype=np.float32)
    dev = np.ascontiguousarray(dev, dtype=np.float32)
    pct = C.c_float(0)
    n = lib().sx_oracle_verify_f32(cpu.size, _ptr(cpu, C.c_float), _ptr(dev, C.c_float),
                                   M, N, C.byref(pct))
    return int(n), float(pct.value), bool(pct.value < 2.0)


def _take(ptr, n, dtype, free):
    arr = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dtype, copy=True)
    free(ptr)
    return arr


def load_mtx(path, dtype=np.float32):
    """Oracle loader -> (M, K, nnz, rowptr, colidx, val, typecode)."""
    suf, ct = _suffix(dtype)
    M, K, nnz = _I(), _I(), _I()
    rp, ci, v = _PI(), _PI(), C.POINTER(ct)()
    code = (C.c_char * 4)()
    rc = getattr(lib(), f"sx_oracle_load_mtx_{suf}")(
        os.fsencode(path), C.byref(M), C.byref(K), C.byref(nnz), C.byref(rp), C.byref(ci),
        C.byref(v), code)
    if rc:
        raise RuntimeError(f"oracle loader: {ERRORS.get(rc, rc)} ({path})")
    free = lib().sx_oracle_free
    return (M.value, K.value, nnz.value, _take(rp, M.value + 1, np.int32, free),
            _take(ci, nnz.value, np.int32, free), _take(v, nnz.value, dtype, free),
            bytes(code).decode())


def ref_load_csr(path):
    """The reference's read_suitsparse_matrix(CSC)+CSC_2_CSR -> same tuple (f32)."""
    R = ref()
    if R is None:
        raise RuntimeError("oracle/_ref/libsextans_ref.so not built")
    M, K, nnz = _I(), _I(), _I()
    rp, ci, v = _PI(), _PI(), _PF()
    R.sxref_load_csr(os.fsencode(path), C.byref(M), C.byref(K), C.byref(nnz), C.byref(rp),
                     C.byref(ci), C.byref(v))
    return (M.value, K.value, nnz.value, _take(rp, M.value + 1, np.int32, R.sxref_free),
            _take(ci, nnz.value, np.int32, R.sxref_free),
            _take(v, nnz.value, np.float32, R.sxref_free))


def ref_spmm_csr(M, N, K, rowptr, colidx, val, alpha, B, beta, C_inout):
    """The reference's cpu_spmm_CSR (fp32, 1 thread).  Returns seconds inside it."""
    R = ref()
    if R is None:
        raise RuntimeError("oracle/_ref/libsextans_ref.so not built")
    rowptr, colidx, val = _csr_args(rowptr, colidx, val, np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    assert C_inout.dtype == np.float32 and C_inout.flags.c_contiguous
    ns = R.sxref_cpu_spmm_csr(M, N, K, int(colidx.size), float(alpha),
                              _ptr(rowptr, C.c_int), _ptr(colidx, C.c_int),
                              _ptr(val, C.c_float), _ptr(B, C.c_float), float(beta),
                              _ptr(C_inout, C.c_float))
    return ns * 1e-9


# ---- FPGA channel images (checker side of sx_sextans_invoke) -------------------------
def ref_build_images(M, K, rowptr, colidx, val):
    """The reference's own A preprocessing (generate_edge_list_for_all_PEs +
    edge_list_64bit, src/sparse_helper.h:345-473, constants of src/sextans.h:7-12) on a
    CSR with ascending columns -> (ptr int32[NUM_ITE+1], [8 x uint64 image], NUM_A_LEN)."""
    R = ref()
    if R is None or not hasattr(R, "sxref_build_images"):
        raise RuntimeError("oracle/_ref/libsextans_ref.so (with sxref_build_images) not built")
    rowptr, colidx, val = _csr_args(rowptr, colidx, val, np.float32)
    ptr, ptr_len, image_len = _PI(), _I(), C.c_long()
    imgs = (C.POINTER(C.c_ulong) * 8)()
    num_a_len = R.sxref_build_images(M, K, int(colidx.size), _ptr(rowptr, C.c_int),
                                     _ptr(colidx, C.c_int), _ptr(val, C.c_float), C.byref(ptr),
                                     C.byref(ptr_len), imgs, C.byref(image_len))
    p = _take(ptr, ptr_len.value, np.int32, R.sxref_free)
    images = []
    for c in range(8):
        a = np.ctypeslib.as_array(imgs[c], shape=(max(image_len.value, 1),))[:image_len.value]
        images.append(a.astype(np.uint64, copy=True))
        R.sxref_free(imgs[c])
    return p, images, int(num_a_len)


def _round_up(x, m):
    return (x + m - 1) // m * m


def pack_B_images(B, K, N):
    """Column-major B (K x N, N % 8 == 0) -> the 4 channel images the host program builds
    (src/sextans-host.cpp:148-171, NUM_CH_B == 4 branch), chunk-padded like there."""
    assert N % 8 == 0
    B = np.asarray(B, dtype=np.float32).reshape(N, K)         # [n, k]
    colsz = _round_up(K, 8) * 2
    chunk = _round_up(colsz * (N // 8), 1024)
    imgs = [np.zeros(chunk, dtype=np.float32) for _ in range(4)]
    k = np.arange(K, dtype=np.int64)
    for n in range(N):
        pos = (k // 8) * 16 + (n % 2) * 8 + k % 8 + colsz * (n // 8)
        imgs[(n // 2) % 4][pos] = B[n]
    return imgs


def pack_C_images(Cm, M, N, fill_pad=None):
    """Column-major C (M x N) -> the 8 channel images (src/sextans-host.cpp:173-195).
    fill_pad: value for the rows that pad M to a multiple of 16 (the host leaves 0)."""
    assert N % 8 == 0
    Cm = np.asarray(Cm, dtype=np.float32).reshape(N, M)
    colsz = _round_up(M, 16)
    chunk = _round_up(colsz * (N // 8), 1024)
    imgs = [np.zeros(chunk, dtype=np.float32) for _ in range(8)]
    m = np.arange(M, dtype=np.int64)
    mp = np.arange(M, colsz, dtype=np.int64)
    for n in range(N):
        pos = colsz * (n // 8) + (m // 8) * 8 + n % 8
        for c in range(8):
            sel = m % 8 == c
            imgs[c][pos[sel]] = Cm[n][sel]
        if fill_pad is not None:
            ppos = colsz * (n // 8) + (mp // 8) * 8 + n % 8
            for c in range(8):
                imgs[c][ppos[mp % 8 == c]] = fill_pad
    return imgs


def unpack_C_images(imgs, M, N):
    """The read-back un-interleave of the verification loop (src/sextans-host.cpp:264-270)
    -> column-major C (M x N)."""
    colsz = _round_up(M, 16)
    out = np.empty((N, M), dtype=np.float32)
    m = np.arange(M, dtype=np.int64)
    for n in range(N):
        pos = colsz * (n // 8) + (m // 8) * 8 + n % 8
        for c in range(8):
            sel = m % 8 == c
            out[n][sel] = imgs[c][pos[sel]]
    return out.ravel()


def sextans_images(ptr, A_images, B_images, Cin_images, M, K, P_N, alpha_u, beta_u):
    """Functional model of the accelerator dataflow (src/sextans.cpp:285-570, 196-233) on
    the channel images themselves -> the 8 C output images (new arrays, zero beyond the
    words the hardware writes)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    a = [np.ascontiguousarray(x, dtype=np.uint64) for x in A_images]
    b = [np.ascontiguousarray(x, dtype=np.float32) for x in B_images]
    ci = [np.ascontiguousarray(x, dtype=np.float32) for x in Cin_images]
    co = [np.zeros_like(x) for x in ci]
    pa = (C.POINTER(C.c_uint64) * 8)(*[x.ctypes.data_as(C.POINTER(C.c_uint64)) for x in a])
    pb = (_PF * 4)(*[_ptr(x, C.c_float) for x in b])
    pci = (_PF * 8)(*[_ptr(x, C.c_float) for x in ci])
    pco = (_PF * 8)(*[_ptr(x, C.c_float) for x in co])
    rc = lib().sx_oracle_sextans_images(_ptr(ptr, C.c_int), pa, pb, pci, pco, ptr.size - 1, M, K,
                                        P_N, alpha_u, beta_u)
    if rc:
        raise RuntimeError(f"dataflow model failed: {rc}")
    return co


def pack_scalars(N, rp_time, alpha, beta):
    """P_N, alpha_u, beta_u as the host packs them (src/sextans-host.cpp:223-229)."""
    f = np.array([alpha, beta], dtype=np.float32).view(np.int32)
    return (int(rp_time) << 16) | int(N), int(f[0]), int(f[1])
