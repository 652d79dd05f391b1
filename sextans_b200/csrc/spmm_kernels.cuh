// spmm_kernels.cuh -- sm_100a kernels of the SpMM hot path.
//
// What of the reference these stand in for (citations relative to the reference
// tree): the whole accelerator dataflow of src/sextans.cpp:836-984 --
//   read_A / read_B / read_C / write_C           (:75-194)  -> global loads/stores here
//   PEG_Bmtx: val * B[col][0..7]                 (:285-423) -> the B-row gather + multiply
//   PEG_Cmtx: local_C[row] += abvec              (:425-570) -> register accumulators per row
//   FloatvMultConst / FloatvAddFloatv epilogue   (:196-233) -> fused alpha/beta epilogue
// and they compute exactly what cpu_spmm_CSR (src/sparse_helper.h:262-290) computes.
//
// Device data layout: A in CSR (int32 rowptr/colidx, T values); B, C ROW-major with
// a leading dimension that is a multiple of 16 bytes, so that one nonzero's B row is
// read with 16-byte vector loads by a sub-warp "row group" of G lanes.  Each lane
// owns 16 bytes (4 fp32 / 2 fp64 columns) of VPL interleaved vectors of its row's
// accumulator; the group walks the row's nonzeros IN STORED ORDER, so in strict mode
// (separately rounded multiply and add, sparse_helper.h:283) the result is
// bit-identical to the reference loop.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sx {

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int E = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int E = 2; };

// ---- scalar arithmetic with an explicit rounding contract -------------------
template <bool STRICT> __device__ __forceinline__ float mac(float acc, float a, float b) {
    if (STRICT) return __fadd_rn(acc, __fmul_rn(a, b));  // never contracted to FMA
    return fmaf(a, b, acc);
}
template <bool STRICT> __device__ __forceinline__ double mac(double acc, double a, double b) {
    if (STRICT) return __dadd_rn(acc, __dmul_rn(a, b));
    return fma(a, b, acc);
}
// alpha*acc + beta*c: two rounded products and a rounded sum (sparse_helper.h:287,
// sextans.cpp:211,229)
template <bool STRICT> __device__ __forceinline__ float axpby(float alpha, float acc, float beta, float c) {
    if (STRICT) return __fadd_rn(__fmul_rn(alpha, acc), __fmul_rn(beta, c));
    return fmaf(alpha, acc, beta * c);
}
template <bool STRICT> __device__ __forceinline__ double axpby(double alpha, double acc, double beta, double c) {
    if (STRICT) return __dadd_rn(__dmul_rn(alpha, acc), __dmul_rn(beta, c));
    return fma(alpha, acc, beta * c);
}
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// ---- 16-byte vector helpers ---------------------------------------------------
__device__ __forceinline__ void vzero(float4 &v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(double2 &v) { v = make_double2(0.0, 0.0); }

template <bool STRICT> __device__ __forceinline__ void vmac(float4 &acc, float a, const float4 &b) {
    acc.x = mac<STRICT>(acc.x, a, b.x); acc.y = mac<STRICT>(acc.y, a, b.y);
    acc.z = mac<STRICT>(acc.z, a, b.z); acc.w = mac<STRICT>(acc.w, a, b.w);
}
template <bool STRICT> __device__ __forceinline__ void vmac(double2 &acc, double a, const double2 &b) {
    acc.x = mac<STRICT>(acc.x, a, b.x); acc.y = mac<STRICT>(acc.y, a, b.y);
}
template <bool STRICT> __device__ __forceinline__ float4 vaxpby(float alpha, const float4 &acc, float beta, const float4 &c) {
    return make_float4(axpby<STRICT>(alpha, acc.x, beta, c.x), axpby<STRICT>(alpha, acc.y, beta, c.y),
                       axpby<STRICT>(alpha, acc.z, beta, c.z), axpby<STRICT>(alpha, acc.w, beta, c.w));
}
template <bool STRICT> __device__ __forceinline__ double2 vaxpby(double alpha, const double2 &acc, double beta, const double2 &c) {
    return make_double2(axpby<STRICT>(alpha, acc.x, beta, c.x), axpby<STRICT>(alpha, acc.y, beta, c.y));
}
__device__ __forceinline__ void vadd(float4 &a, const float4 &b) {
    a.x = add_rn(a.x, b.x); a.y = add_rn(a.y, b.y); a.z = add_rn(a.z, b.z); a.w = add_rn(a.w, b.w);
}
__device__ __forceinline__ void vadd(double2 &a, const double2 &b) {
    a.x = add_rn(a.x, b.x); a.y = add_rn(a.y, b.y);
}
__device__ __forceinline__ float4 vshfl_xor(unsigned m, const float4 &v, int off) {
    return make_float4(__shfl_xor_sync(m, v.x, off), __shfl_xor_sync(m, v.y, off),
                       __shfl_xor_sync(m, v.z, off), __shfl_xor_sync(m, v.w, off));
}
__device__ __forceinline__ double2 vshfl_xor(unsigned m, const double2 &v, int off) {
    return make_double2(__shfl_xor_sync(m, v.x, off), __shfl_xor_sync(m, v.y, off));
}
// read-only 16-byte gather of a B-row piece (ld.global.nc)
__device__ __forceinline__ float4 ldg_vec(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ double2 ldg_vec(const double2 *p) { return __ldg(p); }

// Accumulate nonzeros [begin, end) of one row, visiting chunks of G nonzeros at
// positions begin + first*G, begin + (first+stride)*G, ...  The G lanes of the group
// fetch a chunk's (col, val) pairs with one coalesced load each and hand them round
// by shuffle; the next chunk's pair is fetched before the current one is consumed.
template <typename T, int G, int VPL, bool STRICT>
__device__ __forceinline__ void accumulate_range(
    typename VecOf<T>::type (&acc)[VPL], const int begin, const int end, const int first,
    const int stride, const int lg, const unsigned gmask, const int nvec,
    const int *__restrict__ colidx, const T *__restrict__ val, const T *__restrict__ B,
    const int64_t ldb) {
    using V = typename VecOf<T>::type;
    constexpr int U = G < 8 ? G : 8;  // B-row gathers kept in flight per lane
    int base = begin + first * G;
    int c = 0;
    T a = T(0);
    if (base + lg < end) { c = __ldg(colidx + base + lg); a = __ldg(val + base + lg); }
    for (; base < end; base += stride * G) {
        const int nbase = base + stride * G;
        int cn = 0;
        T an = T(0);
        if (nbase + lg < end) { cn = __ldg(colidx + nbase + lg); an = __ldg(val + nbase + lg); }
        const int cnt = end - base;  // >= 1; entries t >= cnt of this chunk do not exist
#pragma unroll
        for (int t0 = 0; t0 < G; t0 += U) {
            if (t0 < cnt) {  // group-uniform
                V b[U][VPL];
                T av[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int cc = __shfl_sync(gmask, c, t0 + u, G);
                    av[u] = __shfl_sync(gmask, a, t0 + u, G);
                    const V *brow = reinterpret_cast<const V *>(B + (int64_t)cc * ldb);
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        const int vi = lg + v * G;
                        if (t0 + u < cnt && vi < nvec) b[u][v] = ldg_vec(brow + vi);
                        else vzero(b[u][v]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (t0 + u < cnt) {
#pragma unroll
                        for (int v = 0; v < VPL; ++v) vmac<STRICT>(acc[v], av[u], b[u][v]);
                    }
                }
            }
        }
        c = cn;
        a = an;
    }
}

// ---- cache-policy loads/stores ---------------------------------------------------
// A's colidx/val and C are touched once per SpMM: they bypass L1 and are marked
// evict-first in L2, so that the 126 MB L2 (and the L1s) keep B rows, which are the
// only data with reuse.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_stream(const int *p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream(const float *p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_stream(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
// C_in may alias C_out (in place), so no .nc here
__device__ __forceinline__ float4 ld_once(const float4 *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_once(const double2 *p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_once(float4 *p, const float4 &v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_once(double2 *p, const double2 &v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;"
                 :: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on the mbarrier; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// L2 prefetch of a contiguous piece of global memory (a hint: nothing is consumed, and L2 is
// the point of coherence, so it may be issued before the data's producer has finished)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); completion with cp_async_wait_all
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// element-sized asynchronous copies (the one-kernel host call: straight from page-locked host memory), commit groups
__device__ __forceinline__ void cp_async_elem(float *smem_dst, const float *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double *smem_dst, const double *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// at most n (0..7) of the most recently committed groups still pending
__device__ __forceinline__ void cp_async_wait_pending(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_l2(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ double2 ld_l2(const double2 *p) { return __ldcg(p); }
// programmatic dependent launch (no-ops when the grid was launched without the attribute)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// step flags in peer-visible memory
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// generic-proxy writes (also a peer's, once acquired) before later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- main kernel, staged: A streams through shared memory by TMA -------------------
// A work item is a run of consecutive whole rows, or a piece of one long row, holding
// about the same number of nonzeros as every other item (built at upload, sx_api.cu:
// build_plan) -- the GPU analogue of the reference's row -> PE assignment with equal-
// length PE lists (src/sparse_helper.h:345-403).  One lane group (G lanes = one B/C
// row of 16-byte vectors) owns one item:
//   * its slice of colidx[] and val[] streams through a double-buffered shared-memory
//     tile, filled by TMA bulk copies that complete on the group's own mbarriers (the
//     TAPA stream channels read_A -> Scatter -> PEG of src/sextans.cpp:75-100,785-800
//     become TMA + smem staging);
//   * B rows are gathered with 16-byte LDGs through a register ring: the gathers of
//     batch q+1 are issued before batch q is accumulated, so a lane keeps up to 2*U
//     vectors in flight (scripts/micro/gather_bench.cu: plain LDG reaches 16-19 TB/s
//     from L2, per-row TMA bulk copies 6.5-11 TB/s -- so B rows do not go through TMA);
//   * a row is closed (fused alpha/beta epilogue, src/sextans.cpp:196-233) when the
//     stream position reaches the next row pointer.  Every row is accumulated in stored
//     order by one accumulator per output column, so strict mode is bit-identical to
//     cpu_spmm_CSR for every row that is not split into pieces.
//
// item = {row_begin, row_end | ~partial_slot, nnz_begin, nnz_end}.  row_end >= 0: whole
// rows [row_begin, row_end).  row_end < 0: a piece of long row row_begin whose raw sum
// goes to partial[~row_end] for spmm_finalize_kernel.
// ts = tile size in entries (power of two, a multiple of U); dynamic smem per block:
//   GPB * (16 + 2*ts*(sizeof(T)+4)) bytes.
//
// The hot loop is kept lean on purpose (the first version of this kernel spent ~58
// instructions per nonzero on liveness tests, 64-bit index arithmetic and divergence
// bookkeeping and was issue-bound): a batch of U entries takes the CLEAN path -- vector
// LDS of U columns and values, U x (IMAD.WIDE + LDG.128), U multiply-adds -- unless the
// item starts/ends inside it or a row ends inside it.
template <int U> __device__ __forceinline__ void lds_vec(const int *p, int (&c)[U]) {
    if constexpr (U % 4 == 0) {
#pragma unroll
        for (int i = 0; i < U; i += 4) {
            const int4 t = *reinterpret_cast<const int4 *>(p + i);
            c[i] = t.x; c[i + 1] = t.y; c[i + 2] = t.z; c[i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < U; i += 2) {
            const int2 t = *reinterpret_cast<const int2 *>(p + i);
            c[i] = t.x; c[i + 1] = t.y;
        }
    }
}
template <int U> __device__ __forceinline__ void lds_vec(const float *p, float (&a)[U]) {
    if constexpr (U % 4 == 0) {
#pragma unroll
        for (int i = 0; i < U; i += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(p + i);
            a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < U; i += 2) {
            const float2 t = *reinterpret_cast<const float2 *>(p + i);
            a[i] = t.x; a[i + 1] = t.y;
        }
    }
}
template <int U> __device__ __forceinline__ void lds_vec(const double *p, double (&a)[U]) {
#pragma unroll
    for (int i = 0; i < U; i += 2) {
        const double2 t = *reinterpret_cast<const double2 *>(p + i);
        a[i] = t.x; a[i + 1] = t.y;
    }
}

//
// WIN = true is the column-window pass (sx_api.cu: spmm_windows; the reference's K windows,
// src/sextans.cpp:57,337-381, with L2 in the role of the on-chip B buffer): A has been cut
// into windows of W consecutive columns so that the B rows one pass gathers (W * N * s
// bytes) stay L2-resident, and a row's running sum travels from pass to pass through the
// row-major buffer P (leading dimension = C's).  wflags bit 0 (SX_WIN_INIT): the
// accumulator of a row starts from P[row] instead of 0; bit 1 (SX_WIN_RAW): the row's raw
// sum is stored to P[row] instead of the epilogue to C_out.  First window: RAW; middle:
// INIT|RAW; last: INIT.  The chain of additions of a row is the same as in one pass
// (windows ascend, columns ascend inside a window), so strict mode stays bit-identical.
// P[row + 1] is fetched when row opens, like C_in, so its latency hides behind the row.
constexpr int SX_WIN_INIT = 1, SX_WIN_RAW = 2;
// Build-time tuning knobs of the staged kernel (defaults = the measured configuration; an
// alternative library for an A/B run is `make EXTRA_DEFS="-DSX_STAGED_UMAX=16 -DSX_STAGED_MINBLOCKS_F64=2"`):
// most B-row gathers a lane keeps in flight per batch, and the blocks per SM the register
// allocation is held to (fp64 / fp32, one vector per lane).
#ifndef SX_STAGED_UMAX
#define SX_STAGED_UMAX 8
#endif
// (A software-pipelined batch loop -- the gathers of batch q+1 issued before batch q is accumulated,
// 2*U in flight per lane -- was built and measured on a B200: at ~122 registers only two blocks fit
// an SM and it LOST, C5 2.28 ms against 1.59 ms, C4 1.66 against 1.46; more threads with one batch
// each keep more bytes in flight than fewer threads with two.  Removed.)
#ifndef SX_STAGED_MINBLOCKS_F64
#define SX_STAGED_MINBLOCKS_F64 3
#endif
#ifndef SX_STAGED_MINBLOCKS_F32
#define SX_STAGED_MINBLOCKS_F32 4
#endif
// gathers per batch per lane for a lane-group shape (shared with the host's tile sizing)
template <int G, int VPL> struct StagedBatch {
    static constexpr int U = (G < SX_STAGED_UMAX ? G : SX_STAGED_UMAX) / (VPL > 2 ? 4 : VPL);
    static_assert(U == 2 || U == 4 || U == 8 || U == 16, "a batch is 2, 4, 8 or 16 nonzeros");
};
// wflags bit 2 (any instantiation): while a batch's gathers are in flight, ask L2 for the
// B rows of the NEXT batch (its columns already sit in the shared-memory tile), so that
// the next batch's gathers find them in L2 instead of paying the DRAM latency on the
// group's critical path.  No registers are held across the prefetch.
constexpr int SX_FLAG_PREFETCH = 4;
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <typename T, int G, int VPL, bool STRICT, bool WIN = false>
__global__ void __launch_bounds__(256, (VPL > 1 ? 2 : (sizeof(T) == 8 ? SX_STAGED_MINBLOCKS_F64 : SX_STAGED_MINBLOCKS_F32)))
spmm_staged_kernel(const int nitems, const int4 *__restrict__ items, const int ts,
                   const int *__restrict__ rowptr, const int *__restrict__ colidx,
                   const T *__restrict__ val, const T *__restrict__ B, const uint32_t ldbv, const T *Cin,
                   T *Cout, const uint32_t ldcv, T *__restrict__ partial, const uint32_t ldpv,
                   const T alpha, const T beta, const int nvec, T *P, const int wflags) {
    using V = typename VecOf<T>::type;
    const bool w_init = WIN && (wflags & SX_WIN_INIT);
    const bool w_raw = WIN && (wflags & SX_WIN_RAW);
    const bool pf = (wflags & SX_FLAG_PREFETCH) != 0;
    constexpr int U = StagedBatch<G, VPL>::U;  // gathers per batch per lane
    constexpr int GPB = 256 / G;                               // lane groups per block
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [GPB][2] mbarriers | [GPB][2*ts] T values | [GPB][2*ts] int columns
    const int lane = threadIdx.x & 31;
    const int lg = lane & (G - 1);
    const int gi = threadIdx.x / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - lg));
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw) + 2 * gi;
    const T *sval = reinterpret_cast<const T *>(smem_raw + GPB * 16) + (size_t)gi * 2 * ts;
    const int *scol = reinterpret_cast<const int *>(smem_raw + GPB * 16 + (size_t)GPB * 2 * ts * sizeof(T)) + (size_t)gi * 2 * ts;

    const int item = blockIdx.x * GPB + gi;
    if (item >= nitems) return;  // whole groups leave together; no block-wide barrier below
    const uint64_t pol = policy_evict_first();
    const int4 it = __ldg(items + item);
    const int jb = it.z, je = it.w;
    const int jal = jb & ~(U < 4 ? 3 : U - 1);  // batch- and 16-byte-aligned start of both streams
    const int off = jb - jal;     // leading entries that belong to the previous item
    const int len = je - jal;     // stream length counted from the aligned start
    const int ring = 2 * ts - 1;  // entry e lives at smem index e & ring

    auto load_tile = [&](const int k) {  // lane 0 of the group only
        const int e0 = k * ts;
        const uint32_t cnt = (uint32_t)min(ts, (len - e0 + 3) & ~3);
        uint64_t *bk = bar + (k & 1);
        mbar_expect_tx(bk, cnt * (uint32_t)(sizeof(T) + 4));
        tma_bulk_g2s(const_cast<int *>(scol) + (e0 & ring), colidx + jal + e0, cnt * 4u, bk, pol);
        tma_bulk_g2s(const_cast<T *>(sval) + (e0 & ring), val + jal + e0, cnt * (uint32_t)sizeof(T), bk, pol);
    };
    // ts and U are powers of two: tile/batch bookkeeping by shift and mask (as runtime
    // divisions these cost ~70 of the ~250 instructions a batch executed; profiles/r01_colwindow.md)
    const int tsh = 31 - __clz(ts);
    const int nt = (len + ts - 1) >> tsh;
    if (je > jb) {
        if (lg == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            load_tile(0);
            if (nt > 1) load_tile(1);
        }
        __syncwarp(gmask);
    }

    int r = it.x;
    const bool piece = it.y < 0;
    const int re = piece ? r + 1 : it.y;
    int rend = piece ? -1 : __ldg(rowptr + r + 1) - jal;  // in stream coordinates
    // lanes beyond the dense row's last vector (lg >= nvec; one vector per lane only) gather lane 0's piece:
    // their loads are unconditional in the clean path, and must not run past the last row of B
    const V *Bv = reinterpret_cast<const V *>(B) + ((VPL > 1 || lg < nvec) ? lg : 0);
    // a B row's address is base + column * (row bytes): one IMAD.WIDE.U32 per gather
    const unsigned char *Bb = reinterpret_cast<const unsigned char *>(Bv);
    const uint32_t ldbb = ldbv * 16u;
    const V *Cv = reinterpret_cast<const V *>(Cin) + lg;
    V *Pv = reinterpret_cast<V *>(P) + lg;  // dereferenced in window passes only
    V acc[VPL], cin[VPL], pin[VPL];          // pin: next row's running sum (window passes)
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        vzero(acc[v]);
        vzero(cin[v]);
        vzero(pin[v]);
        if (!piece && lg + v * G < nvec) {
            if (!w_raw) cin[v] = ld_once(Cv + (size_t)r * ldcv + v * G, pol);
            if (w_init) {
                acc[v] = ld_once(Pv + (size_t)r * ldcv + v * G, pol);
                if (r + 1 < re) pin[v] = ld_once(Pv + (size_t)(r + 1) * ldcv + v * G, pol);
            }
        }
    }
    auto close_row = [&]() {
        V *cout = reinterpret_cast<V *>(Cout) + (size_t)r * ldcv + lg;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (lg + v * G < nvec) {
                if (w_raw) st_once(Pv + (size_t)r * ldcv + v * G, acc[v], pol);
                else st_once(cout + v * G, vaxpby<STRICT>(alpha, acc[v], beta, cin[v]), pol);
            }
            vzero(acc[v]);
        }
        ++r;
        if (r < re) {
            rend = __ldg(rowptr + r + 1) - jal;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (lg + v * G < nvec) {
                    if (!w_raw) cin[v] = ld_once(Cv + (size_t)r * ldcv + v * G, pol);
                    if (w_init) {
                        acc[v] = pin[v];
                        if (r + 1 < re) pin[v] = ld_once(Pv + (size_t)(r + 1) * ldcv + v * G, pol);
                    }
                }
        } else {
            rend = -1;
        }
    };

    if (je > jb) {
        constexpr int USH = U == 16 ? 4 : (U == 8 ? 3 : (U == 4 ? 2 : 1));
        static_assert((1 << USH) == U, "U is 2, 4, 8 or 16");
        const int bsh = tsh - USH;           // log2(batches per tile)
        const int bmask = (1 << bsh) - 1;
        const int nb = (len + U - 1) / U;
        for (int q = 0; q < nb; ++q) {
            const int e0 = q * U;
            if ((q & bmask) == 0) mbar_wait(bar + ((q >> bsh) & 1), (uint32_t)(((q >> bsh) >> 1) & 1));
            const int idx = e0 & ring;
            int cc[U];
            T av[U];
            V b[U][VPL];
            lds_vec<U>(scol + idx, cc);
            if (e0 >= off && e0 + U <= len) {
                // every entry of the batch belongs to the item
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const V *brow = reinterpret_cast<const V *>(Bb + (uint64_t)(uint32_t)cc[u] * ldbb);
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (VPL == 1 || lg + v * G < nvec) b[u][v] = ldg_vec(brow + v * G);
                        else vzero(b[u][v]);
                    }
                }
                // next batch fully inside the item and inside this tile: its columns are here
                if (pf && ((q + 1) & bmask) != 0 && e0 + 2 * U <= len) {
                    int cn[U];
                    lds_vec<U>(scol + idx + U, cn);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const unsigned char *nrow = Bb + (uint64_t)(uint32_t)cn[u] * ldbb;
#pragma unroll
                        for (int v = 0; v < VPL; ++v)
                            if (VPL == 1 || lg + v * G < nvec) prefetch_l2(nrow + v * G * 16);
                    }
                }
                lds_vec<U>(sval + idx, av);
                if ((unsigned)(rend - e0) >= (unsigned)U) {
                    // clean: no row ends inside the batch
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int v = 0; v < VPL; ++v) vmac<STRICT>(acc[v], av[u], b[u][v]);
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        while (e0 + u == rend) close_row();  // also steps over empty rows
#pragma unroll
                        for (int v = 0; v < VPL; ++v) vmac<STRICT>(acc[v], av[u], b[u][v]);
                    }
                }
            } else {
                // first / last batch of the item: per-entry liveness
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool live = e0 + u >= off && e0 + u < len;
                    const V *brow = reinterpret_cast<const V *>(Bb + (uint64_t)(uint32_t)(live ? cc[u] : 0) * ldbb);
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (live && lg + v * G < nvec) b[u][v] = ldg_vec(brow + v * G);
                        else vzero(b[u][v]);
                    }
                }
                lds_vec<U>(sval + idx, av);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (e0 + u >= off && e0 + u < len) {
                        while (e0 + u == rend) close_row();
#pragma unroll
                        for (int v = 0; v < VPL; ++v) vmac<STRICT>(acc[v], av[u], b[u][v]);
                    }
                }
            }
            if (((q + 1) & bmask) == 0) {
                // the whole group is done with this tile: refill its buffer with tile +2
                const int k = q >> bsh;
                __syncwarp(gmask);
                if (lg == 0 && k + 2 < nt) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    load_tile(k + 2);
                }
            }
        }
    }
    if (piece) {
        V *out = reinterpret_cast<V *>(partial) + (size_t)(~it.y) * ldpv + lg;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (lg + v * G < nvec) out[v * G] = acc[v];
    } else {
        while (r < re) close_row();  // the last row and any trailing empty rows
    }
}

// ---- variant 3: banded matrices, the B window of a row block staged by TMA ----------
// The reference keeps a 4096-row window of B on chip and streams the nonzeros whose
// column falls inside it (src/sextans.h:11, src/sextans.cpp:337-381).  Where a block of 32
// consecutive rows of A only touches a narrow, CONTIGUOUS range of columns -- banded
// matrices: both shipped FEM matrices, span <= 822 / 954 columns -- the same idea fits a
// B200 SM exactly: one thread block owns the 32 rows, thread 0 issues TMA bulk copies of
// (i) rows [cmin, cmin+span) of B, which are one contiguous piece of the row-major image,
// (ii) the block's slice of colidx[] and (iii) of val[], all completing on one mbarrier,
// and the lane groups then walk their rows entirely out of shared memory: no global
// gather, no shuffles, one row per group in stored order (bit-exact in strict mode).
// The dependent chain is block record -> TMA -> shared-memory arithmetic.
// block record = {cmin, span, nnz_begin, nnz_end}; dynamic smem = window + A slice + 16.
//
// PDL = true (experimental, SX_OPT_PDL): the launch carries the programmatic-stream-
// serialization attribute, so this grid may start while the previous kernel of the stream
// is still running.  Everything that only touches A -- block record, row pointers, the TMA
// of the block's colidx/val slice -- happens before griddepcontrol.wait; B and C_in, which
// the previous kernel may have produced, are only touched after it.  launch_dependents is
// issued right after, so that the NEXT kernel's A-side prologue overlaps this kernel's body:
// for the launch-bound SuiteSparse configs a step is a chain of latencies (block record ->
// TMA -> arithmetic), and this takes the A-side part of it off the critical path.
template <typename T, int G, bool STRICT, bool PDL = false>
__global__ void __launch_bounds__(32 * G)
spmm_window_kernel(const int M, const int4 *__restrict__ blocks, const int *__restrict__ rowptr,
                   const int *__restrict__ colidx, const T *__restrict__ val, const T *__restrict__ B,
                   const uint32_t ldbv, const T *Cin, T *Cout, const uint32_t ldcv, const T alpha,
                   const T beta, const int nvec) {
    using V = typename VecOf<T>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const int lg = threadIdx.x & (G - 1);
    const int row = blockIdx.x * 32 + threadIdx.x / G;
    const int4 blk = __ldg(blocks + blockIdx.x);
    const int jal = blk.z & ~3;  // 16-byte aligned start of the A slice
    const uint32_t cnt = (uint32_t)((blk.w - jal + 3) & ~3);
    const uint32_t wbytes = (uint32_t)blk.y * ldbv * 16u;  // the window: span whole rows of B
    const V *win = reinterpret_cast<const V *>(smem_raw);
    const T *sval = reinterpret_cast<const T *>(smem_raw + wbytes);
    const int *scol = reinterpret_cast<const int *>(smem_raw + wbytes + (size_t)cnt * sizeof(T));
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && blk.w > blk.z) {
        const uint64_t pol_a = policy_evict_first();
        mbar_expect_tx(&bar, wbytes + cnt * (uint32_t)(sizeof(T) + 4));
        const unsigned char *src = reinterpret_cast<const unsigned char *>(B) + (size_t)blk.x * ldbv * 16u;
        uint64_t pol_b;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_b));
        if (PDL) {  // A slice first, then wait for the previous kernel, then the B window
            tma_bulk_g2s(const_cast<T *>(sval), val + jal, cnt * (uint32_t)sizeof(T), &bar, pol_a);
            tma_bulk_g2s(const_cast<int *>(scol), colidx + jal, cnt * 4u, &bar, pol_a);
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
        for (uint32_t o = 0; o < wbytes; o += 32768u)  // several copies in flight
            tma_bulk_g2s(smem_raw + o, src + o, min(32768u, wbytes - o), &bar, pol_b);
        if (!PDL) {
            tma_bulk_g2s(const_cast<T *>(sval), val + jal, cnt * (uint32_t)sizeof(T), &bar, pol_a);
            tma_bulk_g2s(const_cast<int *>(scol), colidx + jal, cnt * 4u, &bar, pol_a);
        }
    }
    if (row >= M) {  // after the barrier above; no block-wide barrier below
        if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        return;
    }
    const int begin = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const bool mine = lg < nvec;
    V acc, cin;
    vzero(acc);
    vzero(cin);
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");               // C_in may come from the previous kernel
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next one may start its prologue
    }
    if (mine) cin = reinterpret_cast<const V *>(Cin)[(size_t)row * ldcv + lg];
    if (blk.w > blk.z) mbar_wait(&bar, 0);
    if (mine) {
        const V *w = win + lg;                 // w[(col - cmin) * ldbv] = this lane's piece of B row col
        const int cmin = blk.x;
        int j = begin;
        for (; j + 4 <= end; j += 4) {
            int c[4];
            T a[4];
            V b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { c[u] = scol[j + u - jal]; a[u] = sval[j + u - jal]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) b[u] = w[(uint32_t)(c[u] - cmin) * ldbv];
#pragma unroll
            for (int u = 0; u < 4; ++u) vmac<STRICT>(acc, a[u], b[u]);
        }
        for (; j < end; ++j) vmac<STRICT>(acc, sval[j - jal], w[(uint32_t)(scol[j - jal] - cmin) * ldbv]);
        reinterpret_cast<V *>(Cout)[(size_t)row * ldcv + lg] = vaxpby<STRICT>(alpha, acc, beta, cin);
    }
}

// ---- variant 5: edge lists -- a row block's DISTINCT B rows staged in shared memory, 16-bit local columns ----
// The reference cuts A into column windows, keeps the window of B on chip and stores every
// nonzero as a packed word whose column is LOCAL to the window (col14 | row18 | val32,
// src/sparse_helper.h:419-443; decoded in src/sextans.cpp:398-402), so that a PE indexes its
// on-chip B directly.  Here the "window" of a row block (ROWS consecutive rows, one thread
// block) is the ascending list of the distinct columns its nonzeros touch: exactly those rows
// of B are staged into shared memory, and every nonzero carries a 16-bit index into that
// compacted window (sx_host.cpp: sx_plan_edge_lists).  On FEM-type matrices that is a third of
// the contiguous span variant 3 stages (nasa4704: 142 columns per 32 rows against a span of 456;
// pcrystk02: 317 against 918): a third of the bytes through L2 and of the shared memory, so 4-6
// blocks share an SM instead of 1-2, and the index stream of A shrinks from 4 to 2 bytes per nonzero.
//   block record (two int4): {row_begin, nrows, nnz_begin, nnz_end} {col_begin, ncols, -, smem}
//   shared memory:           window ncols x (G x 16 bytes) | values | local columns | column list | row starts | row ends
//                            (A slice from the 8-entry boundary at or below nnz_begin: whole 16-byte units)
// One lane group per row, stored order, so strict mode is bit-identical to cpu_spmm_CSR.
// (Round 2 also built "super-rows": the 2-3 consecutive rows of a FEM node share their column
// sequence -- nasa4704 1.7 rows per pattern, pcrystk02 2.9 -- so one lane group walked them together
// and read each B piece from shared memory once for all of them.  Bit-exact, and SLOWER on a B200:
// nasa4704 4.40 against 3.68 us, pcrystk02 N=16 9.4 against 7.0 us -- a third of the lane groups
// doing three times the work lengthens the dependent chains more than the saved LDS traffic buys.
// Reverted; the planner and kernel are in the history.)
// 256 threads per block (128 for 2-lane, 512 for 16-lane groups), i.e. ROWS = 64 / 64 / 32 / 32 lane groups for G = 2 / 4 / 8 / 16;
// a lane group takes rows rl, rl + ROWS, ... of its block (blocks are cut by nonzeros, not by rows: a
// small matrix becomes one block per SM with equal work, the reference's equal-length PE lists).
//
// How the window gets on chip was decided by measurement (scripts/micro/launch_floor.cu, B200):
// behind a dependent-launch wait a TMA bulk copy of a cold 16 KB piece costs ~1.05 us per graph
// node, a plain global load ~0.16 us -- and a warp whose lanes hold different copy operands
// issues bulk copies one lane at a time.  So only the A side (values, local columns, the column
// list: contiguous, known before the wait) travels by TMA; the B rows are copied by ALL threads
// with 16-byte cp.async (LDGSTS), a lane group per row, 16*G contiguous bytes per row.
//
// Launch chains: the kernel is written for programmatic dependent launch.  Everything that only
// touches A (records, row pointers, the TMA of the A slice and the column list) and the L2
// PREFETCH of the block's B rows and C_in rows (a hint, consumes nothing) happens before
// griddepcontrol.wait; B and C_in themselves are read after it.  launch_dependents is the first
// instruction: the next kernel of the stream becomes resident beside this one and has its A side
// staged and its B rows on the way to L2 by the time this kernel completes.  Launched without the
// attribute the two instructions do nothing.
//
// HOSTC = true is the host-facing call's kernel (sx_spmm_* with page-locked operands and no
// kernel time asked for): C never exists as a device image.  A block reads the C_in tile of its
// rows straight from the caller's column-major array over PCIe in its prologue -- before the
// dependent-launch wait, i.e. while the kernel that stages B is still running -- and writes the
// result tile back the same way, so C's two PCIe directions overlap each other, B's transfer and
// the arithmetic, and the call is two launches (B staging, this) instead of three.  (Letting the
// blocks take turns on the inbound link in four groups, so that early groups' results leave while
// late groups' C_in arrives, was measured: 74 us against 62 us per call on nasa4704 -- each turn
// pays the PCIe read latency again.  Not kept.)
//
// Multi-GPU, the rank that holds B (npush > 0): the exchange is part of THIS kernel.  After the
// dependent-launch wait every block waits until the peers have finished with the previous
// contents of their images (done[p] >= *pushes), copies its 1/gridDim share of the B image into
// every peer's image with plain 16-byte stores through the NVLink peer mappings (posted writes),
// and goes on with its rows; a one-warp kernel right behind it (publish_push_kernel, a programmatic
// dependent) stores *pushes + 1 into every peer's ready flag -- the kernel boundary is the fence, so
// no block waits for NVLink acknowledgements.  No stream and no collective for the exchange.
// Multi-GPU, a receiving rank (ready != nullptr): B is pushed into this GPU's image by the rank
// that holds it.  *epoch counts the SpMMs this image has served; lane 0 of every warp
// block waits until the local ready flag reaches *epoch + 1 (the push that follows the last SpMM on
// this image) before the window copies are issued, and the last block to finish advances *epoch
// and stores the new value into the pusher's done flag, which lets the pusher overwrite the
// image again.  Counters live in device memory, so a captured launch can be replayed; the
// exchange costs this rank no launch of its own.
constexpr int SX_EDGE_PREFETCH = 1;
// where a pushed B image goes: up to 15 peers' images and ready flags (peer-mapped addresses)
struct PushList { int4 *dst[15]; uint32_t *ready[15]; };
// deferred publication (sx_spmm_fuse_publish): the ready flags of a push that an EARLIER kernel of the stream carried
struct PubList { uint32_t *ready[15]; };
// -DSX_EDGE_TRACE (scripts/build_variant.sh): every block records %globaltimer at its phase
// boundaries into a device array read back by sx_debug_edge_trace -- how a ~3.5 us step splits
// into launch, prologue, dependent-launch wait, window staging and arithmetic.
#ifdef SX_EDGE_TRACE
constexpr int SX_TRACE_ROWS = 1 << 16;
__device__ unsigned long long sx_edge_trace[SX_TRACE_ROWS][8];
__device__ unsigned int sx_edge_trace_count;
__device__ __forceinline__ unsigned long long sx_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SX_TRACE_MARK(i) do { if (threadIdx.x == 0 && trace_row < SX_TRACE_ROWS) sx_edge_trace[trace_row][i] = sx_now(); } while (0)
#else
#define SX_TRACE_MARK(i) do { } while (0)
#endif
// One row of the edge-list kernels out of shared memory: sc[j] / sv[j] = local column / value of entry j of the
// ROW-ALIGNED streams (the row starts at `begin`, a multiple of 8, holds end - begin entries and is padded to a multiple
// of 8 in storage), w[local column * G] = this lane's 16-byte piece of that B row.  A chunk of 8 (column, value) pairs is
// ONE 16-byte load of columns and 8 * sizeof(T) / 16 loads of values -- against 8 + 8 scalar loads when rows start
// anywhere -- next to its 8 loads of B pieces.  Stored order; the columns of chunk k+1 (the critical path: the B pieces'
// addresses) are fetched while chunk k is added up, the B pieces and values four at a time; the additions of pad entries
// are predicated off (never an explicit +0, which would turn a -0 sum into +0).
// Measured on a B200 (profiles/r02_aligned_walk.txt) against the walk over rows that start anywhere (8 + 8 + 8 scalar and
// vector loads per chunk, 71 registers): nasa4704 N=16 fp64 3.51 -> 3.36 us, pcrystk02 N=8/16/32 7.03/6.75/9.13 ->
// 5.94/6.50/8.45 us.  Eight B pieces in flight with a look-ahead of columns AND values took 104 registers (two resident
// blocks instead of three: pcrystk02 N=32 12.1 us, a batch of 20 SpMMs 2.06 instead of 1.49 us each); with a look-ahead
// of columns only 85-87 registers (3.52 us, N=32 12.0); held to 80 registers 3.71 us.
// The walk with scalar loads of the pairs (it does not need the rows aligned, only contiguous): kept for 16-lane groups,
// whose 512-thread blocks are held to 64 registers -- there the eight B pieces in flight it affords beat the vector
// loads' fewer instructions (pcrystk02 N=64 fp32: 18.3 against 19.6 us).
template <typename T, int G, bool STRICT>
__device__ __forceinline__ typename VecOf<T>::type edge_row_walk_scalar(const uint16_t *sc, const T *sv, const typename VecOf<T>::type *w,
                                                                 const int begin, const int end) {
    using V = typename VecOf<T>::type;
    V acc;
    vzero(acc);
    constexpr int UC = 8;
    int j = begin;
    if (j + UC <= end) {
        uint32_t c[UC];
        T a[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) { c[u] = sc[j + u]; a[u] = sv[j + u]; }
        for (;;) {
            V b[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) b[u] = w[c[u] * G];
            const int jn = j + UC;
            const bool more = jn + UC <= end;
            uint32_t c2[UC];
            T a2[UC];
            const int jl = more ? jn : j;  // unconditional loads (this chunk again when there is no next one)
#pragma unroll
            for (int u = 0; u < UC; ++u) { c2[u] = sc[jl + u]; a2[u] = sv[jl + u]; }
#pragma unroll
            for (int u = 0; u < UC; ++u) vmac<STRICT>(acc, a[u], b[u]);
            j = jn;
            if (!more) break;
#pragma unroll
            for (int u = 0; u < UC; ++u) { c[u] = c2[u]; a[u] = a2[u]; }
        }
    }
    if (j < end) {
        // the last, partial chunk: its loads all in flight at once (indices clamped to the row),
        // the additions predicated -- an explicit +0 would turn a -0 sum into +0
        uint32_t c[UC];
        T a[UC];
        V b[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) { const int ju = min(j + u, end - 1); c[u] = sc[ju]; a[u] = sv[ju]; }
#pragma unroll
        for (int u = 0; u < UC; ++u) b[u] = w[c[u] * G];
#pragma unroll
        for (int u = 0; u < UC; ++u)
            if (j + u < end) vmac<STRICT>(acc, a[u], b[u]);
    }
    return acc;
}
// eight 16-bit local columns in four registers; col8(q, u) = column u
__device__ __forceinline__ uint4 lds_cols8(const uint16_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint32_t col8(const uint4 &q, const int u) {
    const uint32_t wd = (u >> 1) == 0 ? q.x : (u >> 1) == 1 ? q.y : (u >> 1) == 2 ? q.z : q.w;
    return (u & 1) ? wd >> 16 : wd & 0xffffu;
}
template <typename T, int G, bool STRICT>
__device__ __forceinline__ typename VecOf<T>::type edge_row_walk(const uint16_t *sc, const T *sv, const typename VecOf<T>::type *w,
                                                                 const int begin, const int end) {
    using V = typename VecOf<T>::type;
    if constexpr (G >= 16) return edge_row_walk_scalar<T, G, STRICT>(sc, sv, w, begin, end);
    V acc;
    vzero(acc);
    constexpr int UC = 8;
    int j = begin;
    if (j >= end) return acc;
    uint4 c = lds_cols8(sc + j);
    for (;;) {
        const int jn = j + UC;
        const bool more = jn < end;
        const uint4 c2 = lds_cols8(sc + (more ? jn : j));  // unconditional load (this chunk again when there is no next one)
#pragma unroll
        for (int h = 0; h < UC; h += 4) {
            if (j + h < end) {
                T a[4];
                V b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = w[col8(c, h + u) * G];
                lds_vec<4>(sv + j + h, a);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j + h + u < end) vmac<STRICT>(acc, a[u], b[u]);
            }
        }
        if (!more) break;
        j = jn;
        c = c2;
    }
    return acc;
}
template <int G> struct EdgeShape {
    // 2-lane groups (32-byte dense rows) in 128-thread blocks: 64 rows per block instead of 128 -- twice the blocks for the
    // same matrix (nasa4704 N=4 fp64 3.54 -> 3.12 us, N=8 fp32 3.49 -> 3.28; pcrystk02 N=8 unchanged at 6.0 us)
    static constexpr int THREADS = G >= 16 ? 512 : (G == 2 ? 128 : 256);
    static constexpr int ROWS = THREADS / G;
};
// blocks per SM the register allocation of the edge-list kernel is held to (build-time knob for A/B runs)
#ifndef SX_EDGE_MINBLOCKS
#define SX_EDGE_MINBLOCKS 2
#endif
template <typename T, int G, bool STRICT, bool HOSTC = false>
__global__ void __launch_bounds__(EdgeShape<G>::THREADS, (G >= 16 ? 2 : SX_EDGE_MINBLOCKS))
spmm_edgelist_kernel(const int4 *__restrict__ blocks, const int *__restrict__ cols, const int *__restrict__ rowptr,
                     const int *__restrict__ prow, const uint16_t *__restrict__ lcol, const T *__restrict__ val, const T *__restrict__ B0,
                     const uint32_t ldbv, const T *Cin0, T *Cout0, const uint32_t ldcv, const T alpha, const T beta,
                     const int nvec, const int flags, const uint32_t *ready, uint32_t *epoch, uint32_t *done_remote,
                     unsigned int *sync_words, const int npush, const PushList push, const int64_t push_n16,
                     const uint32_t *push_done, uint32_t *pushes, T *Ch, const int64_t ldh, const int N,
                     const uint32_t tile_off, const int tile_ld, const int64_t batch_strideB, const int64_t batch_strideC,
                     const PubList pub, const int npub, uint32_t *pub_pushes) {
    using V = typename VecOf<T>::type;
    constexpr int THREADS = EdgeShape<G>::THREADS, ROWS = EdgeShape<G>::ROWS, E = VecOf<T>::E;
    // several B's at once (sx_spmm_device_batch_*): blockIdx.y picks the (B, C_in, C_out) triple; the block's
    // slice of A is the same for every one of them and stays in L2 between the blocks that share it
    const T *__restrict__ B = B0 + (size_t)blockIdx.y * batch_strideB;
    const T *Cin = HOSTC ? Cin0 : Cin0 + (size_t)blockIdx.y * batch_strideC;
    T *Cout = HOSTC ? Cout0 : Cout0 + (size_t)blockIdx.y * batch_strideC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;  // the A slice and the column list
#ifdef SX_EDGE_TRACE
    unsigned int trace_row = SX_TRACE_ROWS;
    if (threadIdx.x == 0) {
        trace_row = atomicAdd(&sx_edge_trace_count, 1u);
        if (trace_row < SX_TRACE_ROWS) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            sx_edge_trace[trace_row][6] = blockIdx.x;
            sx_edge_trace[trace_row][7] = smid;
        }
    }
    SX_TRACE_MARK(0);
#endif
    pdl_launch_dependents();
    const int lg = threadIdx.x & (G - 1);
    const int rl = threadIdx.x / G;
    const int4 b0 = __ldg(blocks + 2 * blockIdx.x), b1 = __ldg(blocks + 2 * blockIdx.x + 1);
    const int row0 = b0.x, nrows = b0.y, jb = b0.z, je = b0.w;
    const int ncols = b1.y;
    const uint32_t rowbytes = ldbv * 16u;                     // a row of the B image in global memory
    const uint32_t wbytes = (uint32_t)ncols * (G * 16u);      // a staged row: G vectors, whatever the image's leading dimension
    const int jal = jb;  // the block's slice of the row-aligned streams: [jb, je) in padded coordinates, multiples of 8
    const bool has = je > jb;
    const uint32_t na = (uint32_t)(je - jb);
    const uint32_t ncp = (uint32_t)(ncols + 3) & ~3u;
    unsigned char *win = smem_raw;
    const T *sval = reinterpret_cast<const T *>(smem_raw + wbytes);
    const uint16_t *scol = reinterpret_cast<const uint16_t *>(smem_raw + wbytes + (size_t)na * sizeof(T));
    const int *scols = reinterpret_cast<const int *>(smem_raw + wbytes + (size_t)na * (sizeof(T) + 2));
    int *srp = const_cast<int *>(scols) + ncp;  // padded starts of the block's rows ...
    int *send = srp + ((nrows + 4) & ~3);       // ... and where their entries end (start + length)
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- A side and hints: before the previous kernel of the stream is known to be complete ----
    if (threadIdx.x == 0 && has) {
        const uint64_t pol_a = policy_evict_first();
        mbar_expect_tx(&bar, na * (uint32_t)(sizeof(T) + 2) + ncp * 4u);
        tma_bulk_g2s(const_cast<int *>(scols), cols + b1.x, ncp * 4u, &bar, pol_a);
        tma_bulk_g2s(const_cast<T *>(sval), val + jal, na * (uint32_t)sizeof(T), &bar, pol_a);
        tma_bulk_g2s(const_cast<uint16_t *>(scol), lcol + jal, na * 2u, &bar, pol_a);
    }
    for (int i = threadIdx.x; i < nrows; i += THREADS) {
        const int ps = __ldg(prow + row0 + i);
        srp[i] = ps;
        send[i] = ps + (__ldg(rowptr + row0 + i + 1) - __ldg(rowptr + row0 + i));
    }
    T *tile = reinterpret_cast<T *>(smem_raw + tile_off);  // HOSTC: tile[column * tile_ld + row of the block]
    if (HOSTC) {
        // the caller's C_in, straight from its page-locked column-major array over PCIe: a warp per
        // column, lanes along the rows (consecutive addresses).  It is the CALLER's data, not the
        // previous kernel's, so it is fetched here, while that kernel (the staging of B) still runs.
        for (int cidx = threadIdx.x >> 5; cidx < N; cidx += THREADS / 32)
            for (int r = threadIdx.x & 31; r < nrows; r += 32) tile[cidx * tile_ld + r] = Ch[(size_t)cidx * ldh + row0 + r];
    }
    const unsigned char *Bb = reinterpret_cast<const unsigned char *>(B) + lg * 16;
    if (has) mbar_wait(&bar, 0);
    if (flags & SX_EDGE_PREFETCH) {
        if (lg * 16 < (int)rowbytes && (lg & 7) == 0)  // one prefetch per 128-byte line of a row
            for (int lr = rl; lr < ncols; lr += ROWS) prefetch_l2(Bb + (size_t)(uint32_t)scols[lr] * rowbytes);
        if (!HOSTC && threadIdx.x == THREADS - 1 && nrows > 0)
            bulk_prefetch_l2(reinterpret_cast<const unsigned char *>(Cin) + (size_t)row0 * ldcv * 16u, (uint32_t)nrows * ldcv * 16u);
    }
    // ---- B and C_in: only after the previous kernel is complete ----
    SX_TRACE_MARK(1);
    pdl_wait();
    SX_TRACE_MARK(2);
    // ---- multi-GPU, deferred publication: the push an earlier kernel of this stream carried is complete (the wait above
    // has returned: everything that kernel stored, the peer stores included, has been performed), so one thread tells
    // the peers -- the job of publish_push_kernel, without a kernel of its own in the chain of dependent launches
    // (a one-warp kernel between two SpMMs costs the chain ~1.1 us per step: its completion gates the next wait).
    if (npub > 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 32) {
        const uint32_t t = *reinterpret_cast<volatile uint32_t *>(pub_pushes);
#pragma unroll
        for (int p = 0; p < 15; ++p)  // constant indices: a by-value list indexed by a variable is copied to local memory at kernel entry
            if (p < npub) st_relaxed_sys(pub.ready[p], t + 1u);
        *reinterpret_cast<volatile uint32_t *>(pub_pushes) = t + 1u;
    }
    // ---- multi-GPU: one thread polls.  No system-scope FENCE anywhere in this kernel: on sm_100a
    // fence.acq_rel.sys / st.release.sys are MEMBAR.ALL.SYS, measured at 4-5 us per block with peer
    // mappings live (profiles/r02_exchange_trace_n2.txt), more than the whole SpMM.  The flag is
    // polled with relaxed loads and read once more with ld.acquire.sys (LDG.STRONG.SYS + an L1
    // invalidate; the B rows are then fetched with cp.async.cg, i.e. from L2, where the pushed data is).
    uint32_t step = 0;
    if (ready != nullptr || npush > 0) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            if (ready != nullptr) {  // a receiving rank: the pushed B image of this step has landed
                step = *reinterpret_cast<volatile uint32_t *>(epoch) + 1u;  // advanced only after every block is done
                while ((int)(ld_relaxed_sys(ready) - step) < 0) {
                    __nanosleep(20);
                    if (clock64() - t0 > 4000000000ll) { atomicExch(sync_words + 1, 1u); break; }  // ~2 s: never hang the GPU on a lost peer
                }
                (void)ld_acquire_sys(ready);
            }
            if (npush > 0) {         // the rank that holds B: the peers are done READING the previous contents of their
                                     // images (nothing of theirs is read here, so relaxed loads are all it takes)
                const uint32_t pushed = *reinterpret_cast<volatile uint32_t *>(pushes);  // advanced by publish_push_kernel
                for (int p = 0; p < npush; ++p)
                    while ((int)(ld_relaxed_sys(push_done + p) - pushed) < 0) {
                        __nanosleep(20);
                        if (clock64() - t0 > 4000000000ll) { atomicExch(sync_words + 1, 1u); break; }
                    }
            }
        }
        __syncthreads();
    }
    SX_TRACE_MARK(4);  // multi-GPU: the step flag has been seen (else: right after the wait)
    if (lg < nvec)
        for (int lr = rl; lr < ncols; lr += ROWS)
            cp_async_16(win + ((size_t)lr * G + lg) * 16, Bb + (size_t)(uint32_t)scols[lr] * rowbytes);
    if (npush > 0) {  // this block's share of the B image goes to every peer: posted 16-byte stores over NVLink
        const int4 *src = reinterpret_cast<const int4 *>(B);
        const int64_t lo = push_n16 * blockIdx.x / gridDim.x, hi = push_n16 * (blockIdx.x + 1) / gridDim.x;
        for (int64_t i = lo + threadIdx.x; i < hi; i += THREADS) {
            const int4 v = __ldg(src + i);
#pragma unroll
            for (int p = 0; p < 15; ++p)
                if (p < npush) push.dst[p][i] = v;
        }
    }
    // a lane group takes rows rl, rl + ROWS, ... of the block; C_in of the next one is fetched while this one is computed
    const bool lane_on = lg < nvec;
    const V *Cv = reinterpret_cast<const V *>(Cin) + (size_t)row0 * ldcv + lg;
    V *Ov = reinterpret_cast<V *>(Cout) + (size_t)row0 * ldcv + lg;
    V cin_next;
    vzero(cin_next);
    if (!HOSTC && lane_on && rl < nrows) cin_next = Cv[(size_t)rl * ldcv];
    cp_async_wait_all();
    __syncthreads();
    SX_TRACE_MARK(3);
    const T *sv = sval - jal;  // sv[j] = value of nonzero j
    const uint16_t *sc = scol - jal;
    const V *w = reinterpret_cast<const V *>(win) + lg;  // w[local column * G] = this lane's piece of that B row
    if (lane_on)
        for (int rr = rl; rr < nrows; rr += ROWS) {
            const int begin = srp[rr], end = send[rr];
            V cin = cin_next;
            if (HOSTC) {
                T *cp = reinterpret_cast<T *>(&cin);
#pragma unroll
                for (int e = 0; e < E; ++e) cp[e] = (lg * E + e < N) ? tile[(lg * E + e) * tile_ld + rr] : T(0);
            } else if (rr + ROWS < nrows) {
                cin_next = Cv[(size_t)(rr + ROWS) * ldcv];
            }
            const V acc = edge_row_walk<T, G, STRICT>(sc, sv, w, begin, end);
            const V out = vaxpby<STRICT>(alpha, acc, beta, cin);
            if (HOSTC) {
                const T *op = reinterpret_cast<const T *>(&out);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (lg * E + e < N) tile[(lg * E + e) * tile_ld + rr] = op[e];
            } else {
                Ov[(size_t)rr * ldcv] = out;
            }
        }
    if (HOSTC) {  // the result tile back into the caller's array, a warp per column again
        __syncthreads();
        for (int cidx = threadIdx.x >> 5; cidx < N; cidx += THREADS / 32)
            for (int r = threadIdx.x & 31; r < nrows; r += 32) Ch[(size_t)cidx * ldh + row0 + r] = tile[cidx * tile_ld + r];
    }
#ifdef SX_EDGE_TRACE
    __syncthreads();
    SX_TRACE_MARK(5);      // every row of the block done
#endif
    if (ready != nullptr) {  // a receiving rank: tell the pusher that this rank is done with the image
        __syncthreads();     // (its reads of the image have all returned: a plain store is enough)
        if (threadIdx.x == 0) {
            if (atomicAdd(sync_words + 2, 1u) == gridDim.x - 1) {
                sync_words[2] = 0;
                *reinterpret_cast<volatile uint32_t *>(epoch) = step;
                st_relaxed_sys(done_remote, step);
            }
        }
    }
    // (the rank that holds B publishes the step from publish_push_kernel, launched right behind this
    // kernel: the kernel boundary is the fence, and this kernel's tail stays free of system-scope fences)
}

// (A PERSISTENT form of the batched launch -- a block keeps its row block for a slice of the operands, the A side staged
// once, the window of operand k+1 copied by cp.async while the rows of operand k are walked -- was built, emulated and run:
// bit-exact, and no faster than grid.y = nb blocks of spmm_edgelist_kernel on nasa4704 (1.49 us per SpMM either way) and
// slower on pcrystk02 (N=16: 5.79 against 3.78 us; its second window buffer costs a resident block per SM).  What bounds a
// batch is the walk itself -- shared-memory instructions of three co-resident blocks, ncu: L1 58 % busy -- not the
// staging latency.  Removed; commit d557125 holds it, profiles/r02_batch_persistent.txt the numbers.)

// ---- the host-facing call as ONE kernel (sx_spmm_* with small page-locked operands, kernel_ns == NULL) ----
// The pieces of the call were measured on the host's clock (scripts/micro/pcie_floor.cu, profiles/r02_pcie_floor.txt):
// an empty launch + cudaStreamSynchronize costs 10 us, every further launch ~3.7 us, reading B and C_in of
// nasa4704 N=16 fp64 (1.2 MB) over PCIe 25 us, writing C (0.6 MB) 14 us -- one after the other 52 us -- and both
// directions AT ONCE 43 us: the link is full duplex, the two-launch call (staging kernel, then the HOSTC form of
// spmm_edgelist_kernel: 60 us) only ever uses one direction at a time.  So:
//   * ONE launch.  Every block first fetches a 1/gridDim share of the rows of the caller's column-major B and the
//     C_in tile of its own rows into shared memory with cp.async straight from host memory (LDGSTS over PCIe;
//     nothing waits on them yet), one commit group per COLUMN GROUP of gw columns, issued in group order.
//   * Column groups are independent SpMMs (C[:, n] depends on B[:, n] only).  For group g = 0, 1, ...: wait for
//     that group's copies, write the B share into the row-major device image, fence, count the block in on
//     counters[g]; when all blocks are in, fetch the group's slice of the window rows from L2 (ld.global.cg -- the
//     image was written by other SMs), walk the rows (same order of operations as everywhere: bit-identical to
//     cpu_spmm_CSR), and store the result columns into the caller's array -- posted PCIe writes that leave
//     while the later groups' operands are still arriving.
// Measured (scripts/micro/e2e_c.cpp, nasa4704 N=16 fp64, median per call on the host's clock; profiles/r02_e2e_call.txt):
// two launches 58.9 us; this kernel with 1 / 2 / 4 / 8 groups 58.6 / 56.5 / 61.8 / 77.6 us (one group requested ahead;
// two or three ahead: 58.8-59.1 / 59.3-60.3 / 75.8-76.6) -- every group costs a grid-wide wait (~2 us), and the floor of
// this box for "1.2 MB in, then 0.6 MB out" in ONE launch is 9.6 + 25 + 14 = 49 us before any arithmetic.  Default: 2 groups.
// The grid must be co-resident (blocks wait for one another): the host checks the occupancy and falls back to the
// two-launch path otherwise; a wait that sees nothing for ~2 s gives up and raises *timeout_flag (host memory).
// counters[0..7]: every block adds 1 to each of them exactly once per launch, so they all stand at `target -
// gridDim.x` when a launch starts, whatever the number of groups of the launches before it.
constexpr int SX_HOST_MAX_GROUPS = 8;

template <typename T, int G, bool STRICT>
__global__ void __launch_bounds__(EdgeShape<G>::THREADS, 2)
spmm_edgelist_host_kernel(const int4 *__restrict__ blocks, const int *__restrict__ cols, const int *__restrict__ rowptr,
                          const int *__restrict__ prow, const uint16_t *__restrict__ lcol, const T *__restrict__ val, const T *Bh, T *Bimg,
                          const uint32_t ldbv, T *Ch, const int64_t M, const int64_t K, const int N, const T alpha,
                          const T beta, const int gw, uint32_t *counters, const uint32_t target, uint32_t *timeout_flag,
                          const uint32_t tile_off, const int tile_ld, const uint32_t share_off, const int share_ld, const int depth) {
    using V = typename VecOf<T>::type;
    constexpr int THREADS = EdgeShape<G>::THREADS, ROWS = EdgeShape<G>::ROWS, E = VecOf<T>::E, NWARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const int lg = threadIdx.x & (G - 1);
    const int rl = threadIdx.x / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int4 b0 = __ldg(blocks + 2 * blockIdx.x), b1 = __ldg(blocks + 2 * blockIdx.x + 1);
    const int row0 = b0.x, nrows = b0.y, jb = b0.z, je = b0.w;
    const int ncols = b1.y;
    const uint32_t wbytes = (uint32_t)ncols * (G * 16u);
    const int jal = jb;
    const bool has = je > jb;
    const uint32_t na = (uint32_t)(je - jb);
    const uint32_t ncp = (uint32_t)(ncols + 3) & ~3u;
    V *win = reinterpret_cast<V *>(smem_raw);
    const T *sval = reinterpret_cast<const T *>(smem_raw + wbytes);
    const uint16_t *scol = reinterpret_cast<const uint16_t *>(smem_raw + wbytes + (size_t)na * sizeof(T));
    const int *scols = reinterpret_cast<const int *>(smem_raw + wbytes + (size_t)na * (sizeof(T) + 2));
    int *srp = const_cast<int *>(scols) + ncp;
    int *send = srp + ((nrows + 4) & ~3);
    T *tile = reinterpret_cast<T *>(smem_raw + tile_off);    // tile[column * tile_ld + row of the block]
    T *share = reinterpret_cast<T *>(smem_raw + share_off);  // share[column * share_ld + row of the share]
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && has) {  // the A side by TMA, as in spmm_edgelist_kernel
        const uint64_t pol_a = policy_evict_first();
        mbar_expect_tx(&bar, na * (uint32_t)(sizeof(T) + 2) + ncp * 4u);
        tma_bulk_g2s(const_cast<int *>(scols), cols + b1.x, ncp * 4u, &bar, pol_a);
        tma_bulk_g2s(const_cast<T *>(sval), val + jal, na * (uint32_t)sizeof(T), &bar, pol_a);
        tma_bulk_g2s(const_cast<uint16_t *>(scol), lcol + jal, na * 2u, &bar, pol_a);
    }
    // ---- inbound: this block's share of B's rows and its C_in tile, group by group, nothing waited for yet ----
    const int ks0 = (int)(K * (int64_t)blockIdx.x / gridDim.x), ks1 = (int)(K * (int64_t)(blockIdx.x + 1) / gridDim.x);
    const int nshare = ks1 - ks0;
    const int ngroups = (N + gw - 1) / gw;
    // `depth` groups in flight: with everything requested at once the link serves all groups side by side and the
    // first one is complete no earlier than the last; two in flight keep the link busy and the groups in order
    auto fetch_group = [&](const int g) {
        const int n_lo = g * gw, n_hi = min(N, n_lo + gw);
        for (int cidx = n_lo + warp; cidx < n_hi; cidx += NWARPS) {  // a warp per column, lanes along the rows
            for (int r = lane; r < nshare; r += 32) cp_async_elem(share + cidx * share_ld + r, Bh + (size_t)cidx * K + ks0 + r);
            for (int r = lane; r < nrows; r += 32) cp_async_elem(tile + cidx * tile_ld + r, Ch + (size_t)cidx * M + row0 + r);
        }
        cp_async_commit();
    };
    for (int g = 0; g < min(depth, ngroups); ++g) fetch_group(g);
    for (int i = threadIdx.x; i < nrows; i += THREADS) {
        const int ps = __ldg(prow + row0 + i);
        srp[i] = ps;
        send[i] = ps + (__ldg(rowptr + row0 + i + 1) - __ldg(rowptr + row0 + i));
    }
    if (has) mbar_wait(&bar, 0);
    const T *sv = sval - jal;
    const uint16_t *sc = scol - jal;
    V *Bv = reinterpret_cast<V *>(Bimg);
    bool gave_up = false;
    for (int g = 0; g < ngroups; ++g) {
        const int n_lo = g * gw, n_hi = min(N, n_lo + gw);
        const int v_lo = n_lo / E, nvg = ((g == ngroups - 1 ? (int)ldbv * E : n_hi) + E - 1) / E - v_lo;  // the last group zero-fills the padding
        cp_async_wait_pending(min(depth, ngroups - g) - 1);
        __syncthreads();
        if (g + depth < ngroups) fetch_group(g + depth);  // (its tile and share slots are its own: no hazard with the groups in progress)
        // the share of this group's columns of B into the row-major device image, 16 bytes per store
        for (int i = threadIdx.x; i < nshare * nvg; i += THREADS) {
            const int r = i / nvg, v = v_lo + i % nvg;
            V x;
            T *xp = reinterpret_cast<T *>(&x);
#pragma unroll
            for (int e = 0; e < E; ++e) xp[e] = (v * E + e < N) ? share[(v * E + e) * share_ld + r] : T(0);
            Bv[(size_t)(ks0 + r) * ldbv + v] = x;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(counters + g, 1u);
            const long long t0 = clock64();
            while ((int)(ld_acquire_gpu(counters + g) - target) < 0) {
                __nanosleep(20);
                if (clock64() - t0 > 4000000000ll) { *reinterpret_cast<volatile uint32_t *>(timeout_flag) = 1u; break; }
            }
        }
        __syncthreads();
        // this group's slice of the window rows, from L2
        const int nvw = min(nvg, G - v_lo);
        for (int i = threadIdx.x; i < ncols * nvw; i += THREADS) {
            const int lr = i / nvw, v = v_lo + i % nvw;
            win[lr * G + v] = ld_l2(Bv + (size_t)(uint32_t)scols[lr] * ldbv + v);
        }
        __syncthreads();
        if (lg >= v_lo && lg < v_lo + nvw && lg * E < N)
            for (int rr = rl; rr < nrows; rr += ROWS) {
                V cin;
                T *cp = reinterpret_cast<T *>(&cin);
#pragma unroll
                for (int e = 0; e < E; ++e) cp[e] = (lg * E + e < N) ? tile[(lg * E + e) * tile_ld + rr] : T(0);
                const V acc = edge_row_walk<T, G, STRICT>(sc, sv, win + lg, srp[rr], send[rr]);
                const V out = vaxpby<STRICT>(alpha, acc, beta, cin);
                const T *op = reinterpret_cast<const T *>(&out);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (lg * E + e < N) tile[(lg * E + e) * tile_ld + rr] = op[e];
            }
        __syncthreads();
        for (int cidx = n_lo + warp; cidx < n_hi; cidx += NWARPS)  // the result columns into the caller's array: posted writes
            for (int r = lane; r < nrows; r += 32) Ch[(size_t)cidx * M + row0 + r] = tile[cidx * tile_ld + r];
    }
    (void)gave_up;
    if (threadIdx.x == 0)
        for (int g = ngroups; g < SX_HOST_MAX_GROUPS; ++g) atomicAdd(counters + g, 1u);  // keep the eight counters level
}

// ---- variant 4 (experimental, SX_OPT_SLIDE): long banded matrices, a SLIDING B window -------
// Variant 3 gives every 32-row block its own copy of its B window; on a long banded matrix
// consecutive blocks' windows are almost the same rows, so each SM keeps re-fetching ~100 KB
// from L2 to do 32 rows of work, one block per SM, load and compute in turns (FEM-like
// band=100, M=1e6, fp64: 1.11 ms against 0.22 ms at the HBM roof).  Here a thread block walks
// a CHAIN of consecutive 32-row steps and keeps B in a shared-memory RING indexed by
// (row & rmask): step s only brings in the rows above the highest row loaded so far
// (~32 new rows instead of the whole window), together with its slice of colidx/val into one
// of two A buffers, all completing on that buffer's mbarrier; the loads of step s+2 are
// issued as soon as step s is done, so they fly while step s+1 computes.  This is the
// reference's scheme in its native form -- A streamed once through a small buffer, B held
// on chip while the rows that use it go by (src/sextans.cpp:337-420) -- with the window
// following the band instead of standing still.  The host plan (sx_host.cpp: sx_plan_slide)
// sizes the ring so that a step's columns stay resident while the next step's rows arrive.
//   chain = {first step, last step + 1};  step = {load_lo, load_hi, nnz_begin, nnz_end}
//   dynamic smem: ring (rmask+1) * ldbv * 16 | 2 * abuf values | 2 * abuf columns
// One row per lane group, stored order, so strict mode is bit-identical to cpu_spmm_CSR.
template <typename T, int G, bool STRICT>
__global__ void __launch_bounds__(32 * G)
spmm_slide_kernel(const int M, const int2 *__restrict__ chains, const int4 *__restrict__ steps,
                  const int *__restrict__ rowptr, const int *__restrict__ colidx, const T *__restrict__ val,
                  const T *__restrict__ B, const uint32_t ldbv, const T *Cin, T *Cout, const uint32_t ldcv,
                  const T alpha, const T beta, const int nvec, const uint32_t rmask, const uint32_t abuf) {
    using V = typename VecOf<T>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[2];
    const int lg = threadIdx.x & (G - 1);
    const int rl = threadIdx.x / G;
    const uint32_t rowbytes = ldbv * 16u;
    const V *ring = reinterpret_cast<const V *>(smem_raw);
    unsigned char *abase = smem_raw + (size_t)(rmask + 1) * rowbytes;
    T *svals = reinterpret_cast<T *>(abase);                                            // [2][abuf]
    int *scols = reinterpret_cast<int *>(abase + (size_t)2 * abuf * sizeof(T));           // [2][abuf]
    const int2 ch = __ldg(chains + blockIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: everything step s needs, into A buffer `slot` and the ring, on full[slot]
    auto issue = [&](const int s, const int slot) {
        const int4 st = __ldg(steps + s);
        const int jal = st.z & ~3;
        const uint32_t cnt = st.w > st.z ? (uint32_t)((st.w - jal + 3) & ~3) : 0u;
        const uint32_t nrows = (uint32_t)(st.y - st.x);
        const uint64_t pol_a = policy_evict_first();
        uint64_t pol_b;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_b));
        mbar_expect_tx(&full[slot], nrows * rowbytes + cnt * (uint32_t)(sizeof(T) + 4));
        const uint32_t chunk = 32768u / rowbytes;  // rows per bulk copy
        for (uint32_t r = (uint32_t)st.x; r < (uint32_t)st.y;) {
            const uint32_t at = r & rmask;
            const uint32_t n = min(min((uint32_t)st.y - r, rmask + 1 - at), chunk);  // up to the ring's wrap
            tma_bulk_g2s(smem_raw + (size_t)at * rowbytes, reinterpret_cast<const unsigned char *>(B) + (size_t)r * rowbytes,
                         n * rowbytes, &full[slot], pol_b);
            r += n;
        }
        if (cnt) {
            tma_bulk_g2s(svals + (size_t)slot * abuf, val + jal, cnt * (uint32_t)sizeof(T), &full[slot], pol_a);
            tma_bulk_g2s(scols + (size_t)slot * abuf, colidx + jal, cnt * 4u, &full[slot], pol_a);
        }
    };
    if (threadIdx.x == 0) {
        issue(ch.x, 0);
        if (ch.x + 1 < ch.y) issue(ch.x + 1, 1);
    }
    // the row pointers and the A-slice origin of a step are fetched one step ahead, so that no
    // global-memory latency sits between the arrival of a step's data and its arithmetic
    int nbegin = 0, nend = 0, njal = 0;
    {
        const int row0 = ch.x * 32 + rl;
        if (ch.x < ch.y) {
            njal = __ldg(&steps[ch.x].z) & ~3;
            if (row0 < M && lg < nvec) { nbegin = __ldg(rowptr + row0); nend = __ldg(rowptr + row0 + 1); }
        }
    }
    for (int s = ch.x, i = 0; s < ch.y; ++s, ++i) {
        const int slot = i & 1;
        const int row = s * 32 + rl;
        const bool mine = row < M && lg < nvec;
        const int begin = nbegin, end = nend, jal = njal;
        V acc, cin;
        vzero(acc);
        vzero(cin);
        if (mine) cin = reinterpret_cast<const V *>(Cin)[(size_t)row * ldcv + lg];
        if (s + 1 < ch.y) {
            const int nrow = row + 32;
            njal = __ldg(&steps[s + 1].z) & ~3;
            nbegin = nend = 0;
            if (nrow < M && lg < nvec) { nbegin = __ldg(rowptr + nrow); nend = __ldg(rowptr + nrow + 1); }
        }
        mbar_wait(&full[slot], (uint32_t)((i >> 1) & 1));
        if (mine) {
            const T *sv = svals + (size_t)slot * abuf - jal;  // sv[j] = value of nonzero j
            const int *sc = scols + (size_t)slot * abuf - jal;
            const V *w = ring + lg;
            // chunks of 8 nonzeros, software-pipelined: the (col, val) pairs of chunk k+1 and the
            // eight B-row pieces of chunk k are in flight while the strictly ordered chain of
            // additions of chunk k runs -- with one block of 8 warps per SM there is little else
            // to hide shared-memory latency behind
            constexpr int UC = 8;
            int j = begin;
            if (j + UC <= end) {
                int c[UC];
                T a[UC];
#pragma unroll
                for (int u = 0; u < UC; ++u) { c[u] = sc[j + u]; a[u] = sv[j + u]; }
                for (;;) {
                    V b[UC];
#pragma unroll
                    for (int u = 0; u < UC; ++u) b[u] = w[((uint32_t)c[u] & rmask) * ldbv];
                    const int jn = j + UC;
                    const bool more = jn + UC <= end;
                    int c2[UC];
                    T a2[UC];
                    const int jl = more ? jn : j;  // unconditional loads (re-reading this chunk when there is no next one)
#pragma unroll
                    for (int u = 0; u < UC; ++u) { c2[u] = sc[jl + u]; a2[u] = sv[jl + u]; }
#pragma unroll
                    for (int u = 0; u < UC; ++u) vmac<STRICT>(acc, a[u], b[u]);
                    j = jn;
                    if (!more) break;
#pragma unroll
                    for (int u = 0; u < UC; ++u) { c[u] = c2[u]; a[u] = a2[u]; }
                }
            }
            for (; j < end; ++j) vmac<STRICT>(acc, sv[j], w[((uint32_t)sc[j] & rmask) * ldbv]);
            reinterpret_cast<V *>(Cout)[(size_t)row * ldcv + lg] = vaxpby<STRICT>(alpha, acc, beta, cin);
        }
        __syncthreads();  // everybody is done with A buffer `slot` and with the ring rows step s+2 will replace
        if (threadIdx.x == 0 && s + 2 < ch.y) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(s + 2, slot);
        }
    }
}

// ---- one row group per row (variant 1: matrices that fill less than one wave) ---
// Latency-oriented walk of one row for G <= 8.  Entries travel in chunks of 8 (each
// lane holds 8/G (col,val) pairs of a chunk), one chunk is one batch of 8 B-row gathers;
// the gathers of chunk k+1 are issued before chunk k is accumulated and chunks are
// fetched two ahead, so a short row costs rowptr -> entries -> B rows instead of one
// B-row latency per chunk.
template <typename T, int G, int VPL, bool STRICT>
__device__ __forceinline__ void accumulate_row_pipelined(
    typename VecOf<T>::type (&acc)[VPL], const int begin, const int end, const int lg,
    const unsigned gmask, const int nvec, const int *__restrict__ colidx, const T *__restrict__ val,
    const T *__restrict__ B, const int64_t ldb) {
    using V = typename VecOf<T>::type;
    constexpr int CH = 8;      // entries per chunk
    constexpr int R = CH / G;  // (col,val) pairs per lane per chunk
    struct Chunk { int c[R]; T a[R]; };
    auto fetch = [&](const int b, Chunk &k) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = b + r * G + lg;
            k.c[r] = 0;
            k.a[r] = T(0);
            if (p < end) { k.c[r] = __ldg(colidx + p); k.a[r] = __ldg(val + p); }
        }
    };
    auto gather = [&](V (&b)[CH][VPL], const Chunk &k, const int base) {
#pragma unroll
        for (int t = 0; t < CH; ++t) {
            const int cc = __shfl_sync(gmask, k.c[t / G], t % G, G);
            const V *brow = reinterpret_cast<const V *>(B + (int64_t)cc * ldb);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int vi = lg + v * G;
                if (base + t < end && vi < nvec) b[t][v] = ldg_vec(brow + vi);
                else vzero(b[t][v]);
            }
        }
    };
    auto mac_chunk = [&](V (&b)[CH][VPL], const Chunk &k, const int base) {
#pragma unroll
        for (int t = 0; t < CH; ++t) {
            const T av = __shfl_sync(gmask, k.a[t / G], t % G, G);
            if (base + t < end) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) vmac<STRICT>(acc[v], av, b[t][v]);
            }
        }
    };
    if (begin >= end) return;
    Chunk k0, k1, k2;
    fetch(begin, k0);
    fetch(begin + CH, k1);
    fetch(begin + 2 * CH, k2);
    V bA[CH][VPL], bB[CH][VPL];
    gather(bA, k0, begin);
    for (int base = begin; base < end; base += 2 * CH) {
        Chunk k3, k4;
        if (base + CH < end) gather(bB, k1, base + CH);
        fetch(base + 3 * CH, k3);
        mac_chunk(bA, k0, base);
        if (base + CH >= end) break;
        if (base + 2 * CH < end) gather(bA, k2, base + 2 * CH);
        fetch(base + 4 * CH, k4);
        mac_chunk(bB, k1, base + CH);
        k0 = k2;
        k1 = k3;
        k2 = k4;
    }
}

// Rows longer than split_nnz (when > 0) are left to the segment kernels below.
template <typename T, int G, int VPL, bool STRICT>
__global__ void __launch_bounds__(256)
spmm_rows_kernel(const int M, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                 const T *__restrict__ val, const T *__restrict__ B, const int64_t ldb,
                 const T *Cin, T *Cout, const int64_t ldc, const T alpha, const T beta,
                 const int nvec, const int split_nnz) {
    using V = typename VecOf<T>::type;
    const int lane = threadIdx.x & 31;
    const int lg = lane & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - lg));
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (row >= M) return;
    const int begin = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    if (split_nnz > 0 && end - begin > split_nnz) return;

    // C_in does not depend on A: fetch it now, use it in the epilogue
    const V *cin = reinterpret_cast<const V *>(Cin + row * ldc);
    V acc[VPL], cv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        vzero(acc[v]);
        vzero(cv[v]);
        if (lg + v * G < nvec) cv[v] = cin[lg + v * G];
    }
    if constexpr (G <= 8 && VPL == 1)
        accumulate_row_pipelined<T, G, VPL, STRICT>(acc, begin, end, lg, gmask, nvec, colidx, val, B, ldb);
    else
        accumulate_range<T, G, VPL, STRICT>(acc, begin, end, 0, 1, lg, gmask, nvec, colidx, val, B, ldb);

    V *cout = reinterpret_cast<V *>(Cout + row * ldc);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int vi = lg + v * G;
        if (vi < nvec) cout[vi] = vaxpby<STRICT>(alpha, acc[v], beta, cv[v]);
    }
}

// ---- long rows: one warp per segment of <= split_nnz nonzeros -------------------
// The 32/G groups of the warp take interleaved chunks of the segment and their
// partial vectors are combined by a fixed xor-shuffle tree, so the result is
// deterministic but NOT in the oracle's summation order.
template <typename T, int G, int VPL, bool STRICT>
__global__ void __launch_bounds__(256)
spmm_segments_kernel(const int nseg, const int *__restrict__ seg_begin,
                     const int *__restrict__ seg_end, const int *__restrict__ colidx,
                     const T *__restrict__ val, const T *__restrict__ B, const int64_t ldb,
                     T *__restrict__ partial, const int64_t ldp, const int nvec) {
    using V = typename VecOf<T>::type;
    const int lane = threadIdx.x & 31;
    const int lg = lane & (G - 1);
    const int gi = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - lg));
    const int seg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (seg >= nseg) return;  // warp-uniform
    V acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) vzero(acc[v]);
    accumulate_range<T, G, VPL, STRICT>(acc, __ldg(seg_begin + seg), __ldg(seg_end + seg), gi,
                                        32 / G, lg, gmask, nvec, colidx, val, B, ldb);
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            V o = vshfl_xor(0xffffffffu, acc[v], off);
            vadd(acc[v], o);
        }
    }
    if (gi == 0) {
        V *out = reinterpret_cast<V *>(partial + (int64_t)seg * ldp);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int vi = lg + v * G;
            if (vi < nvec) out[vi] = acc[v];
        }
    }
}

// Sum a split row's segment partials in segment order and apply the epilogue.
// WIN: the column-window pass of a split row -- the pieces are added to the running sum
// in P (SX_WIN_INIT) and the result goes back to P (SX_WIN_RAW) or through the epilogue.
template <typename T, int G, int VPL, bool STRICT, bool WIN = false>
__global__ void __launch_bounds__(256)
spmm_finalize_kernel(const int nsplit, const int *__restrict__ split_row,
                     const int *__restrict__ split_seg_ptr, const T *__restrict__ partial,
                     const int64_t ldp, const T *Cin, T *Cout, const int64_t ldc, const T alpha,
                     const T beta, const int nvec, T *P, const int wflags) {
    using V = typename VecOf<T>::type;
    const int lg = threadIdx.x & (G - 1);
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (i >= nsplit) return;
    const int64_t row = __ldg(split_row + i);
    const int s0 = __ldg(split_seg_ptr + i), s1 = __ldg(split_seg_ptr + i + 1);
    const V *cin = reinterpret_cast<const V *>(Cin + row * ldc);
    V *cout = reinterpret_cast<V *>(Cout + row * ldc);
    V *prow = reinterpret_cast<V *>(P + row * ldc);  // dereferenced in window passes only
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int vi = lg + v * G;
        if (vi >= nvec) continue;
        V acc;
        vzero(acc);
        if (WIN && (wflags & SX_WIN_INIT)) acc = prow[vi];
        for (int s = s0; s < s1; ++s) {
            const V p = reinterpret_cast<const V *>(partial + (int64_t)s * ldp)[vi];
            vadd(acc, p);
        }
        if (WIN && (wflags & SX_WIN_RAW)) prow[vi] = acc;
        else cout[vi] = vaxpby<STRICT>(alpha, acc, beta, cin[vi]);
    }
}

// ---- dense-tile variant: 8-row panels on the FP64 tensor cores (DMMA) ---------------
// Where A admits dense sub-blocks (FEM-type matrices: the rows of a node share their
// columns) the upload splits A into  A = A_tiles + A_rest  (sx_api.cu: build_panels):
// in every panel of 8 consecutive rows, a column used by at least `tau` of the 8 rows
// becomes a column of the panel's dense 8 x w tile (missing entries are explicit zeros),
// everything else stays in a CSR remainder that the staged kernel adds afterwards.
// One warp owns one panel and walks its tile four columns at a time with
//     mma.sync.aligned.m8n8k4.row.col.f64  (SASS DMMA.884)
// -- tcgen05.mma has no fp64 kind, so the legacy warp-level path IS the fp64 tensor path
// on sm_100a.  The point of the format is not the flops but the gathers: one B row is
// fetched once per panel column instead of once per nonzero, i.e. up to 8x fewer bytes
// through L2, which is what bounds SpMM on this machine.
// Storage per k-step (4 tile columns): 4 column indices + 32 values in A-fragment order
// (lane l holds A[l/4][l%4], explicit zeros included), so a step is one coalesced
// 256-byte load with no index arithmetic.  (A packed variant -- occupancy mask + only the
// stored values, found by popc -- halves the bytes of A at fill 0.5 but doubles the
// instructions per step and ran 1.9x SLOWER: the kernel is latency/issue-bound, not
// DRAM-bound; profiles/r01_pass16_tiles.md.)
// Output columns are dealt to the 8-wide DMMA n-tiles in even/odd pairs (tile 2p takes
// columns 16p + {0,2,..,14}, tile 2p+1 the odd ones), so that a lane's two B operands of
// a pair are adjacent in memory (one 16-byte gather instead of two 8-byte ones) and its
// four C values are 32 contiguous bytes.
// Output: C_out = alpha * A_tiles * B + beta * C_in for EVERY row (panels without tile
// columns just run the epilogue); the remainder is added in place afterwards.
// Summation order differs from cpu_spmm_CSR (tolerance-level parity, not bit parity),
// and a padded zero times a non-finite B entry gives NaN where the reference has none.
__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// volatile loads: the compiler must not sink them towards their uses (it did, to save
// registers, and with that serialised the gathers the unrolling was meant to overlap)
__device__ __forceinline__ int ldv_s32(const int *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldv_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldv_v2f64(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// NP = number of 16-column pairs handled by one launch (N <= 16*NP)
template <int NP, bool STRICT>
__global__ void __launch_bounds__(256)
spmm_panels_dmma_kernel(const int npanels, const int M, const int *__restrict__ step_ptr,
                        const int *__restrict__ tcols, const double *__restrict__ tvals,
                        const double *__restrict__ B, const uint32_t ldb, const double *Cin, double *Cout,
                        const int64_t ldc, const double alpha, const double beta, const int N) {
    constexpr int UNR = 4;  // k-steps per batch
    const int lane = threadIdx.x & 31;
    const int panel = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (panel >= npanels) return;  // warp-uniform
    const int kq = lane & 3, rq = lane >> 2;
    const int s0 = __ldg(step_ptr + panel), s1 = __ldg(step_ptr + panel + 1);
    double acc[2 * NP][2];
#pragma unroll
    for (int t = 0; t < 2 * NP; ++t) acc[t][0] = acc[t][1] = 0.0;
    const double *Bq = B + 2 * rq;  // this lane's even/odd column pair inside a 16-wide pair of tiles
    const bool colok[4] = {2 * rq < N, 16 + 2 * rq < N, 32 + 2 * rq < N, 48 + 2 * rq < N};
    // column index and A value of a batch of steps (steps past the panel's end: value 0,
    // column of the last real step, so that every load stays unconditional)
    auto load_meta = [&](const int s, int (&col)[UNR], double (&a)[UNR]) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int su = min(s + u, s1 - 1);
            col[u] = ldv_s32(tcols + (size_t)su * 4 + kq);
            a[u] = ldv_f64(tvals + (size_t)su * 32 + lane);
            if (s + u >= s1) a[u] = 0.0;
        }
    };
    if (s0 < s1) {
        int col[UNR], coln[UNR];
        double a[UNR], an[UNR];
        load_meta(s0, col, a);
        for (int s = s0; s < s1; s += UNR) {
            double2 b[UNR][NP];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const double *brow = Bq + (size_t)(uint32_t)col[u] * ldb;
#pragma unroll
                for (int p = 0; p < NP; ++p) b[u][p] = colok[p] ? ldv_v2f64(brow + 16 * p) : make_double2(0.0, 0.0);
            }
            if (s + UNR < s1) load_meta(s + UNR, coln, an);  // next batch's indices fly with this batch's gathers
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    dmma884(acc[2 * p][0], acc[2 * p][1], a[u], b[u][p].x);
                    dmma884(acc[2 * p + 1][0], acc[2 * p + 1][1], a[u], b[u][p].y);
                }
#pragma unroll
            for (int u = 0; u < UNR; ++u) { col[u] = coln[u]; a[u] = an[u]; }
        }
    }
    const int row = panel * 8 + rq;
    if (row < M) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            // this lane's four consecutive output columns of pair p
            const int n = 16 * p + 4 * kq;
            const double v[4] = {acc[2 * p][0], acc[2 * p + 1][0], acc[2 * p][1], acc[2 * p + 1][1]};
            const double *ci = Cin + (int64_t)row * ldc + n;
            double *co = Cout + (int64_t)row * ldc + n;
            if (n + 3 < N) {
                const double2 c01 = *reinterpret_cast<const double2 *>(ci);
                const double2 c23 = *reinterpret_cast<const double2 *>(ci + 2);
                *reinterpret_cast<double2 *>(co) = vaxpby<STRICT>(alpha, make_double2(v[0], v[1]), beta, c01);
                *reinterpret_cast<double2 *>(co + 2) = vaxpby<STRICT>(alpha, make_double2(v[2], v[3]), beta, c23);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (n + i < N) co[i] = axpby<STRICT>(alpha, v[i], beta, ci[i]);
            }
        }
    }
}

// ---- push of the B image to the other ranks: the exchange step of the row-block partition -----
// The rank that holds B runs ONE kernel per exchange: wait until every peer has finished the SpMM
// that used the previous contents of its image (done[p] >= *pushes, stored by the peer's SpMM
// kernel), copy the local image into every peer's image with plain 16-byte stores through the
// NVLink peer mappings (posted writes: no round trip), fence, and store *pushes + 1 into every
// peer's ready flag from the last block to finish.  The peers launch nothing for the exchange:
// their SpMM kernel waits on the flag (spmm_edgelist_kernel) or a one-warp kernel does
// (wait_push_kernel).  The multi-GPU form of the reference's daisy chain that hands the B window
// from PEG to PEG (src/sextans.cpp:909-941).  All counters are in device memory.
__global__ void __launch_bounds__(256)
push_image_kernel(const int4 *__restrict__ src, const int64_t n16, const PushList peers, const int npeers,
                  const uint32_t *done, uint32_t *pushes, unsigned int *sync_words) {
    __shared__ uint32_t t_sh;
    if (threadIdx.x == 0) {
        const uint32_t t = *reinterpret_cast<volatile uint32_t *>(pushes);
        const long long t0 = clock64();
        for (int p = 0; p < npeers; ++p)
            while ((int)(ld_acquire_sys(done + p) - t) < 0) {
                __nanosleep(64);
                if (clock64() - t0 > 4000000000ll) { atomicExch(sync_words + 1, 1u); break; }
            }
        t_sh = t;
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 v = __ldg(src + i);
        for (int p = 0; p < npeers; ++p) peers.dst[p][i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(sync_words + 3, 1u) == gridDim.x - 1) {
            sync_words[3] = 0;
            __threadfence_system();
            for (int p = 0; p < npeers; ++p) st_release_sys(peers.ready[p], t_sh + 1u);
            *reinterpret_cast<volatile uint32_t *>(pushes) = t_sh + 1u;
        }
    }
}
// publication of a push that an SpMM kernel carried (spmm_edgelist_kernel, npush > 0): launched right
// behind it as a programmatic dependent; everything that kernel stored to the peers is complete when
// the wait returns.
__global__ void publish_push_kernel(const PushList peers, const int npeers, uint32_t *pushes) {
    pdl_launch_dependents();
    pdl_wait();  // the carrying kernel is complete: its stores, the peer stores included, have been performed
    if (threadIdx.x == 0) {
        const uint32_t t = *reinterpret_cast<volatile uint32_t *>(pushes);
        for (int p = 0; p < npeers; ++p) st_relaxed_sys(peers.ready[p], t + 1u);
        *reinterpret_cast<volatile uint32_t *>(pushes) = t + 1u;
    }
}
// the same from a list of ready flags alone: a deferred publication that no SpMM launch picked up (sx_push_publish)
__global__ void publish_list_kernel(const PubList peers, const int npeers, uint32_t *pushes) {
    if (threadIdx.x == 0) {
        const uint32_t t = *reinterpret_cast<volatile uint32_t *>(pushes);
        for (int p = 0; p < npeers; ++p) st_relaxed_sys(peers.ready[p], t + 1u);
        *reinterpret_cast<volatile uint32_t *>(pushes) = t + 1u;
    }
}
// the same handshake around the SpMM kernels that do not carry it themselves
__global__ void wait_push_kernel(const uint32_t *ready, const uint32_t *epoch, unsigned int *sync_words) {
    if (threadIdx.x == 0) {
        const uint32_t step = *reinterpret_cast<const volatile uint32_t *>(epoch) + 1u;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(ready) - step) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000ll) { atomicExch(sync_words + 1, 1u); break; }
        }
    }
}
__global__ void ack_push_kernel(uint32_t *epoch, uint32_t *done_remote) {
    if (threadIdx.x == 0) {
        const uint32_t step = *reinterpret_cast<volatile uint32_t *>(epoch) + 1u;
        *reinterpret_cast<volatile uint32_t *>(epoch) = step;
        __threadfence_system();
        st_release_sys(done_remote, step);
    }
}

// ---- layout changes at the host boundary ---------------------------------------
// column-major (ld = rows) -> row-major (ld = ld_dst, pad columns zero-filled); the
// device-side stand-in for the reference's B/C channel repacking
// (src/sextans-host.cpp:152-195).
template <typename T>
__global__ void __launch_bounds__(256)
colmajor_to_rowmajor_kernel(const int64_t rows, const int cols, const T *__restrict__ src,
                            T *__restrict__ dst, const int64_t ld_dst) {
    __shared__ T tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t r = r0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r + rows * (int64_t)c] : T(0);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t r = r0 + j;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < ld_dst) dst[r * ld_dst + c] = tile[threadIdx.x][j];
    }
}

// B (rowsB x cols) and C_in (rowsC x cols) in one launch: blocks [0, tilesB) take B,
// the rest take C.  Used with the sources in page-locked HOST memory.  threadIdx.x runs
// along a column and every thread moves VEC consecutive rows with one 16-byte access
// when VEC > 1 (the host passes VEC = 16/sizeof(T) only if both row counts are multiples
// of it, which keeps every column start 16-byte aligned), so a warp reads 512 contiguous
// bytes over PCIe per instruction.  dst may point at a column group of a wider image (ld = the
// image's leading dimension, wcols = columns of the group): the host-facing call's column pipeline.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
colmajor_to_rowmajor_pair_kernel(const int64_t rowsB, const int64_t rowsC, const int cols,
                                 const T *__restrict__ srcB, const T *__restrict__ srcC,
                                 T *__restrict__ dstB, T *__restrict__ dstC, const int64_t ld,
                                 const int tcol, const int64_t tilesB, const int wcols) {
    constexpr int TR = 32 * VEC;  // rows per tile
    __shared__ T tile[32][TR + 1];
    pdl_launch_dependents();  // a dependent launch (the fused SpMM) may run its A-side prologue beside this kernel
    int64_t t = blockIdx.x;
    const bool isC = t >= tilesB;
    if (isC) t -= tilesB;
    const int64_t rows = isC ? rowsC : rowsB;
    const T *src = isC ? srcC : srcB;
    T *dst = isC ? dstC : dstB;
    const int64_t r0 = (t / tcol) * TR;
    const int c0 = (int)(t % tcol) * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t r = r0 + (int64_t)threadIdx.x * VEC;
        if (VEC > 1 && c < cols && r + VEC <= rows) {
            struct alignas(16) Pack { T v[VEC]; };
            const Pack p = *reinterpret_cast<const Pack *>(src + r + rows * (int64_t)c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) tile[j][threadIdx.x * VEC + e] = p.v[e];
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                tile[j][threadIdx.x * VEC + e] = (r + e < rows && c < cols) ? src[r + e + rows * (int64_t)c] : T(0);
        }
    }
    __syncthreads();
    for (int j = threadIdx.y; j < TR; j += 8) {
        const int64_t r = r0 + j;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < wcols) dst[r * ld + c] = tile[threadIdx.x][j];  // wcols: columns written (cols..wcols-1 zero-filled)
    }
}

// row-major device image -> column-major HOST memory, 16 bytes per thread (see above)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
rowmajor_to_colmajor_vec_kernel(const int64_t rows, const int cols, const T *__restrict__ src,
                                const int64_t ld_src, T *__restrict__ dst) {
    constexpr int TR = 32 * VEC;
    __shared__ T tile[32][TR + 1];
    const int64_t r0 = (int64_t)blockIdx.x * TR;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < TR; j += 8) {
        const int64_t r = r0 + j;
        const int c = c0 + threadIdx.x;
        tile[threadIdx.x][j] = (r < rows && c < cols) ? src[r * ld_src + c] : T(0);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t r = r0 + (int64_t)threadIdx.x * VEC;
        if (c >= cols) continue;
        if (VEC > 1 && r + VEC <= rows) {
            struct alignas(16) Pack { T v[VEC]; };
            Pack p;
#pragma unroll
            for (int e = 0; e < VEC; ++e) p.v[e] = tile[j][threadIdx.x * VEC + e];
            *reinterpret_cast<Pack *>(dst + r + rows * (int64_t)c) = p;
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                if (r + e < rows) dst[r + e + rows * (int64_t)c] = tile[j][threadIdx.x * VEC + e];
        }
    }
}

// row-major (ld = ld_src) -> column-major (ld = rows); the read-back un-interleave
// (src/sextans-host.cpp:264-270).
template <typename T>
__global__ void __launch_bounds__(256)
rowmajor_to_colmajor_kernel(const int64_t rows, const int cols, const T *__restrict__ src,
                            const int64_t ld_src, T *__restrict__ dst) {
    __shared__ T tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t r = r0 + j;
        const int c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * ld_src + c] : T(0);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[r + rows * (int64_t)c] = tile[threadIdx.x][j];
    }
}

}  // namespace sx
