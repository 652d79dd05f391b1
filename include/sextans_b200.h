/*
 * sextans_b200.h -- C ABI of the B200-native SpMM engine (libsextans_b200.so).
 *
 *   C = alpha * A * B + beta * C      A sparse M x K (CSR), B dense K x N, C dense M x N
 *
 * This is the drop-in boundary for the one device call of the Sextans host
 * program.  Citations are relative to the reference tree (linghaosong/Sextans):
 *
 *   reference interface                                        replaced by
 *   ---------------------------------------------------------  --------------------------
 *   tapa::invoke(Sextans, bitstream, ptr, A[8], B[4], Cin[8],  sx_spmm_f32 / sx_spmm_f64
 *     Cout[8], NUM_ITE, NUM_A_LEN, M, K, P_N, alpha_u, beta_u)
 *     src/sextans-host.cpp:237-251, proto src/sextans.h:20-26
 *   its return value (elapsed ns for all rp_time repeats)      *kernel_ns
 *     src/sextans-host.cpp:237,252
 *   P_N = (rp_time << 16) | N      src/sextans-host.cpp:223    int N, int rp_time
 *   alpha_u / beta_u bit-casts     src/sextans-host.cpp:225-229 float/double alpha, beta
 *   generate_edge_list_for_all_PEs + edge_list_64bit           sx_upload_csr_f32/_f64
 *     (A -> device image) src/sextans-host.cpp:119-146           (CSR goes up as is)
 *   B / C channel-image repacking  src/sextans-host.cpp:152-202 done on the device inside
 *     and read-back un-interleave  src/sextans-host.cpp:264-270   sx_spmm_* (col-major in/out)
 *   tapa::aligned_allocator<T>     src/sextans-host.cpp:23-24  sx_host_alloc / sx_host_free
 *   env TAPAB (backend select)     src/sextans-host.cpp:231-234 int device of sx_create
 *   read_suitsparse_matrix + CSC_2_CSR                         sx_load_mtx_f32/_f64
 *     src/sparse_helper.h:169-259,475-509
 *
 * Dense operands at this boundary are exactly the host program's: column-major,
 * B[k + K*n], C[m + M*n] (src/sextans-host.cpp:102,109; src/sparse_helper.h:283,287).
 *
 * Conventions: every function returns SX_OK (0) or a non-zero sx_status and never
 * throws; sx_last_error() returns the text of the calling thread's last failure.
 * Host pointers stay caller-owned.  One context drives one GPU; calls on one
 * context must be serialised by the caller (the reference makes one blocking call
 * from main).  There is NO CPU fallback: without a usable CUDA device sx_create
 * fails with SX_ERR_NO_DEVICE.
 */
#ifndef SEXTANS_B200_H
#define SEXTANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SX_ABI_VERSION 1

typedef struct sx_ctx sx_ctx;

enum sx_status {
    SX_OK = 0,
    SX_ERR_INVALID = 1,   /* bad argument */
    SX_ERR_CUDA = 2,      /* a CUDA runtime call failed; text in sx_last_error() */
    SX_ERR_NO_DEVICE = 3, /* no CUDA device / ordinal out of range */
    SX_ERR_STATE = 4,     /* call order: e.g. SpMM before sx_upload_csr_* */
    SX_ERR_NOMEM = 5,
    SX_ERR_IO = 6,        /* loader: cannot open / premature end */
    SX_ERR_FORMAT = 7     /* loader: unsupported Matrix Market content */
};

enum sx_dtype { SX_F32 = 0, SX_F64 = 1 };

enum sx_option {
    /* 0 (default) "strict": every product and every sum separately rounded, each
     * row's nonzeros accumulated in stored order -- bit-identical to cpu_spmm_CSR
     * (src/sparse_helper.h:283,287) for rows up to SX_OPT_SPLIT_ROW_NNZ nonzeros.
     * 1 "fast": fused multiply-add. */
    SX_OPT_ARITH = 0,
    /* rows with more nonzeros than this are split across a whole thread block and
     * tree-reduced (summation order differs from the oracle; error ~ 1e-7 fp32 /
     * 1e-16 fp64 relative to the row's |a||b| sum).  0 disables splitting. */
    SX_OPT_SPLIT_ROW_NNZ = 1,
    /* kernel variant: 0 auto; 1 one lane group per row (+ one warp per
     * long-row segment); 2 TMA-staged nnz-balanced work items; 3 B window of each
     * 32-row block staged into shared memory by TMA (banded matrices whose windows fit;
     * falls back to 1 otherwise); 4 the sliding-window kernel for long banded matrices, see
     * SX_OPT_SLIDE; 5 edge lists: a row block's DISTINCT B rows staged in shared memory, 16-bit
     * window-local column indices (any matrix whose rows fit; dense rows of at most 256 bytes).
     * Auto: 5 where a staged B row serves at least two nonzeros (FEM-type matrices); else 3 for
     * matrices below one wave of lane groups whose windows fit, else 1 below one wave, else 2. */
    SX_OPT_KERNEL = 2,
    /* nonzeros per work item; 0 = auto (512 for N*sizeof(T) <= 128 bytes, else 256;
     * smaller for small matrices) */
    SX_OPT_ITEM_NNZ = 3,
    /* sx_spmm_*: page-locked host B and C of up to this many bytes together are read
     * and written by the kernels directly over PCIe (no copy-engine transfers);
     * larger or pageable operands go through cudaMemcpyAsync.  Default 1.5 MiB
     * (measured cross-over on PCIe Gen5: SM-driven transfers reach ~43 GB/s in and
     * ~31 GB/s out, the copy engines more, but each memcpy costs a launch); 0 disables. */
    SX_OPT_ZEROCOPY_BYTES = 4,
    /* dense-tile (blocked) variant on the FP64 tensor cores, fp64 only; read by the NEXT
     * sx_upload_csr_f64.  0 (default) off.  t in 1..8: in every panel of 8 consecutive
     * rows a column used by >= t rows goes into the panel's dense tile (DMMA m8n8k4),
     * the rest of A stays CSR and is added afterwards.  Summation order then differs from
     * cpu_spmm_CSR (fp64 rounding-level differences), and an explicit zero of a tile
     * times a non-finite B entry yields NaN. */
    SX_OPT_TILE_MIN_ROWS = 5,
    /* column windows (the reference's K windows, src/sextans.h:11, src/sextans.cpp:57,337-381,
     * with L2 in the role of the on-chip B buffer); read by the NEXT sx_upload_csr_*.
     * 0 (default) off.  W >= 1: A is cut into windows of W consecutive columns and one SpMM
     * becomes ceil(K/W) passes of the TMA-staged kernel, window by window, a row's running
     * sum travelling through a device buffer, so that the B rows a pass gathers -- W rows,
     * W * ld * sizeof(T) bytes -- stay resident in the 126 MB L2 instead of coming from HBM
     * once per nonzero.  Worth it when B is much larger than L2's reach and rows hold many
     * nonzeros per window (power-law / wide-band matrices); pick W so that a window of B is
     * ~32 MiB.  Strict mode stays bit-identical to cpu_spmm_CSR for rows stored in ascending
     * column order (the loader's order); rows stored otherwise are summed window by window
     * (rounding-level differences).  Ignored when SX_OPT_TILE_MIN_ROWS is in effect or when
     * K <= W. */
    SX_OPT_COL_WINDOW_ROWS = 6,
    /* L2 prefetch hints.  Edge-list kernel: before the dependent-launch wait, ask L2 for the B rows
     * and C_in rows the block is about to read (on unless 0).
     * TMA-staged kernel: 1 = while a batch's B-row gathers are in flight, prefetch the next
     * batch's B rows into L2 (prefetch.global.L2; their column indices are already in the
     * shared-memory tile), taking the DRAM latency of the gathers off the lane group's
     * critical path.  Results are unaffected.  0 = off.  -1 (default) = auto: on when B is
     * larger than 32 MiB and a dense row is at most 256 bytes (measured: power-law
     * M=K=1e6, nnz=9.6e7, N=16 fp64 1.63 -> 1.49 ms; uniform N=128 fp32, DRAM-bound,
     * 1.456 -> 1.470 ms, hence left off for wide rows). */
    SX_OPT_PREFETCH = 7,
    /* An sx_spmm_* call that takes the zero-copy path on a matrix that runs the edge-list kernel, with
     * rp_time <= 1 and kernel_ns == NULL (nobody asks for the kernel-only time), does not stage C on the
     * device at all.
     * 1: two launches -- a kernel stages B, and the SpMM kernel (its programmatic dependent) reads C_in
     *    from and writes C to the caller's page-locked array itself.  SX_INFO_HOST_PATH reports 2.
     * -1 (default) / 2: ONE launch where the kernel's grid is resident at once (else as 1): every block
     *    fetches its share of B and its C_in tile from the caller's arrays with cp.async over PCIe, the
     *    blocks exchange B through the device image (L2) behind a counter, and the call is pipelined over
     *    SX_OPT_HOST_GROUPS column groups, so that the first columns of C leave over PCIe while the last
     *    columns of B and C_in are still arriving -- the link's two directions overlap (measured floors,
     *    scripts/micro/pcie_floor.cu: in then out 52 us, both at once 43 us, for nasa4704 N=16 fp64).
     *    SX_INFO_HOST_PATH reports 3.
     * 0: off (three launches, C staged on the device).  Same arithmetic, same results in every mode. */
    SX_OPT_HOST_FUSED = 8,
    /* Programmatic dependent launch.  -1 (default): the edge-list kernel is launched with
     * programmatic stream serialization -- its A-side prologue (records, row pointers, TMA of its
     * slice of A, L2 prefetch hints for B and C_in) runs while the previous kernel of the stream
     * is still finishing; B and C_in are only read after griddepcontrol.wait (nasa4704: 4.3 ->
     * 3.5 us per step).  1: variant 3 as well.  0: never.  Results are unaffected. */
    SX_OPT_PDL = 9,
    /* retired (taller row blocks of variant 3: measured 10 % at best, superseded by variant 5);
     * the value is accepted and ignored */
    SX_OPT_WINDOW_ROWS = 10,
    /* Read by the NEXT sx_upload_csr_*.
     * n in 1..8: plan n chains per SM for the sliding-window kernel (variant 4, selected with
     * SX_OPT_KERNEL = 4): a thread block walks a run of consecutive 32-row steps with B held in
     * a shared-memory ring that follows the band -- each step loads only the B rows above the
     * highest one loaded so far -- and the next steps' loads overlap the current step's
     * arithmetic.  For LONG banded matrices, where variant 3 keeps re-fetching nearly the same
     * window (FEM-like band of 100, M = 1e6, fp64: 0.78 ms against 1.11 ms).  Falls back to the
     * automatic choice where the ring the plan needs does not fit in shared memory for the N in
     * use.  Results are unaffected.  0 (default): no plan. */
    SX_OPT_SLIDE = 11,
    /* EXPERIMENTAL (not yet run on hardware).  1: the first SpMM for a column count N (device
     * operands with C_out != C_in, outside graph capture) times every kernel variant that
     * applies to this matrix -- 1, 2 without / with the L2 prefetch, 3, and 4 when a plan
     * exists -- on the caller's own operands (one warm-up, three timed launches each) and
     * later calls with that N use the fastest; SX_INFO_TUNED_KERNEL reports it.  The selection
     * rules of SX_OPT_KERNEL = 0 are thresholds measured on a handful of matrices; this
     * measures the matrix at hand.  Cleared by the next upload. */
    SX_OPT_AUTOTUNE = 12,
    /* N passes: columns of B and C per pass of the SpMM kernel (a multiple of 8).  The reference runs
     * every SpMM as ceil(N/8) passes over A, 8 columns of B and C at a time (rp_time_N,
     * src/sextans.cpp:57,84,328,474), so that a window of B fits its on-chip buffers; here a pass
     * gathers a panel-wide slice of every B row, and the slice of the WHOLE of B (K * panel * sizeof(T)
     * bytes) stays resident in the 126 MB L2 -- it comes from HBM once per pass instead of once per
     * nonzero, at the price of streaming A once per pass.  0 (default) = auto: ONE pass (on a B200 a
     * pass costs its nonzeros, not its bytes: C4 1.51 / 2.68 / 4.27 ms at 64 / 32 / 16 columns per
     * pass against 1.46 ms in one; profiles/r02_n_passes.txt).  Results are unaffected
     * (the same chain of operations per element of C). */
    SX_OPT_PANEL_COLS = 13,
    /* Column groups of the fused host-facing call (SX_OPT_HOST_FUSED): the columns of C are independent
     * and the caller's arrays are column-major, so the call is pipelined over this many groups of
     * consecutive columns.  One-launch form: 0 (default) = auto (2), at most 8.  Two-launch form: that
     * many (B staging, SpMM) kernel pairs chained by programmatic dependent launch; 0 = auto (1: on
     * the measured host every further launch costs ~3.7 us, more than the overlap buys).
     * Results are unaffected. */
    SX_OPT_HOST_GROUPS = 14
};

enum sx_info {
    SX_INFO_LAUNCHES = 0,    /* kernels launched by this context so far */
    SX_INFO_M = 1,
    SX_INFO_K = 2,
    SX_INFO_NNZ = 3,
    SX_INFO_DTYPE = 4,
    SX_INFO_SPLIT_ROWS = 5,  /* rows that take the split path */
    SX_INFO_LAST_KERNEL = 6, /* variant id of the last SpMM launch */
    SX_INFO_LD = 7,          /* leading dimension (elements) of the context's row-major B/C */
    SX_INFO_ITEMS = 8,       /* work items of the main kernel */
    SX_INFO_ITEM_NNZ = 9,    /* nonzero budget per work item in use */
    SX_INFO_HOST_PATH = 10,  /* last sx_spmm_* call: 0 copy engines, 1 zero-copy kernels, 2 zero-copy with
                              * C carried by the SpMM kernel, 3 the whole call as one kernel (SX_OPT_HOST_FUSED) */
    SX_INFO_TILE_NNZ = 11,   /* nonzeros held in dense tiles */
    SX_INFO_TILE_SLOTS = 12, /* tile slots incl. explicit zeros (fill = TILE_NNZ / TILE_SLOTS) */
    SX_INFO_REST_NNZ = 13,   /* nonzeros left to the CSR kernels */
    SX_INFO_UPLOAD_SERIAL = 14, /* process-wide serial number of the matrix this context holds
                                 * (every successful sx_upload_csr_* draws a new one; 0: none) */
    SX_INFO_COL_WINDOWS = 15,   /* column windows in use (0: the matrix is not windowed) */
    SX_INFO_EXCHANGE_TIMEOUTS = 17, /* nonzero if a device-side wait on a peer flag ever gave up (~2 s) */
    SX_INFO_EDGE_BLOCKS = 18,   /* row blocks of the edge-list plan of the last launch (variant 5) */
    SX_INFO_EDGE_COLS = 19,     /* B rows those blocks stage per SpMM (sum of their distinct columns) */
    SX_INFO_PUSH_PENDING = 20,  /* 1 if the last SpMM launch carried a push whose publication was deferred
                                 * (sx_spmm_fuse_push_deferred) and is still owed; 0 if nothing is owed */
    SX_INFO_TUNED_KERNEL = 16   /* SX_OPT_AUTOTUNE's choice for the current N: 10 * variant + (1 if
                                 * with the L2 prefetch), 0 if nothing has been tuned */
};

/* ---- library ------------------------------------------------------------- */
int sx_abi_version(void);
const char *sx_last_error(void);
const char *sx_status_name(int status);
int sx_device_count(int *count);

/* ---- context ------------------------------------------------------------- */
int sx_create(int device, sx_ctx **out);
int sx_destroy(sx_ctx *ctx);
/* Run this context's work on an existing cudaStream_t (pass it as void*); NULL
 * returns to the context's own stream. */
int sx_set_stream(sx_ctx *ctx, void *cuda_stream);
int sx_set_option(sx_ctx *ctx, int option, int64_t value);
int sx_get_info(sx_ctx *ctx, int what, int64_t *value);
int sx_synchronize(sx_ctx *ctx);

/* ---- A: CSR upload (replaces the FPGA edge-list preprocessing) ------------ */
/* rowptr has M+1 entries with rowptr[0] == 0 and rowptr[M] == nnz; column indices
 * must lie in [0, K).  Rows keep their stored nonzero order (the loader's order is
 * ascending column).  Re-uploading replaces the previous matrix. */
int sx_upload_csr_f32(sx_ctx *ctx, int M, int K, int64_t nnz, const int32_t *rowptr,
                      const int32_t *colidx, const float *val);
int sx_upload_csr_f64(sx_ctx *ctx, int M, int K, int64_t nnz, const int32_t *rowptr,
                      const int32_t *colidx, const double *val);

/* ---- the SpMM call (replaces tapa::invoke(Sextans, ...)) ------------------ */
/* B: K x N column-major (ld K), read only.  C: M x N column-major (ld M), in/out.
 * The kernel is run rp_time times (rp_time < 1 is treated as 1, like
 * src/sextans.cpp:52-54), every repeat starting from the ORIGINAL C (the FPGA
 * re-reads mat_C_ch_in each repeat, src/sextans.cpp:143), so the result does not
 * depend on rp_time.  *kernel_ns (may be NULL) receives the device time of all
 * repeats together, measured with CUDA events on the context's stream; host<->
 * device copies and layout changes are outside it, as the FPGA's DMA is outside
 * the reference's figure. */
int sx_spmm_f32(sx_ctx *ctx, int N, float alpha, const float *B, float beta, float *C,
                int rp_time, double *kernel_ns);
int sx_spmm_f64(sx_ctx *ctx, int N, double alpha, const double *B, double beta, double *C,
                int rp_time, double *kernel_ns);

/* The same call without its final host synchronisation (rp_time = 1, no kernel time): it returns once
 * the work is enqueued on the context's stream; B and C must stay valid and untouched until
 * sx_synchronize(ctx) returns, which also reports a failed call.  For callers that keep several
 * calls in flight -- two contexts on two streams with double-buffered page-locked operands: one
 * call's results leave over PCIe while the next call's operands arrive (the link is full duplex).
 * Pageable operands make the call blocking (cudaMemcpyAsync from/to pageable memory is). */
int sx_spmm_enqueue_f32(sx_ctx *ctx, int N, float alpha, const float *B, float beta, float *C);
int sx_spmm_enqueue_f64(sx_ctx *ctx, int N, double alpha, const double *B, double beta, double *C);

/* ---- the same call in stages (what sx_spmm_* does internally) ------------- */
int sx_stage_B_f32(sx_ctx *ctx, int N, const float *B_colmajor);
int sx_stage_B_f64(sx_ctx *ctx, int N, const double *B_colmajor);
int sx_stage_C_f32(sx_ctx *ctx, int N, const float *C_colmajor);
int sx_stage_C_f64(sx_ctx *ctx, int N, const double *C_colmajor);
int sx_launch_f32(sx_ctx *ctx, float alpha, float beta, int rp_time, double *kernel_ns);
int sx_launch_f64(sx_ctx *ctx, double alpha, double beta, int rp_time, double *kernel_ns);
int sx_fetch_C_f32(sx_ctx *ctx, float *C_colmajor);
int sx_fetch_C_f64(sx_ctx *ctx, double *C_colmajor);
/* One blocking call on the B image the context already holds -- staged by sx_stage_B_*, pushed by a
 * peer (sx_spmm_expect_push) or filled by a collective through sx_device_B -- and the caller's
 * column-major host C (M x N, in/out): what a rank that does NOT hold the host B calls in the
 * row-block path.  Page-locked C on a matrix that runs the edge-list kernel: one launch, C never
 * staged on the device (SX_OPT_HOST_FUSED); otherwise C is staged, multiplied and fetched. */
int sx_spmm_staged_B_f32(sx_ctx *ctx, int N, float alpha, float beta, float *C);
int sx_spmm_staged_B_f64(sx_ctx *ctx, int N, double alpha, double beta, double *C);
/* Device address of the context's staged row-major B (K x ld, ld from SX_INFO_LD):
 * the buffer a collective (NCCL broadcast over NVLink) fills on the non-root
 * ranks.  Allocates it for N columns if needed. */
int sx_device_B(sx_ctx *ctx, int N, void **dptr, size_t *bytes);

/* ---- device-resident operands (stream-ordered, no copies, no sync) -------- */
/* dB: K rows of ldb elements, dCin/dCout: M rows of ldc elements, all ROW-major
 * device memory, 16-byte aligned, ldb/ldc multiples of 16/sizeof(T) elements and
 * >= N.  dCin == dCout is allowed (in place).  Enqueues on the context's stream
 * and returns. */
int sx_spmm_device_f32(sx_ctx *ctx, int N, float alpha, const float *dB, int64_t ldb,
                       float beta, const float *dCin, float *dCout, int64_t ldc);
int sx_spmm_device_f64(sx_ctx *ctx, int N, double alpha, const double *dB, int64_t ldb,
                       double beta, const double *dCin, double *dCout, int64_t ldc);
/* Several B's at once (the reference's call is one B per invoke, src/sextans-host.cpp:237-251; a
 * caller with nb right-hand-side blocks invokes it nb times): C_out[b] = alpha * A * B[b] + beta *
 * C_in[b] for b < nb, operand b at base + b * stride (elements; strides multiples of 16/sizeof(T),
 * strideC >= M * ldc, strideB may be 0).  A matrix that runs the edge-list kernel takes the whole
 * batch in ONE launch (grid.y = nb): the launch cost that dominates a small SpMM is paid once and
 * the slice of A a block reads is shared by the nb blocks that use it.  Other kernels are launched
 * once per operand triple.  Every result is bitwise what nb sx_spmm_device_* calls produce. */
int sx_spmm_device_batch_f32(sx_ctx *ctx, int N, int nb, float alpha, const float *dB, int64_t ldb,
                             int64_t strideB, float beta, const float *dCin, float *dCout,
                             int64_t ldc, int64_t strideC);
int sx_spmm_device_batch_f64(sx_ctx *ctx, int N, int nb, double alpha, const double *dB, int64_t ldb,
                             int64_t strideB, double beta, const double *dCin, double *dCout,
                             int64_t ldc, int64_t strideC);
/* Layout changes between the host program's column-major operands and the
 * engine's row-major ones, on device memory, enqueued on the context's stream:
 * src is rows x cols column-major (ld rows); dst is row-major with leading
 * dimension ld_dst (columns cols..ld_dst-1 are zero-filled), and the reverse. */
int sx_colmajor_to_rowmajor(sx_ctx *ctx, int dtype, int64_t rows, int cols, const void *d_src,
                            void *d_dst, int64_t ld_dst);
int sx_rowmajor_to_colmajor(sx_ctx *ctx, int dtype, int64_t rows, int cols, const void *d_src,
                            int64_t ld_src, void *d_dst);

/* ---- peer memory: B to the other GPUs' contexts over NVLink, no collective ---- */
/* One process per GPU.  For a small B the launch latency of a collective dwarfs the
 * transfer (600 KB: ~60 us through NCCL against ~4 us of SpMM), so the row-block path
 * instead lets the rank that holds B PUSH its image into the other ranks' images through
 * CUDA-IPC peer mappings (below: "exchange of B by PUSH").  Handles are CUDA IPC handles
 * (64 bytes) to be exchanged by whatever transport the host program has
 * (torch.distributed object gather in sextans_b200/rowblock.py). */
#define SX_IPC_HANDLE_BYTES 64
int sx_device_alloc(sx_ctx *ctx, size_t bytes, void **dptr);  /* zero-filled */
int sx_device_free(sx_ctx *ctx, void *dptr);
/* The handle names the whole driver allocation that contains dptr (small cudaMalloc requests are
 * carved out of larger blocks); sx_ipc_offset gives dptr's offset inside it, which the importing
 * side adds to the address sx_ipc_import returns.  A process must open a given handle only once. */
int sx_ipc_export(sx_ctx *ctx, const void *dptr, unsigned char handle[SX_IPC_HANDLE_BYTES]);
int sx_ipc_offset(sx_ctx *ctx, const void *dptr, size_t *offset);
int sx_ipc_import(sx_ctx *ctx, const unsigned char handle[SX_IPC_HANDLE_BYTES], void **dptr);
int sx_ipc_close(sx_ctx *ctx, void *dptr);
/* ---- exchange of B by PUSH (the row-block partition's one exchange step, small B) -------------
 * The rank that holds B copies its image into every peer's image with ONE kernel (posted 16-byte
 * stores over NVLink); the peers launch nothing for it: their next SpMM waits on a flag in its own
 * prologue and acknowledges from its last block.  The multi-GPU form of the reference's chain that
 * hands the B window from PEG to PEG (src/sextans.cpp:909-941).  All counters are 32-bit words in
 * device memory (sx_device_alloc, zero-filled; exchanged with sx_ipc_*), so captured launches can
 * be replayed:
 *   on every peer, per image:   ready  (written by the pusher), epoch (SpMMs served; local)
 *   on the pusher, per image:   pushes (pushes done; local), done[peer] (written by the peers)
 * sx_push_B: waits until done[p] >= pushes for every peer (they have finished the SpMM that used the
 *   previous contents), copies `bytes` bytes of the row-major B image `image` (16-byte units; what
 *   sx_device_B reports) to peer_images[p], then stores pushes + 1 into peer_ready_flags[p] and
 *   into pushes.  npeers <= 15.
 * sx_spmm_expect_push: the NEXT SpMM launch of this context (sx_spmm_device_* / sx_launch_*) waits
 *   until *ready_flag >= *epoch_counter + 1 before it reads B, and when it is complete advances
 *   *epoch_counter and stores it into done_flag (the pusher's done[this peer], peer-mapped).
 *   One-shot.  A wait that sees nothing for ~2 s gives up (SX_INFO_EXCHANGE_TIMEOUTS). */
int sx_push_B(sx_ctx *ctx, const void *image, size_t bytes, void *const *peer_images,
              void *const *peer_ready_flags, int npeers, const void *done_flags, void *pushes_counter);
int sx_spmm_expect_push(sx_ctx *ctx, const void *ready_flag, void *epoch_counter, void *done_flag);
/* sx_spmm_fuse_push: on the rank that holds B, make the push PART OF the next SpMM launch of this
 *   context (sx_spmm_device_* / sx_launch_*): its thread blocks copy the B image that launch reads
 *   (all K rows, from column 0) into the peers' images on their way to their rows, and its last
 *   block publishes the step -- compute and exchange in one kernel, nothing else launched on any
 *   rank.  Same arguments and counters as sx_push_B.  One-shot.  (Kernels other than the
 *   edge-list variant run the push as a kernel of its own right before them.) */
int sx_spmm_fuse_push(sx_ctx *ctx, void *const *peer_images, void *const *peer_ready_flags, int npeers,
                      const void *done_flags, void *pushes_counter);
/* Deferred publication: in a chain of dependent launches the one-warp publish kernel behind every
 * pushing SpMM gates the next SpMM's dependent-launch wait (~1.1 us per step).  A caller that keeps
 * at least two images in rotation can take it out of the chain:
 * sx_spmm_fuse_push_deferred: as sx_spmm_fuse_push, but nothing publishes the push behind the launch.
 *   SX_INFO_PUSH_PENDING tells afterwards whether the publication is still owed (1) or the launch
 *   published it itself after all (0: a kernel other than the edge-list variant ran).
 * sx_spmm_fuse_publish: the NEXT SpMM launch of this context -- on the same stream as the launch
 *   that carried the push, a different image's counter -- stores pushes + 1 into the peers' ready
 *   flags and into the counter right after its dependent-launch wait, i.e. once the carrying
 *   kernel is complete.  One-shot.  (Kernels other than the edge-list variant run a one-warp
 *   kernel for it first.)
 * sx_push_publish: the same as a one-warp kernel of its own -- what ends a sequence (before a host
 *   sync, at the end of a captured graph). */
int sx_spmm_fuse_push_deferred(sx_ctx *ctx, void *const *peer_images, void *const *peer_ready_flags,
                               int npeers, const void *done_flags, void *pushes_counter);
int sx_spmm_fuse_publish(sx_ctx *ctx, void *const *peer_ready_flags, int npeers, void *pushes_counter);
int sx_push_publish(sx_ctx *ctx, void *const *peer_ready_flags, int npeers, void *pushes_counter);
/* Page-locked host memory for B and C (stands in for tapa::aligned_allocator). */
int sx_host_alloc(size_t bytes, void **ptr);
int sx_host_free(void *ptr);
/* Contiguous row blocks with ~equal nonzeros: bounds[0]=0 <= ... <= bounds[parts]=M.
 * The GPU-count analogue of the reference's row -> PE map (src/sparse_helper.h:370). */
int sx_partition_rows(int M, const int32_t *rowptr, int parts, int32_t *bounds);
/* Column windows of W = window_rows columns (what SX_OPT_COL_WINDOW_ROWS makes of A; host
 * only): *nwin = ceil(K/W) windows, each a CSR of its own over all M rows.
 *   win_rowptr  nwin x (M+1): row pointers of window w at [w*(M+1), (w+1)*(M+1)), from 0
 *   win_base    nwin + 1: window w's entries are order[win_base[w] .. win_base[w+1])
 *   order       nnz: position in the caller's colidx/val of every entry, window-major;
 *               inside (window, row) the stored order is kept
 *   ascending   (may be NULL) 1 if every row is stored in non-decreasing column order
 * Arrays are malloc'ed; release each with sx_free. */
/* Plan of the sliding-window kernel (what SX_OPT_SLIDE builds at upload; host only): rows in
 * steps of 32, chains = runs of consecutive steps with about equal nonzeros.
 *   steps   4 ints per step: {load_lo, load_hi, nnz_begin, nnz_end} -- B rows [load_lo, load_hi)
 *           enter the ring at this step (everything above the highest row loaded so far, up to
 *           the highest column the step touches)
 *   chains  2 ints per chain: {first step, last step + 1}
 *   ring_rows         rows the ring must hold so that a step's columns stay resident while the
 *                     next step's rows arrive (the kernel rounds it up to a power of two)
 *   max_step_entries  capacity (entries, a multiple of 4) of one of the two A buffers
 * Arrays are malloc'ed; release each with sx_free. */
int sx_plan_slide(int M, const int32_t *rowptr, const int32_t *colidx, int nchains_wanted, int *nsteps,
                  int32_t **steps, int *nchains, int32_t **chains, int *ring_rows, int *max_step_entries);
/* Plan of the edge-list kernel (variant 5; host only).  Row blocks of consecutive rows holding
 * about the same number of nonzeros, each with the ascending list of the DISTINCT columns its
 * nonzeros touch -- the block's compacted B window -- and, per nonzero, the 16-bit index of its
 * column inside that window: the GPU form of the reference's window-local column field (col14 of
 * the packed edge word, src/sparse_helper.h:419-443, src/sextans.cpp:398-402) and of its
 * equal-length PE lists (:345-403).
 *   row_bytes    bytes of one staged B row in shared memory (lanes per row x 16)
 *   max_rows     most rows a block may hold
 *   nnz_target   nonzeros per block to aim for (0: blocks of max_rows rows)
 *   smem_budget  shared memory one block may use (window + its slices of values and local
 *                columns + its column list + its row pointers)
 * The streams of A are ROW-ALIGNED: row r's entries start at prow[r], a multiple of 8 entries, and
 * are padded to a multiple of 8, so that the kernel fetches 8 (local column, value) pairs with whole
 * 16-byte shared-memory loads; the pad entries' additions are predicated off (the reference pads its
 * PE lists with bubbles, src/sparse_helper.h:345-403).
 *   blocks  8 ints per block: {row_begin, nrows, pnz_begin, pnz_end, col_begin, ncols, 0, smem_bytes}
 *           with pnz_* in padded coordinates (prow[row_begin], prow[row_begin + nrows])
 *   cols    *ncols ints: the blocks' column lists back to back, each starting at a multiple of 4
 *           entries; pad entries repeat the block's last column
 *   prow    M + 1 ints: padded row starts
 *   lcol    prow[M] uint16: entry k of row r at prow[r] + k, pad entries 0
 *   total_cols  sum of ncols over the blocks (B rows staged per SpMM)
 *   max_smem    largest smem_bytes
 * *nblocks = 0 (and SX_OK) if some single row does not fit the budget.  Arrays are malloc'ed;
 * release each with sx_free. */
int sx_plan_edge_lists(int M, int K, const int32_t *rowptr, const int32_t *colidx, int row_bytes, int elem_bytes,
                       int max_rows, int64_t nnz_target, int smem_budget, int *nblocks, int32_t **blocks,
                       int64_t *ncols, int32_t **cols, uint16_t **lcol, int64_t *total_cols, int *max_smem,
                       int32_t **prow);
int sx_split_col_windows(int M, int K, const int32_t *rowptr, const int32_t *colidx, int window_rows,
                         int *nwin, int32_t **win_rowptr, int64_t **win_base, int32_t **order,
                         int *ascending);
/* Matrix Market -> CSR with the reference loader's semantics (symmetric expansion,
 * +0.0 entries dropped, 1-based -> 0-based, columns ascending within a row,
 * duplicates kept).  Arrays are malloc'ed; release each with sx_free. */
int sx_load_mtx_f32(const char *path, int *M, int *K, int64_t *nnz, int32_t **rowptr,
                    int32_t **colidx, float **val);
int sx_load_mtx_f64(const char *path, int *M, int *K, int64_t *nnz, int32_t **rowptr,
                    int32_t **colidx, double **val);
void sx_free(void *ptr);

/* ---- the literal Sextans(...) argument list: FPGA channel images -------------- */
/* For a host program that keeps the reference's OWN preprocessing
 * (generate_edge_list_for_all_PEs + edge_list_64bit, src/sparse_helper.h:345-473, and
 * the B/C channel repacking of src/sextans-host.cpp:152-202) and wants to change
 * nothing but the device call: sx_sextans_invoke takes exactly the arguments of
 *     void Sextans(mmap<int> edge_list_ptr, mmaps<ap_uint<512>,8> edge_list_ch,
 *                  mmaps<float_v16,4> mat_B_ch, mmaps<float_v16,8> mat_C_ch_in,
 *                  mmaps<float_v16,8> mat_C_ch, int NUM_ITE, int NUM_A_LEN, int M, int K,
 *                  int P_N, int alpha_u, int beta_u)              src/sextans.h:20-26
 * as plain pointers, and *kernel_ns is what tapa::invoke returns
 * (src/sextans-host.cpp:237).  integration/sextans_kernel_b200.cpp defines Sextans()
 * itself on top of it, and include/tapa_compat/ supplies the few TAPA names the
 * unmodified host source uses.  fp32 only, like the images.
 *
 * Image layouts:
 *   edge word  [63:50] column inside its 4096-column window, [49:32] row/64 (bit 17 set
 *              = bubble), [31:0] fp32 bits                 src/sparse_helper.h:419-443
 *   A          slot i of channel ch is the 8 words [8i, 8i+8); word bitrev3(q) belongs to
 *              PE ch + 8q, which owns the rows r with r % 64 == PE
 *                                                          src/sparse_helper.h:451-464
 *   ptr        ptr[w]..ptr[w+1] = slots of column window w, NUM_ITE windows, ptr[NUM_ITE]
 *              == NUM_A_LEN                                 src/sparse_helper.h:359,400
 *   B          channel (n/2)%4, element (k/8)*16 + (n%2)*8 + k%8 + 2*roundup(K,8)*(n/8)
 *                                                          src/sextans-host.cpp:158-171
 *   C in/out   channel m%8, element roundup(M,16)*(n/8) + (m/8)*8 + n%8
 *                                                          src/sextans-host.cpp:181-195,269
 *   P_N        (rp_time << 16) | N, rp_time 0 read as 1; N is processed in whole blocks
 *              of 8 columns                src/sextans-host.cpp:223, src/sextans.cpp:52-54
 *   alpha_u, beta_u   fp32 bit patterns                    src/sextans-host.cpp:225-229
 *
 * A images are decoded to CSR on the host (row order inside a PE stream is the order
 * the FPGA accumulates in, so the CSR rows come out in ascending column order) and
 * uploaded; a second call with the same images (same contents) reuses the uploaded A.
 * Rows M..roundup(M,16)-1 of mat_C_ch receive alpha*0 + beta*C_in like the FPGA's
 * whole-word writes; nothing beyond is touched. */
#define SX_IMAGES_A_CHANNELS 8
#define SX_IMAGES_B_CHANNELS 4
#define SX_IMAGES_C_CHANNELS 8
#define SX_IMAGES_WINDOW 4096
#define SX_IMAGES_PES 64
int sx_sextans_invoke(sx_ctx *ctx, const int32_t *edge_list_ptr,
                      const uint64_t *const edge_list_ch[SX_IMAGES_A_CHANNELS],
                      const float *const mat_B_ch[SX_IMAGES_B_CHANNELS],
                      const float *const mat_C_ch_in[SX_IMAGES_C_CHANNELS],
                      float *const mat_C_ch[SX_IMAGES_C_CHANNELS], int NUM_ITE, int NUM_A_LEN,
                      int M, int K, int P_N, int alpha_u, int beta_u, double *kernel_ns);
/* kernel time of the calling thread's last successful sx_sextans_invoke (what the
 * tapa::invoke of include/tapa_compat/tapa.h returns) */
double sx_sextans_last_kernel_ns(void);
/* elements every channel buffer must at least hold: 64-bit words of an A channel, floats
 * of a B channel, floats of a C channel (the host pads them further to 512/1024) */
int64_t sx_images_A_words(int NUM_A_LEN);
int64_t sx_images_B_floats(int K, int N);
int64_t sx_images_C_floats(int M, int N);
/* the pieces, usable on their own (host only, no GPU needed): */
int sx_images_decode_A(const int32_t *edge_list_ptr,
                       const uint64_t *const edge_list_ch[SX_IMAGES_A_CHANNELS], int NUM_ITE,
                       int NUM_A_LEN, int M, int K, int64_t *nnz, int32_t **rowptr,
                       int32_t **colidx, float **val); /* arrays malloc'ed: sx_free */
int sx_images_decode_B(const float *const mat_B_ch[SX_IMAGES_B_CHANNELS], int K, int N,
                       float *B_colmajor /* K x roundup(N,8) */);
int sx_images_decode_C(const float *const mat_C_ch[SX_IMAGES_C_CHANNELS], int M, int N,
                       float *C_colmajor /* M x roundup(N,8) */);
/* writes rows 0..M-1 from C_colmajor and rows M..roundup(M,16)-1 as alpha*0 + beta*pad
 * of mat_C_ch_in (NULL: those rows are left alone) */
int sx_images_encode_C(const float *C_colmajor, int M, int N, float alpha, float beta,
                       const float *const mat_C_ch_in[SX_IMAGES_C_CHANNELS],
                       float *const mat_C_ch[SX_IMAGES_C_CHANNELS]);

#ifdef __cplusplus
}
#endif
#endif /* SEXTANS_B200_H */
