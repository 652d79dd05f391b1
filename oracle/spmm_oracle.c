/*
 * oracle/spmm_oracle.c -- CPU restatement of the Sextans golden SpMM path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sextans_b200/ may include, link or
 * call this file; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the reference's own header compiled unmodified into
 * oracle/_ref/libsextans_ref.so (oracle/ref_shim.cpp) and (b) the golden vectors
 * under tests/golden/ that were generated from that library
 * (tests/golden/make_golden.py).
 *
 * What is restated (all citations relative to /root/reference):
 *   sx_oracle_spmm_csr_{f32,f64}   src/sparse_helper.h:262-290  cpu_spmm_CSR
 *   sx_oracle_load_mtx_{f32,f64}   src/sparse_helper.h:89-259   mm_init_read /
 *                                  load_S_matrix / read_suitsparse_matrix(CSC)
 *                                  src/mmio.h:254-367           banner + size line
 *                                  src/sparse_helper.h:475-509  CSC_2_CSR
 *   sx_oracle_verify_f32           src/sextans-host.cpp:262-289 mismatch criterion
 *   sx_oracle_init_dense_*         src/sextans-host.cpp:100-111 B = 1, C = (m+1)(n+1)/M/N
 *   sx_oracle_sextans_images       src/sextans.cpp:285-570      the accelerator's own dataflow
 *                                  (PEG_Bmtx, PEG_Cmtx), :196-233 (alpha/beta), run on
 *                                  the FPGA channel images in the order the hardware
 *                                  streams them (functional model, no timing)
 *
 * The reference is fp32-only; the f64 entry points are
 * the same statements instantiated for double.
 *
 * Arithmetic contract (must be compiled with -ffp-contract=off):
 *   per row, nonzeros in stored order, psum[n] = psum[n] + (a * b)   two roundings
 *   C = (alpha * psum) + (beta * C)                                  three roundings
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SX_ORACLE_OK 0
#define SX_ORACLE_EIO 1
#define SX_ORACLE_EBANNER 2
#define SX_ORACLE_ESIZE 3
#define SX_ORACLE_ENOTCOORD 4
#define SX_ORACLE_ECOMPLEX 5
#define SX_ORACLE_EINDEX 6
#define SX_ORACLE_ENOMEM 7

/* ------------------------------------------------------------------------- */
/* cpu_spmm_CSR  (src/sparse_helper.h:262-290)                               */
/* ------------------------------------------------------------------------- */

/* One row of the reference loop nest (sparse_helper.h:280-288): a zeroed psum
 * of N entries, every stored nonzero of the row in order, B and C column-major
 * with leading dimensions K and M.  psum is caller scratch of N entries. */
#define SX_DEFINE_ROW(NAME, T)                                                 \
    static void NAME(int64_t i, int N, int64_t M, int64_t K, const int *rowptr, \
                     const int *colidx, const T *val, T alpha, const T *B,     \
                     T beta, T *C, T *psum) {                                  \
        for (int n = 0; n < N; ++n) psum[n] = (T)0;                            \
        for (int64_t j = rowptr[i]; j < rowptr[i + 1]; ++j) {                  \
            const T a = val[j];                                                \
            const T *bcol = B + colidx[j];                                     \
            for (int n = 0; n < N; ++n) {                                      \
                const T prod = a * bcol[K * (int64_t)n];                       \
                psum[n] = psum[n] + prod;                                      \
            }                                                                  \
        }                                                                      \
        for (int n = 0; n < N; ++n) {                                          \
            T *c = C + i + M * (int64_t)n;                                     \
            const T left = alpha * psum[n];                                    \
            const T right = beta * (*c);                                       \
            *c = left + right;                                                 \
        }                                                                      \
    }

SX_DEFINE_ROW(sx_row_f32, float)
SX_DEFINE_ROW(sx_row_f64, double)

/* Single-threaded, exactly the reference's traversal (rows ascending). */
#define SX_DEFINE_SPMM(NAME, ROW, T)                                           \
    int NAME(int M, int N, int K, const int *rowptr, const int *colidx,        \
             const T *val, T alpha, const T *B, T beta, T *C) {                \
        T *psum = (T *)malloc(sizeof(T) * (size_t)(N > 0 ? N : 1));            \
        if (!psum) return SX_ORACLE_ENOMEM;                                    \
        for (int64_t i = 0; i < M; ++i)                                        \
            ROW(i, N, M, K, rowptr, colidx, val, alpha, B, beta, C, psum);     \
        free(psum);                                                            \
        return SX_ORACLE_OK;                                                   \
    }

SX_DEFINE_SPMM(sx_oracle_spmm_csr_f32, sx_row_f32, float)
SX_DEFINE_SPMM(sx_oracle_spmm_csr_f64, sx_row_f64, double)

/* Row-parallel variant: rows are independent (sparse_helper.h:279), so running
 * them on several host threads changes no row's arithmetic -- the result is
 * bitwise equal to the single-threaded one.  This is the "all host cores"
 * baseline; it is NOT how the reference runs (the reference is one thread). */
#define SX_DEFINE_SPMM_MT(NAME, ROW, T)                                        \
    int NAME(int M, int N, int K, const int *rowptr, const int *colidx,        \
             const T *val, T alpha, const T *B, T beta, T *C, int threads) {   \
        int failed = 0;                                                        \
        _Pragma("omp parallel num_threads(threads > 0 ? threads : 1)")         \
        {                                                                      \
            T *psum = (T *)malloc(sizeof(T) * (size_t)(N > 0 ? N : 1));        \
            if (!psum) {                                                       \
                _Pragma("omp atomic write") failed = 1;                        \
            } else {                                                           \
                _Pragma("omp for schedule(dynamic, 256)")                      \
                for (int64_t i = 0; i < M; ++i)                                \
                    ROW(i, N, M, K, rowptr, colidx, val, alpha, B, beta, C,    \
                        psum);                                                 \
                free(psum);                                                    \
            }                                                                  \
        }                                                                      \
        return failed ? SX_ORACLE_ENOMEM : SX_ORACLE_OK;                       \
    }

SX_DEFINE_SPMM_MT(sx_oracle_spmm_csr_mt_f32, sx_row_f32, float)
SX_DEFINE_SPMM_MT(sx_oracle_spmm_csr_mt_f64, sx_row_f64, double)

/* Row-sample variant for the 1e6-row configurations: computes only the listed
 * rows (same per-row statements), writing out[s*N + n] for sample s, so that a
 * full-size GPU result can be spot-checked in seconds.  C_in is read from the
 * column-major C; C itself is not modified. */
#define SX_DEFINE_SPMM_ROWS(NAME, T)                                           \
    int NAME(int M, int N, int K, const int *rowptr, const int *colidx,        \
             const T *val, T alpha, const T *B, T beta, const T *C,            \
             const int *rows, int nrows, T *out) {                             \
        for (int s = 0; s < nrows; ++s) {                                      \
            const int64_t i = rows[s];                                         \
            if (i < 0 || i >= M) return SX_ORACLE_EINDEX;                      \
            T *o = out + (int64_t)s * N;                                       \
            for (int n = 0; n < N; ++n) o[n] = (T)0;                           \
            for (int64_t j = rowptr[i]; j < rowptr[i + 1]; ++j) {              \
                const T a = val[j];                                            \
                const T *bcol = B + colidx[j];                                 \
                for (int n = 0; n < N; ++n) {                                  \
                    const T prod = a * bcol[(int64_t)K * n];                   \
                    o[n] = o[n] + prod;                                        \
                }                                                              \
            }                                                                  \
            for (int n = 0; n < N; ++n) {                                      \
                const T left = alpha * o[n];                                   \
                const T right = beta * C[i + (int64_t)M * n];                  \
                o[n] = left + right;                                           \
            }                                                                  \
        }                                                                      \
        return SX_ORACLE_OK;                                                   \
    }

SX_DEFINE_SPMM_ROWS(sx_oracle_spmm_csr_rows_f32, float)
SX_DEFINE_SPMM_ROWS(sx_oracle_spmm_csr_rows_f64, double)

int sx_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Dense operand initialisation of the host driver (sextans-host.cpp:100-111) */
/* ------------------------------------------------------------------------- */

/* B[k + K*n] = 1.0 (host.cpp:102); C[m + M*n] = 1.0*(m+1)*(n+1)/M/N evaluated
 * in double and stored as float (host.cpp:109).  The f64 variant keeps the
 * float-rounded value (promoted) so f32 and f64 runs see identical inputs. */
void sx_oracle_init_dense_f32(int M, int K, int N, float *B, float *C) {
    for (int n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) B[k + (int64_t)K * n] = 1.0f;
    for (int n = 0; n < N; ++n)
        for (int64_t m = 0; m < M; ++m)
            C[m + (int64_t)M * n] = (float)(1.0 * (m + 1) * (n + 1) / M / N);
}

void sx_oracle_init_dense_f64(int M, int K, int N, double *B, double *C) {
    for (int n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) B[k + (int64_t)K * n] = 1.0;
    for (int n = 0; n < N; ++n)
        for (int64_t m = 0; m < M; ++m)
            C[m + (int64_t)M * n] =
                (double)(float)(1.0 * (m + 1) * (n + 1) / M / N);
}

/* ------------------------------------------------------------------------- */
/* Verification criterion (sextans-host.cpp:262-289)                          */
/* ------------------------------------------------------------------------- */

/* Counts elements with |a-b| / (min(|a|,|b|) + 1e-4) > 1e-4, all in float as the
 * reference does (fabs on float promotes to double in C; the reference stores
 * the results back into float variables, host.cpp:272-273).  Returns the count;
 * *percent receives 100*count/M/N as float (host.cpp:281).  pass <=> < 2 %. */
int64_t sx_oracle_verify_f32(int64_t count, const float *cpu, const float *dev,
                             int M, int N, float *percent) {
    int64_t mismatch = 0;
    for (int64_t e = 0; e < count; ++e) {
        const float a = cpu[e], b = dev[e];
        const float dff = (float)fabs((double)(a - b));
        const double fa = fabs((double)a), fb = fabs((double)b);
        const float x = (float)((fa < fb ? fa : fb) + 1e-4);
        if (dff / x > 1e-4) ++mismatch;
    }
    if (percent) *percent = (float)(100.0 * (double)mismatch / M / N);
    return mismatch;
}

/* ------------------------------------------------------------------------- */
/* Matrix Market -> COO -> sorted by (col,row) -> CSC -> CSR                  */
/* ------------------------------------------------------------------------- */

typedef struct {
    int r, c;
    double v; /* holds the float value exactly in the f32 path */
} sx_entry;

static int sx_cmp_col_row(const void *pa, const void *pb) {
    /* cmp_by_column_row, sparse_helper.h:50-62 */
    const sx_entry *a = (const sx_entry *)pa, *b = (const sx_entry *)pb;
    if (a->c != b->c) return a->c > b->c ? 1 : -1;
    if (a->r != b->r) return a->r > b->r ? 1 : -1;
    return 0;
}

static void sx_lower(char *s) {
    for (; *s; ++s) *s = (char)tolower((unsigned char)*s);
}

/* Banner: "%%MatrixMarket matrix <coordinate|array> <real|complex|pattern|integer>
 * <general|symmetric|hermitian|skew-symmetric>" (mmio.h:254-337). */
static int sx_read_banner(FILE *f, char code[4]) {
    char line[1025], banner[64], mtx[64], crd[64], dtype[64], sym[64];
    code[0] = code[1] = code[2] = ' ';
    code[3] = 'G';
    if (!fgets(line, sizeof line, f)) return SX_ORACLE_EBANNER;
    if (sscanf(line, "%63s %63s %63s %63s %63s", banner, mtx, crd, dtype, sym) != 5)
        return SX_ORACLE_EBANNER;
    sx_lower(mtx); sx_lower(crd); sx_lower(dtype); sx_lower(sym);
    if (strncmp(banner, "%%MatrixMarket", 14) != 0) return SX_ORACLE_EBANNER;
    if (strcmp(mtx, "matrix") != 0) return SX_ORACLE_EBANNER;
    code[0] = 'M';
    if (!strcmp(crd, "coordinate")) code[1] = 'C';
    else if (!strcmp(crd, "array")) code[1] = 'A';
    else return SX_ORACLE_EBANNER;
    if (!strcmp(dtype, "real")) code[2] = 'R';
    else if (!strcmp(dtype, "complex")) code[2] = 'C';
    else if (!strcmp(dtype, "pattern")) code[2] = 'P';
    else if (!strcmp(dtype, "integer")) code[2] = 'I';
    else return SX_ORACLE_EBANNER;
    if (!strcmp(sym, "general")) code[3] = 'G';
    else if (!strcmp(sym, "symmetric")) code[3] = 'S';
    else if (!strcmp(sym, "hermitian")) code[3] = 'H';
    else if (!strcmp(sym, "skew-symmetric")) code[3] = 'K';
    else return SX_ORACLE_EBANNER;
    return SX_ORACLE_OK;
}

/* Size line: skip '%' comment lines, then "M N nz", tolerating blank lines
 * (mmio.h:339-367). */
static int sx_read_size(FILE *f, int *M, int *K, int *nz) {
    char line[1025];
    do {
        if (!fgets(line, sizeof line, f)) return SX_ORACLE_ESIZE;
    } while (line[0] == '%');
    if (sscanf(line, "%d %d %d", M, K, nz) == 3) return SX_ORACLE_OK;
    for (;;) {
        int got = fscanf(f, "%d %d %d", M, K, nz);
        if (got == EOF) return SX_ORACLE_ESIZE;
        if (got == 3) return SX_ORACLE_OK;
    }
}

/* Shared loader body.  as_double selects "%lg" parsing (the f64 restatement);
 * otherwise "%f" into a float exactly as sparse_helper.h:140.  Entries whose bit
 * pattern is +0 are dropped (sparse_helper.h:143-145); only 'symmetric' mirrors
 * off-diagonal entries, without negation (sparse_helper.h:156-163); duplicates
 * are kept.  Output arrays are malloc'ed; release with sx_oracle_free. */
static int sx_load(const char *path, int as_double, int *M_out, int *K_out,
                   int *nnz_out, int **rowptr_out, int **colidx_out,
                   void **val_out, char code_out[4]) {
    FILE *f = fopen(path, "r");
    if (!f) return SX_ORACLE_EIO;
    char code[4];
    int M = 0, K = 0, nz = 0;
    int rc = sx_read_banner(f, code);
    if (rc == SX_ORACLE_OK) rc = sx_read_size(f, &M, &K, &nz);
    if (rc == SX_ORACLE_OK && code[1] != 'C') rc = SX_ORACLE_ENOTCOORD;
    if (rc == SX_ORACLE_OK && code[2] == 'C') rc = SX_ORACLE_ECOMPLEX;
    if (rc != SX_ORACLE_OK) { fclose(f); return rc; }
    if (code_out) memcpy(code_out, code, 4);

    const int symmetric = code[3] == 'S';
    const int pattern = code[2] == 'P';
    const size_t cap = (size_t)nz * (symmetric ? 2 : 1);
    sx_entry *coo = (sx_entry *)malloc(sizeof(sx_entry) * (cap ? cap : 1));
    if (!coo) { fclose(f); return SX_ORACLE_ENOMEM; }

    size_t n = 0;
    int r = 0, c = 0;
    float vf = 0.0f;
    double vd = 0.0;
    for (int e = 0; e < nz; ++e) {
        /* a failed conversion leaves the previous values in place, exactly as
         * the unchecked fscanf of the reference does */
        if (pattern) {
            if (fscanf(f, "%d %d\n", &r, &c)) {}
            vf = 1.0f; vd = 1.0;
        } else if (as_double) {
            if (fscanf(f, "%d %d %lg\n", &r, &c, &vd)) {}
        } else {
            if (fscanf(f, "%d %d %f\n", &r, &c, &vf)) {}
        }
        int keep;
        if (as_double) { uint64_t bits; memcpy(&bits, &vd, 8); keep = bits != 0; }
        else { uint32_t bits; memcpy(&bits, &vf, 4); keep = bits != 0; }
        if (!keep) continue;
        if (r < 1 || c < 1) { free(coo); fclose(f); return SX_ORACLE_EINDEX; }
        const double v = as_double ? vd : (double)vf;
        coo[n].r = r - 1; coo[n].c = c - 1; coo[n].v = v; ++n;
        if (symmetric && r != c) {
            coo[n].r = c - 1; coo[n].c = r - 1; coo[n].v = v; ++n;
        }
    }
    fclose(f);

    /* sort_by_fn(..., cmp_by_column_row) (sparse_helper.h:65-87,209-210) */
    qsort(coo, n, sizeof(sx_entry), sx_cmp_col_row);

    /* CSC_2_CSR (sparse_helper.h:475-509): counting pass on rows, then a sweep
     * over the column-sorted entries, which leaves each row's columns ascending
     * and equal (row,col) duplicates in their post-sort order. */
    int *rowptr = (int *)calloc((size_t)M + 1, sizeof(int));
    int *colidx = (int *)malloc(sizeof(int) * (n ? n : 1));
    void *val = malloc((as_double ? 8 : 4) * (n ? n : 1));
    int *fill = (int *)calloc((size_t)(M > 0 ? M : 1), sizeof(int));
    if (!rowptr || !colidx || !val || !fill) {
        free(coo); free(rowptr); free(colidx); free(val); free(fill);
        return SX_ORACLE_ENOMEM;
    }
    for (size_t e = 0; e < n; ++e) rowptr[coo[e].r + 1]++;
    for (int i = 0; i < M; ++i) rowptr[i + 1] += rowptr[i];
    for (size_t e = 0; e < n; ++e) {
        const int row = coo[e].r;
        const int pos = rowptr[row] + fill[row]++;
        colidx[pos] = coo[e].c;
        if (as_double) ((double *)val)[pos] = coo[e].v;
        else ((float *)val)[pos] = (float)coo[e].v;
    }
    free(fill);
    free(coo);
    *M_out = M; *K_out = K; *nnz_out = (int)n;
    *rowptr_out = rowptr; *colidx_out = colidx; *val_out = val;
    return SX_ORACLE_OK;
}

int sx_oracle_load_mtx_f32(const char *path, int *M, int *K, int *nnz,
                           int **rowptr, int **colidx, float **val, char code[4]) {
    void *v = NULL;
    int rc = sx_load(path, 0, M, K, nnz, rowptr, colidx, &v, code);
    *val = (float *)v;
    return rc;
}

int sx_oracle_load_mtx_f64(const char *path, int *M, int *K, int *nnz,
                           int **rowptr, int **colidx, double **val, char code[4]) {
    void *v = NULL;
    int rc = sx_load(path, 1, M, K, nnz, rowptr, colidx, &v, code);
    *val = (double *)v;
    return rc;
}

/*
 * Functional model of the Sextans accelerator on its own channel images: what the
 * task graph of src/sextans.cpp:836-984 computes, in the order it computes it.
 *
 *   for every block of 8 output columns             (l_rp loops, :331, :491; P_N decode :52-54)
 *     local_C <- 0                                  (init_C, :497-507)
 *     for every column window w                     (main loops, :349, :511)
 *       local_B <- rows [4096 w, 4096 w + 4096) x the block's 8 columns of B, taken from
 *                  the 4 B channel images           (read_B, :353-381; layout host.cpp:158-171)
 *       for every slot of the window, for every PE  (computation, :385-420, :516-540)
 *         word = col14 | row18 | fp32; bit 17 of the row field set: bubble, skipped (:404)
 *         abvec[d] = val * local_B[d][col]          (PEcore_Bmtx, :285-295: one rounding)
 *         local_C[PE][row][d] += abvec[d]           (PU2core_Cmtx, :425-447: one rounding)
 *     C_out = alpha * local_C + beta * C_in         (FloatvMultConst x2, FloatvAddFloatv,
 *                                                    :196-233: three roundings)
 *     written as whole 16-row words into the 8 C channel images (:158-194; layout
 *     host.cpp:181-195: channel m % 8, element roundup(M,16)*(n/8) + (m/8)*8 + n % 8)
 *
 * PE p reads word bitrev3(p / 8) of every 8-word slot of channel p % 8
 * (src/sparse_helper.h:451-464 with Scatter_1_2, src/sextans.cpp:785-800) and owns the
 * rows r with r % 64 == p at local address r / 64 (src/sparse_helper.h:314,370).
 * rp_time repeats recompute the same thing from the same C_in (:143) and are not modelled.
 * Returns 0, or 7 (out of memory) / 8 (a word addresses a row >= roundup(M,64) or a column
 * beyond its window's part of K).
 */
int sx_oracle_sextans_images(const int32_t *ptr, const uint64_t *const A[8], const float *const B[4],
                             const float *const Cin[8], float *const Cout[8], int NUM_ITE, int M, int K,
                             int P_N, int alpha_u, int beta_u) {
    const int N = P_N & 0xFFFF;
    const int nblocks = (N + 7) >> 3;
    float alpha, beta;
    memcpy(&alpha, &alpha_u, 4);
    memcpy(&beta, &beta_u, 4);
    const int rows_per_pe = (M + 63) / 64;
    const long colsz_b = (long)((K + 7) / 8) * 16; /* floats per 8-column block in a B image */
    const long colsz_c = (long)((M + 15) / 16) * 16;
    const int Mr = (int)colsz_c;
    float *local_C = (float *)malloc(sizeof(float) * 64 * (size_t)(rows_per_pe ? rows_per_pe : 1) * 8);
    float *local_B = (float *)malloc(sizeof(float) * 8 * 4096);
    if (!local_C || !local_B) { free(local_C); free(local_B); return 7; }
    int rc = 0;
    for (int nb = 0; nb < nblocks && !rc; ++nb) {
        memset(local_C, 0, sizeof(float) * 64 * (size_t)(rows_per_pe ? rows_per_pe : 1) * 8);
        for (int w = 0; w < NUM_ITE && !rc; ++w) {
            const int k0 = w * 4096;
            const int kn = K - k0 < 4096 ? K - k0 : 4096;
            for (int kk = 0; kk < kn; ++kk) {
                const long word = (long)((k0 + kk) / 8) * 16 + (k0 + kk) % 8 + colsz_b * nb;
                for (int d = 0; d < 8; ++d) local_B[d * 4096 + kk] = B[d / 2][word + (d % 2) * 8];
            }
            for (long s = ptr[w]; s < ptr[w + 1] && !rc; ++s) {
                for (int pe = 0; pe < 64; ++pe) {
                    const int q = pe / 8;
                    const int pos = ((q & 1) << 2) | (q & 2) | ((q >> 2) & 1);
                    const uint64_t x = A[pe % 8][s * 8 + pos];
                    const uint32_t row = (uint32_t)(x >> 32) & 0x3FFFFu;
                    if (row & 0x20000u) continue;
                    const uint32_t col = (uint32_t)(x >> 50);
                    if ((int)row >= rows_per_pe || (int)col >= kn) { rc = 8; break; }
                    const uint32_t vb = (uint32_t)x;
                    float val;
                    memcpy(&val, &vb, 4);
                    float *c = local_C + ((size_t)pe * rows_per_pe + row) * 8;
                    for (int d = 0; d < 8; ++d) {
                        const float ab = val * local_B[d * 4096 + col];
                        c[d] = c[d] + ab;
                    }
                }
            }
        }
        for (int m = 0; m < Mr && !rc; ++m) {
            const int pe = m % 64, r = m / 64;
            for (int d = 0; d < 8; ++d) {
                const long p = colsz_c * nb + (long)(m / 8) * 8 + d;
                const float acc = (r < rows_per_pe) ? local_C[((size_t)pe * rows_per_pe + r) * 8 + d] : 0.0f;
                const float t1 = alpha * acc;
                const float t2 = beta * Cin[m % 8][p];
                Cout[m % 8][p] = t1 + t2;
            }
        }
    }
    free(local_C);
    free(local_B);
    return rc;
}

void sx_oracle_free(void *p) { free(p); }
