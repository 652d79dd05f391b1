// sx_images.cpp -- the literal Sextans(...) argument list in front of the engine.
//
// The reference hands its accelerator not A, B and C but their FPGA channel images:
// the hazard-scheduled, bubble-padded 64-bit edge words of 64 PEs interleaved over 8 HBM
// channels (src/sparse_helper.h:345-473), B interleaved over 4 channels and C over 8
// (src/sextans-host.cpp:152-202), plus scalars packed the way the kernel unpacks them
// (src/sextans-host.cpp:221-229, src/sextans.cpp:52-54,200).  sx_sextans_invoke accepts
// exactly that, so a host program can keep ALL of its preprocessing and change nothing
// but the device call.  What happens to the images:
//
//   A   decoded back to CSR on the host, one host thread per group of PEs (a PE owns
//       the rows r % 64 == PE, so PE streams touch disjoint rows and need no locking);
//       walking a PE stream window by window and slot by slot visits each row's nonzeros
//       in the order the FPGA accumulates them -- ascending column, because the hazard
//       scheduler never reorders two edges of one row (src/sparse_helper.h:308-327) --
//       which is the CSR the reference's CSC_2_CSR builds.  The CSR then goes up with
//       sx_upload_csr_f32; the upload is skipped when the images' contents are the ones
//       already uploaded (a hash over every word that matters).
//   B, C_in   un-interleaved to the column-major operands of sx_spmm_f32 in page-locked
//       staging buffers (the engine's device-side transposition takes over from there).
//   C   written back into the 8 output images, whole 16-row words like write_C
//       (src/sextans.cpp:158-194).
//
// The arithmetic is the engine's strict mode: bit-identical to cpu_spmm_CSR and hence
// to the FPGA dataflow for rows up to SX_OPT_SPLIT_ROW_NNZ.
#include "../../include/sextans_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern "C" void sx_internal_set_error(const char *msg);

namespace {

int fail(int status, const std::string &msg) {
    sx_internal_set_error(msg.c_str());
    return status;
}

thread_local double g_last_kernel_ns = 0.0;

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// word position of PE `pe` inside a slot of its channel: bit-reversed pe / 8
// (src/sparse_helper.h:459: (0,4), (2,6), (1,5), (3,7))
inline int slot_word_of_pe(int pe) {
    const int q = pe / 8;
    return ((q & 1) << 2) | (q & 2) | ((q >> 2) & 1);
}

unsigned host_threads(int64_t work) {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    if (const char *e = getenv("SX_HOST_THREADS")) {
        const int v = atoi(e);
        if (v > 0) n = (unsigned)v;
    }
    n = std::min<unsigned>(n, SX_IMAGES_PES);
    if (work < (1 << 16)) n = 1;  // thread start-up costs more than the walk
    return n;
}

template <typename F>
void for_each_pe(unsigned nthreads, F f) {
    if (nthreads <= 1) {
        for (int pe = 0; pe < SX_IMAGES_PES; ++pe) f(pe);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([=] {
            for (int pe = (int)t; pe < SX_IMAGES_PES; pe += (int)nthreads) f(pe);
        });
    for (auto &th : pool) th.join();
}

int check_A_args(const int32_t *ptr, const uint64_t *const ch[], int NUM_ITE, int NUM_A_LEN, int M, int K) {
    if (!ptr || !ch) return fail(SX_ERR_INVALID, "null image pointer");
    if (M < 0 || K < 0 || NUM_ITE < 0 || NUM_A_LEN < 0) return fail(SX_ERR_INVALID, "negative size");
    if (NUM_ITE != (K + SX_IMAGES_WINDOW - 1) / SX_IMAGES_WINDOW)
        return fail(SX_ERR_INVALID, "NUM_ITE = " + std::to_string(NUM_ITE) + " but K = " + std::to_string(K) +
                                        " has " + std::to_string((K + SX_IMAGES_WINDOW - 1) / SX_IMAGES_WINDOW) +
                                        " column windows of 4096");
    if (ptr[0] != 0) return fail(SX_ERR_INVALID, "edge_list_ptr[0] must be 0");
    for (int w = 0; w < NUM_ITE; ++w)
        if (ptr[w + 1] < ptr[w]) return fail(SX_ERR_INVALID, "edge_list_ptr decreases at window " + std::to_string(w));
    if (ptr[NUM_ITE] != NUM_A_LEN)
        return fail(SX_ERR_INVALID, "edge_list_ptr[NUM_ITE] = " + std::to_string(ptr[NUM_ITE]) + " differs from NUM_A_LEN = " +
                                        std::to_string(NUM_A_LEN));
    if (NUM_A_LEN > 0)
        for (int c = 0; c < SX_IMAGES_A_CHANNELS; ++c)
            if (!ch[c]) return fail(SX_ERR_INVALID, "null A channel image");
    return SX_OK;
}

// One walk over PE `pe`'s stream.  visit(row, col, value_bits); returns false on a word
// that cannot have come from the reference's packer.
template <typename V>
bool walk_pe(const int32_t *ptr, const uint64_t *const ch[], int NUM_ITE, int M, int K, int pe, V visit) {
    const uint64_t *img = ch[pe % SX_IMAGES_A_CHANNELS] + slot_word_of_pe(pe);
    for (int w = 0; w < NUM_ITE; ++w) {
        const int64_t base_col = (int64_t)w * SX_IMAGES_WINDOW;
        for (int64_t s = ptr[w]; s < ptr[w + 1]; ++s) {
            const uint64_t x = img[s * 8];
            const uint32_t lrow = (uint32_t)(x >> 32) & 0x3FFFFu;
            if (lrow & 0x20000u) continue;  // bubble: a_row[17] (src/sextans.cpp:404)
            const int64_t row = (int64_t)lrow * SX_IMAGES_PES + pe;
            const int64_t col = base_col + (int64_t)(x >> 50);
            if (row >= M || col >= K) return false;
            visit((int32_t)row, (int32_t)col, (uint32_t)x);
        }
    }
    return true;
}

int decode_A(const int32_t *ptr, const uint64_t *const ch[], int NUM_ITE, int NUM_A_LEN, int M, int K,
             std::vector<int32_t> *rowptr, std::vector<int32_t> *colidx, std::vector<float> *val) {
    int rc = check_A_args(ptr, ch, NUM_ITE, NUM_A_LEN, M, K);
    if (rc) return rc;
    rowptr->assign((size_t)M + 1, 0);
    const unsigned nt = host_threads((int64_t)NUM_A_LEN * SX_IMAGES_PES);
    int bad[SX_IMAGES_PES] = {0};
    int32_t *cnt = rowptr->data() + 1;  // cnt[row] while counting
    for_each_pe(nt, [&](int pe) {
        if (!walk_pe(ptr, ch, NUM_ITE, M, K, pe, [&](int32_t r, int32_t, uint32_t) { ++cnt[r]; })) bad[pe] = 1;
    });
    for (int pe = 0; pe < SX_IMAGES_PES; ++pe)
        if (bad[pe])
            return fail(SX_ERR_INVALID, "A image of PE " + std::to_string(pe) + " addresses a row >= M or a column >= K");
    int64_t total = 0;
    for (int r = 0; r < M; ++r) {
        total += cnt[r];
        if (total > INT32_MAX) return fail(SX_ERR_INVALID, "more than 2^31-1 nonzeros");
        cnt[r] = (int32_t)total;  // rowptr[r+1]
    }
    colidx->resize((size_t)total);
    val->resize((size_t)total);
    std::vector<int32_t> fill(rowptr->begin(), rowptr->end() - 1);  // next free position of every row
    int32_t *ci = colidx->data();
    float *v = val->data();
    for_each_pe(nt, [&](int pe) {
        walk_pe(ptr, ch, NUM_ITE, M, K, pe, [&](int32_t r, int32_t c, uint32_t bits) {
            const int32_t p = fill[r]++;
            ci[p] = c;
            std::memcpy(&v[p], &bits, 4);
        });
    });
    return SX_OK;
}

// FNV-1a over everything that determines the decoded CSR
uint64_t hash_A(const int32_t *ptr, const uint64_t *const ch[], int NUM_ITE, int NUM_A_LEN, int M, int K) {
    uint64_t part[SX_IMAGES_A_CHANNELS];
    const int64_t words = (int64_t)NUM_A_LEN * 8;
    auto run = [&](int c) {
        uint64_t h = 1469598103934665603ull ^ (uint64_t)c;
        const uint64_t *p = ch[c];
        for (int64_t i = 0; i < words; ++i) h = (h ^ p[i]) * 1099511628211ull;
        part[c] = h;
    };
    if (words >= (1 << 16)) {
        std::vector<std::thread> pool;
        for (int c = 0; c < SX_IMAGES_A_CHANNELS; ++c) pool.emplace_back(run, c);
        for (auto &t : pool) t.join();
    } else {
        for (int c = 0; c < SX_IMAGES_A_CHANNELS; ++c) run(c);
    }
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t x) { h = (h ^ x) * 1099511628211ull; };
    mix((uint64_t)M); mix((uint64_t)K); mix((uint64_t)NUM_ITE); mix((uint64_t)NUM_A_LEN);
    for (int w = 0; w <= NUM_ITE; ++w) mix((uint64_t)(uint32_t)ptr[w]);
    for (int c = 0; c < SX_IMAGES_A_CHANNELS; ++c) mix(part[c]);
    return h;
}

// Per-context state of the image path: which images are uploaded, and page-locked
// staging for the column-major operands.
struct ImageState {
    bool valid = false;
    uint64_t hash = 0;
    int64_t upload_serial = -1;
    float *B = nullptr, *C = nullptr;  // sx_host_alloc
    size_t capB = 0, capC = 0;
};
std::mutex g_mu;
std::map<sx_ctx *, ImageState> g_state;

int ensure_pinned(float **p, size_t *cap, size_t elems) {
    if (elems <= *cap && *p) return SX_OK;
    if (*p) { sx_host_free(*p); *p = nullptr; *cap = 0; }
    void *q = nullptr;
    int rc = sx_host_alloc(std::max<size_t>(elems, 1) * sizeof(float), &q);
    if (rc) return rc;
    *p = (float *)q;
    *cap = elems;
    return SX_OK;
}

template <typename T>
T *dup_malloc(const std::vector<T> &v) {
    T *p = (T *)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" {

// called by sx_destroy (sx_api.cu) so that a recycled context address starts clean
void sx_internal_images_forget(sx_ctx *ctx) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_state.find(ctx);
    if (it == g_state.end()) return;
    if (it->second.B) sx_host_free(it->second.B);
    if (it->second.C) sx_host_free(it->second.C);
    g_state.erase(it);
}

int64_t sx_images_A_words(int NUM_A_LEN) { return NUM_A_LEN < 0 ? 0 : (int64_t)NUM_A_LEN * 8; }

int64_t sx_images_B_floats(int K, int N) {
    if (K < 0 || N < 0) return 0;
    return (int64_t)round_up(K, 8) * 2 * (round_up(N, 8) / 8);
}

int64_t sx_images_C_floats(int M, int N) {
    if (M < 0 || N < 0) return 0;
    return (int64_t)round_up(M, 16) * (round_up(N, 8) / 8);
}

double sx_sextans_last_kernel_ns(void) { return g_last_kernel_ns; }

int sx_images_decode_A(const int32_t *edge_list_ptr, const uint64_t *const edge_list_ch[SX_IMAGES_A_CHANNELS],
                       int NUM_ITE, int NUM_A_LEN, int M, int K, int64_t *nnz, int32_t **rowptr, int32_t **colidx,
                       float **val) {
    if (!nnz || !rowptr || !colidx || !val) return fail(SX_ERR_INVALID, "null output pointer");
    *nnz = 0;
    *rowptr = *colidx = nullptr;
    *val = nullptr;
    std::vector<int32_t> rp, ci;
    std::vector<float> v;
    int rc = decode_A(edge_list_ptr, edge_list_ch, NUM_ITE, NUM_A_LEN, M, K, &rp, &ci, &v);
    if (rc) return rc;
    int32_t *prp = dup_malloc(rp), *pci = dup_malloc(ci);
    float *pv = dup_malloc(v);
    if (!prp || !pci || !pv) {
        std::free(prp); std::free(pci); std::free(pv);
        return fail(SX_ERR_NOMEM, "out of host memory");
    }
    *nnz = (int64_t)ci.size();
    *rowptr = prp;
    *colidx = pci;
    *val = pv;
    return SX_OK;
}

int sx_images_decode_B(const float *const mat_B_ch[SX_IMAGES_B_CHANNELS], int K, int N, float *B) {
    if (K < 0 || N < 0) return fail(SX_ERR_INVALID, "negative size");
    const int Nr = round_up(N, 8);
    if ((int64_t)K * Nr == 0) return SX_OK;
    if (!mat_B_ch || !B) return fail(SX_ERR_INVALID, "null pointer");
    for (int c = 0; c < SX_IMAGES_B_CHANNELS; ++c)
        if (!mat_B_ch[c]) return fail(SX_ERR_INVALID, "null B channel image");
    const int64_t colsz = (int64_t)round_up(K, 8) * 2;
    for (int n = 0; n < Nr; ++n) {
        const float *img = mat_B_ch[(n / 2) % 4] + colsz * (n / 8) + (n % 2) * 8;
        float *dst = B + (int64_t)K * n;
        int k = 0;
        for (; k + 8 <= K; k += 8) std::memcpy(dst + k, img + (int64_t)(k / 8) * 16, 32);
        for (; k < K; ++k) dst[k] = img[(int64_t)(k / 8) * 16 + k % 8];
    }
    return SX_OK;
}

int sx_images_decode_C(const float *const mat_C_ch[SX_IMAGES_C_CHANNELS], int M, int N, float *C) {
    if (M < 0 || N < 0) return fail(SX_ERR_INVALID, "negative size");
    const int Nr = round_up(N, 8);
    if ((int64_t)M * Nr == 0) return SX_OK;
    if (!mat_C_ch || !C) return fail(SX_ERR_INVALID, "null pointer");
    for (int c = 0; c < SX_IMAGES_C_CHANNELS; ++c)
        if (!mat_C_ch[c]) return fail(SX_ERR_INVALID, "null C channel image");
    const int64_t colsz = round_up(M, 16);
    for (int n = 0; n < Nr; ++n) {
        float *dst = C + (int64_t)M * n;
        const int64_t off = colsz * (n / 8) + n % 8;
        for (int m = 0; m < M; ++m) dst[m] = mat_C_ch[m % 8][off + (int64_t)(m / 8) * 8];
    }
    return SX_OK;
}

int sx_images_encode_C(const float *C, int M, int N, float alpha, float beta,
                       const float *const mat_C_ch_in[SX_IMAGES_C_CHANNELS], float *const mat_C_ch[SX_IMAGES_C_CHANNELS]) {
    if (M < 0 || N < 0) return fail(SX_ERR_INVALID, "negative size");
    const int Nr = round_up(N, 8);
    if ((int64_t)M * Nr == 0) return SX_OK;
    if (!mat_C_ch || !C) return fail(SX_ERR_INVALID, "null pointer");
    for (int c = 0; c < SX_IMAGES_C_CHANNELS; ++c)
        if (!mat_C_ch[c] || (mat_C_ch_in && !mat_C_ch_in[c])) return fail(SX_ERR_INVALID, "null C channel image");
    const int Mr = round_up(M, 16);
    const int64_t colsz = Mr;
    for (int n = 0; n < Nr; ++n) {
        const float *src = C + (int64_t)M * n;
        const int64_t off = colsz * (n / 8) + n % 8;
        for (int m = 0; m < M; ++m) mat_C_ch[m % 8][off + (int64_t)(m / 8) * 8] = src[m];
        if (mat_C_ch_in) {
            // the rows that pad the last 16-row word: no nonzeros, so the pipeline of
            // src/sextans.cpp:196-233 delivers alpha*0 + beta*C_in (three rounded ops;
            // this file is built with -ffp-contract=off)
            for (int m = M; m < Mr; ++m) {
                const int64_t p = off + (int64_t)(m / 8) * 8;
                const float a = alpha * 0.0f, b = beta * mat_C_ch_in[m % 8][p];
                mat_C_ch[m % 8][p] = a + b;
            }
        }
    }
    return SX_OK;
}

int sx_sextans_invoke(sx_ctx *ctx, const int32_t *edge_list_ptr, const uint64_t *const edge_list_ch[SX_IMAGES_A_CHANNELS],
                      const float *const mat_B_ch[SX_IMAGES_B_CHANNELS], const float *const mat_C_ch_in[SX_IMAGES_C_CHANNELS],
                      float *const mat_C_ch[SX_IMAGES_C_CHANNELS], int NUM_ITE, int NUM_A_LEN, int M, int K, int P_N,
                      int alpha_u, int beta_u, double *kernel_ns) {
    if (kernel_ns) *kernel_ns = 0.0;
    if (!ctx) return fail(SX_ERR_INVALID, "null context");
    const int N = P_N & 0xFFFF;
    const int rp16 = (int)((uint32_t)P_N >> 16);
    const int rp_time = rp16 == 0 ? 1 : rp16;  // src/sextans.cpp:52-54
    if (N < 1) return fail(SX_ERR_INVALID, "P_N carries N = 0");
    const int Nr = round_up(N, 8);  // the accelerator works in blocks of 8 columns (src/sextans.cpp:54)
    float alpha, beta;
    std::memcpy(&alpha, &alpha_u, 4);
    std::memcpy(&beta, &beta_u, 4);
    int rc = check_A_args(edge_list_ptr, edge_list_ch, NUM_ITE, NUM_A_LEN, M, K);
    if (rc) return rc;
    if (!mat_B_ch || !mat_C_ch_in || !mat_C_ch) return fail(SX_ERR_INVALID, "null image pointer");

    ImageState st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        st = g_state[ctx];
    }
    // A: decode + upload unless these very images are what the context holds
    const uint64_t h = hash_A(edge_list_ptr, edge_list_ch, NUM_ITE, NUM_A_LEN, M, K);
    int64_t serial = -1, dtype = -1;
    bool reuse = false;
    if (st.valid && st.hash == h && sx_get_info(ctx, SX_INFO_UPLOAD_SERIAL, &serial) == SX_OK &&
        serial == st.upload_serial && sx_get_info(ctx, SX_INFO_DTYPE, &dtype) == SX_OK && dtype == SX_F32)
        reuse = true;
    if (!reuse) {
        std::vector<int32_t> rp, ci;
        std::vector<float> v;
        if ((rc = decode_A(edge_list_ptr, edge_list_ch, NUM_ITE, NUM_A_LEN, M, K, &rp, &ci, &v))) return rc;
        const float zero = 0.f;
        const int32_t izero = 0;
        st.valid = false;
        if ((rc = sx_upload_csr_f32(ctx, M, K, (int64_t)ci.size(), rp.data(), ci.empty() ? &izero : ci.data(),
                                    v.empty() ? &zero : v.data())))
            return rc;
        if ((rc = sx_get_info(ctx, SX_INFO_UPLOAD_SERIAL, &serial))) return rc;
        st.valid = true;
        st.hash = h;
        st.upload_serial = serial;
    }
    // B, C_in: images -> column-major, page-locked
    if ((rc = ensure_pinned(&st.B, &st.capB, (size_t)K * Nr)) || (rc = ensure_pinned(&st.C, &st.capC, (size_t)M * Nr))) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_state[ctx] = st;
        return rc;
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_state[ctx] = st;
    }
    if ((rc = sx_images_decode_B(mat_B_ch, K, Nr, st.B))) return rc;
    if ((rc = sx_images_decode_C(mat_C_ch_in, M, Nr, st.C))) return rc;
    double ns = 0.0;
    if ((rc = sx_spmm_f32(ctx, Nr, alpha, st.B, beta, st.C, rp_time, &ns))) return rc;
    if ((rc = sx_images_encode_C(st.C, M, Nr, alpha, beta, mat_C_ch_in, mat_C_ch))) return rc;
    g_last_kernel_ns = ns;
    if (kernel_ns) *kernel_ns = ns;
    return SX_OK;
}

}  // extern "C"
