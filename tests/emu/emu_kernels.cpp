// emu_kernels.cpp -- TEST INFRASTRUCTURE: runs the block-level SpMM kernels of
// sextans_b200/csrc/spmm_kernels.cuh (variant 3 with 32/64/128-row blocks, with and without
// the PDL code path, and the host-boundary fusion spmm_window_hostc_kernel) on the CPU
// emulation of tests/emu/cuda_emu.h and compares them, bit for bit, with the plain loop of
// cpu_spmm_CSR (src/sparse_helper.h:262-290: stored order, separately rounded * and +;
// built with -ffp-contract=off).  The kernel source is the product's, textually, with only
// its PTX helper section swapped for emu_helpers.h (tests/test_kernel_emulation_cpu.py).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "cuda_emu.h"
#define asm
#define volatile(...)
#include "spmm_kernels_emu.cuh"  // generated next to this file's object by the test

namespace {

template <typename T> struct Aligned {  // 256-byte aligned like cudaMalloc, zero-filled, with slack
    std::vector<unsigned char> raw;
    T *p = nullptr;
    explicit Aligned(size_t n) : raw(n * sizeof(T) + 512, 0) {
        unsigned char *b = raw.data();
        b += (256 - reinterpret_cast<uintptr_t>(b) % 256) % 256;
        p = reinterpret_cast<T *>(b);
    }
};

struct Csr { int M, K; std::vector<int> rp, ci; };

Csr banded(int M, int K, int half_band, int per_row, unsigned seed) {
    std::mt19937 rng(seed);
    Csr a{M, K, std::vector<int>(M + 1, 0), {}};
    for (int r = 0; r < M; ++r) {
        const int centre = (int)((long long)r * K / M);
        const int lo = std::max(0, centre - half_band), hi = std::min(K, centre + half_band + 1);
        std::vector<int> cols;
        const int want = (r % 7 == 3) ? 0 : std::min(per_row + (int)(rng() % 5), hi - lo);
        while ((int)cols.size() < want) {
            const int c = lo + (int)(rng() % (unsigned)(hi - lo));
            if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c);
        }
        std::sort(cols.begin(), cols.end());
        a.ci.insert(a.ci.end(), cols.begin(), cols.end());
        a.rp[r + 1] = (int)a.ci.size();
    }
    return a;
}

// block records {first column, span, nnz begin, nnz end} for blocks of RB rows -- the rule of
// sx_api.cu: build_window_blocks
void block_records(const Csr &a, int RB, std::vector<int> *blk, int *max_span, int *max_nnz) {
    const int nb = (a.M + RB - 1) / RB;
    blk->assign((size_t)nb * 4, 0);
    *max_span = *max_nnz = 0;
    for (int b = 0; b < nb; ++b) {
        const int r0 = b * RB, r1 = std::min(a.M, r0 + RB);
        const int jb = a.rp[r0], je = a.rp[r1];
        int lo = INT32_MAX, hi = -1;
        for (int j = jb; j < je; ++j) { lo = std::min(lo, a.ci[j]); hi = std::max(hi, a.ci[j]); }
        if (je == jb) { lo = 0; hi = -1; }
        (*blk)[(size_t)b * 4 + 0] = lo;
        (*blk)[(size_t)b * 4 + 1] = hi - lo + 1;
        (*blk)[(size_t)b * 4 + 2] = jb;
        (*blk)[(size_t)b * 4 + 3] = je;
        *max_span = std::max(*max_span, hi - lo + 1);
        *max_nnz = std::max(*max_nnz, je - (jb & ~3));
    }
}

template <typename T>
void reference(const Csr &a, const std::vector<T> &val, int N, const T *B, int64_t ldb, T alpha, T beta,
               const T *Cin, T *Cout, int64_t ldc) {
    for (int i = 0; i < a.M; ++i)
        for (int n = 0; n < N; ++n) {
            T psum = 0;
            for (int j = a.rp[i]; j < a.rp[i + 1]; ++j) {
                const T prod = val[j] * B[(int64_t)a.ci[j] * ldb + n];
                psum = psum + prod;
            }
            const T t1 = alpha * psum, t2 = beta * Cin[(int64_t)i * ldc + n];
            Cout[(int64_t)i * ldc + n] = t1 + t2;
        }
}

template <typename T, int G, int RB, bool PDL>
void run_window(const Csr &a, const T *val, const int *blk, int nblk, size_t smem, const T *B, uint32_t ldbv,
                const T *Cin, T *Cout, uint32_t ldcv, T alpha, T beta, int nvec, const int *rp, const int *ci) {
    sx_emu::launch((unsigned)nblk, RB * G, smem, [&] {
        sx::spmm_window_kernel<T, G, true, PDL, RB>(a.M, reinterpret_cast<const int4 *>(blk), rp, ci, val, B, ldbv, Cin,
                                                     Cout, ldcv, alpha, beta, nvec);
    });
}

template <typename T, int G>
void run_hostc(const Csr &a, const T *val, const int *blk, int nblk, size_t smem, uint32_t tile_off, const T *B,
               uint32_t ldbv, T *Ch, int N, T alpha, T beta, int nvec, const int *rp, const int *ci) {
    sx_emu::launch((unsigned)nblk, 32 * G, smem, [&] {
        sx::spmm_window_hostc_kernel<T, G, true>(a.M, reinterpret_cast<const int4 *>(blk), rp, ci, val, B, ldbv, Ch, N,
                                                 alpha, beta, nvec, tile_off);
    });
}

int failures = 0;

template <typename T>
bool same_bits(const T *x, const T *y, size_t n) { return std::memcmp(x, y, n * sizeof(T)) == 0; }

// one case: every kernel flavour that fits, against the reference loop
template <typename T, int G>
void one_case(const char *tname, int M, int K, int N, int half_band, int per_row, unsigned seed) {
    constexpr int E = 16 / (int)sizeof(T);
    const Csr a = banded(M, K, half_band, per_row, seed);
    const int nnz = a.rp[M];
    std::mt19937 rng(seed * 7 + 1);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    Aligned<T> val((size_t)nnz + 16), B((size_t)K * ((N + 7) / 8 * 8)), Cin((size_t)M * ((N + 7) / 8 * 8)),
        Cout((size_t)M * ((N + 7) / 8 * 8)), Ref((size_t)M * ((N + 7) / 8 * 8));
    Aligned<int> ci((size_t)nnz + 16), rp((size_t)M + 1);
    std::vector<T> hval((size_t)nnz);
    for (int j = 0; j < nnz; ++j) { hval[j] = (T)U(rng); val.p[j] = hval[j]; ci.p[j] = a.ci[j]; }
    for (int i = 0; i <= M; ++i) rp.p[i] = a.rp[i];
    const int64_t ld = (N + 7) / 8 * 8;
    for (int64_t i = 0; i < (int64_t)K * ld; ++i) B.p[i] = (i % ld) < N ? (T)U(rng) : (T)0;
    for (int64_t i = 0; i < (int64_t)M * ld; ++i) Cin.p[i] = (i % ld) < N ? (T)U(rng) : (T)0;
    const T alpha = (T)0.85f, beta = (T)-2.06f;
    reference<T>(a, hval, N, B.p, ld, alpha, beta, Cin.p, Ref.p, ld);
    const int nvec = (N * (int)sizeof(T) + 15) / 16;
    const uint32_t ldv = (uint32_t)(ld / E);

    auto check = [&](const char *what, const T *got) {
        bool ok = true;
        for (int i = 0; i < M && ok; ++i) ok = same_bits(got + (int64_t)i * ld, Ref.p + (int64_t)i * ld, (size_t)N);
        std::printf("%-34s %s M=%d K=%d N=%d G=%d: %s\n", what, tname, M, K, N, G, ok ? "bit-exact" : "MISMATCH");
        if (!ok) ++failures;
    };
    auto window = [&](auto runner, int RB, const char *what) {
        std::vector<int> blk;
        int max_span, max_nnz;
        block_records(a, RB, &blk, &max_span, &max_nnz);
        Aligned<int> dblk(blk.size() + 4);
        std::copy(blk.begin(), blk.end(), dblk.p);
        const size_t smem = (size_t)max_span * ldv * 16 + ((size_t)max_nnz + 8) * (sizeof(T) + 4) + 16;
        if (smem > 200 * 1024 || RB * G > 1024) { std::printf("%-34s %s N=%d G=%d: skipped (does not fit)\n", what, tname, N, G); return; }
        std::fill(Cout.p, Cout.p + (int64_t)M * ld, (T)777);
        runner(dblk.p, (int)(blk.size() / 4), smem);
        check(what, Cout.p);
    };
    window([&](const int *blk, int nb, size_t smem) { run_window<T, G, 32, false>(a, val.p, blk, nb, smem, B.p, ldv, Cin.p, Cout.p, ldv, alpha, beta, nvec, rp.p, ci.p); }, 32, "window RB=32");
    window([&](const int *blk, int nb, size_t smem) { run_window<T, G, 32, true>(a, val.p, blk, nb, smem, B.p, ldv, Cin.p, Cout.p, ldv, alpha, beta, nvec, rp.p, ci.p); }, 32, "window RB=32 PDL");
    if constexpr (G >= 4) {
        window([&](const int *blk, int nb, size_t smem) { run_window<T, G, 64, false>(a, val.p, blk, nb, smem, B.p, ldv, Cin.p, Cout.p, ldv, alpha, beta, nvec, rp.p, ci.p); }, 64, "window RB=64");
        if constexpr (G <= 8)
            window([&](const int *blk, int nb, size_t smem) { run_window<T, G, 128, false>(a, val.p, blk, nb, smem, B.p, ldv, Cin.p, Cout.p, ldv, alpha, beta, nvec, rp.p, ci.p); }, 128, "window RB=128");
    }
    // host-boundary fusion: C column-major, in place
    if (M % E == 0) {
        std::vector<int> blk;
        int max_span, max_nnz;
        block_records(a, 32, &blk, &max_span, &max_nnz);
        Aligned<int> dblk(blk.size() + 4);
        std::copy(blk.begin(), blk.end(), dblk.p);
        const size_t wsmem = (size_t)max_span * ldv * 16 + ((size_t)max_nnz + 8) * (sizeof(T) + 4) + 16;
        const size_t tile_off = (wsmem + 15) & ~(size_t)15;
        const size_t smem = tile_off + (size_t)nvec * E * (32 + E) * sizeof(T);
        Aligned<T> Ch((size_t)M * N + 16);
        for (int i = 0; i < M; ++i)
            for (int n = 0; n < N; ++n) Ch.p[(size_t)M * n + i] = Cin.p[(int64_t)i * ld + n];
        run_hostc<T, G>(a, val.p, dblk.p, (int)(blk.size() / 4), smem, (uint32_t)tile_off, B.p, ldv, Ch.p, N, alpha, beta, nvec, rp.p, ci.p);
        for (int i = 0; i < M; ++i)
            for (int n = 0; n < N; ++n) Cout.p[(int64_t)i * ld + n] = Ch.p[(size_t)M * n + i];
        check("hostc (C column-major, in place)", Cout.p);
        bool clean = true;  // nothing written past the M*N array
        for (int i = 0; i < 16; ++i) clean = clean && Ch.p[(size_t)M * N + i] == (T)0;
        if (!clean) { std::printf("hostc wrote past the end of C\n"); ++failures; }
    }
}

template <typename T>
void by_shape(const char *tname, int M, int K, int N, int half_band, int per_row, unsigned seed) {
    const int nvec = (N * (int)sizeof(T) + 15) / 16;
    int G = 2;
    while (G < 32 && G < nvec) G <<= 1;
    switch (G) {
        case 2: one_case<T, 2>(tname, M, K, N, half_band, per_row, seed); break;
        case 4: one_case<T, 4>(tname, M, K, N, half_band, per_row, seed); break;
        case 8: one_case<T, 8>(tname, M, K, N, half_band, per_row, seed); break;
        case 16: one_case<T, 16>(tname, M, K, N, half_band, per_row, seed); break;
        default: std::printf("N=%d needs more than 16 lanes per row: not a variant-3 shape\n", N);
    }
}

}  // namespace

int main() {
    const struct { int M, K, N, hb, per; } cases[] = {
        {200, 200, 16, 40, 9}, {130, 150, 8, 30, 6}, {96, 96, 4, 20, 5}, {257, 300, 24, 50, 11}, {64, 64, 32, 30, 7},
        {300, 280, 3, 25, 4}, {128, 128, 64, 20, 6}};
    unsigned seed = 1;
    for (const auto &c : cases) {
        by_shape<float>("f32", c.M, c.K, c.N, c.hb, c.per, seed++);
        by_shape<double>("f64", c.M, c.K, c.N, c.hb, c.per, seed++);
    }
    std::printf(failures ? "EMULATION: %d FAILURES\n" : "EMULATION: all bit-exact\n", failures);
    return failures ? 1 : 0;
}
