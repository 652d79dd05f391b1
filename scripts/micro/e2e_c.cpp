// e2e_c.cpp -- the host-facing call timed from C (no Python around it): sx_spmm_f64 on a Matrix Market file
// with page-locked B and C, kernel_ns = NULL, std::chrono around the blocking call, median of 300.
//   e2e_c <path.mtx> [N=16] [host_groups=0] [host_fused=-1]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sextans_b200.h"

#define SXC(x) do { int rc_ = (x); if (rc_) { printf("%s: %s\n", #x, sx_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { printf("usage: e2e_c file.mtx [N] [host_groups] [host_fused]\n"); return 2; }
    const int N = argc > 2 ? atoi(argv[2]) : 16;
    int M, K; int64_t nnz; int32_t *rp, *ci; double *v;
    SXC(sx_load_mtx_f64(argv[1], &M, &K, &nnz, &rp, &ci, &v));
    sx_ctx *c;
    SXC(sx_create(0, &c));
    if (argc > 3) SXC(sx_set_option(c, SX_OPT_HOST_GROUPS, atoi(argv[3])));
    if (argc > 4) SXC(sx_set_option(c, SX_OPT_HOST_FUSED, atoi(argv[4])));
    SXC(sx_upload_csr_f64(c, M, K, nnz, rp, ci, v));
    double *B, *C, *C0;
    SXC(sx_host_alloc((size_t)K * N * 8, (void **)&B));
    SXC(sx_host_alloc((size_t)M * N * 8, (void **)&C));
    C0 = (double *)malloc((size_t)M * N * 8);
    for (size_t i = 0; i < (size_t)K * N; ++i) B[i] = 1.0;
    for (int n = 0; n < N; ++n) for (int m = 0; m < M; ++m) C0[(size_t)n * M + m] = (float)(1.0 * (m + 1) * (n + 1) / M / N);
    std::vector<double> t;
    double sum = 0;
    for (int it = 0; it < 330; ++it) {
        memcpy(C, C0, (size_t)M * N * 8);
        auto t0 = std::chrono::steady_clock::now();
        SXC(sx_spmm_f64(c, N, 0.85f, B, -2.06f, C, 1, nullptr));
        auto t1 = std::chrono::steady_clock::now();
        if (it >= 30) t.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
        if (it == 329) for (size_t i = 0; i < (size_t)M * N; ++i) sum += C[i];
    }
    std::sort(t.begin(), t.end());
    int64_t path = 0;
    sx_get_info(c, SX_INFO_HOST_PATH, &path);
    printf("%s N=%d groups=%s fused=%s: median %.2f us, min %.2f us, p90 %.2f us per call; host path %lld; checksum %.6f\n", argv[1], N,
           argc > 3 ? argv[3] : "auto", argc > 4 ? argv[4] : "auto", t[t.size() / 2], t[0], t[t.size() * 9 / 10], (long long)path, sum);
    sx_destroy(c);
    return 0;
}
