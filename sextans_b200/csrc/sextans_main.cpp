// sextans_main.cpp -- the `sextans` host program on top of libsextans_b200.so.
//
// Keeps the call surface of the reference's main (src/sextans-host.cpp:26-292):
//   sextans <A.mtx> <N>                      (host.cpp:33-48: argc 3)
//   sextans <A.mtx> <N> <rp_time>            (argc 4)
//   sextans <A.mtx> <N> <alpha> <beta>       (argc 5)
//   sextans <A.mtx> <N> <rp_time> <alpha> <beta>   (argc 6)
// the same stdout lines in the same order, the same B / C_in initialisation
// (host.cpp:100-111), the CPU-vs-device verification with the same criterion
// (host.cpp:262-289), and exit status 0 after verification (host.cpp:291).  What is
// gone is the FPGA preprocessing (edge lists, channel images, host.cpp:115-202): A goes
// to the device as CSR and the one device call, tapa::invoke(Sextans, ...) at
// host.cpp:237-251, becomes sx_spmm_f32 / sx_spmm_f64.
//
// Extensions that do not disturb the positional forms (options may appear anywhere):
//   --dtype f32|f64   arithmetic of the device path (default f32, as the reference)
//   --gpus G          row-block partition over G GPUs (one context and one host thread
//                     per GPU; default 1; env SEXTANS_GPUS)
//   --fast            fused multiply-add instead of the bit-exact strict arithmetic
//   --strict-exit     exit status 1 when verification fails (reference: always 0)
//   --json            one machine-readable result line at the end
//
// The CPU loop in check_spmm() below exists for the verification step only, as in the
// reference's main; no result the program reports as the device's ever comes from it.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "../../include/sextans_b200.h"

using std::cout;

namespace {

// golden result for the verification step: per output element, the row's nonzeros in
// stored order, a rounded product then a rounded sum, then alpha*psum + beta*c --
// the arithmetic of src/sparse_helper.h:279-289 (built with -ffp-contract=off)
template <typename T>
void check_spmm(int M, int N, const int32_t *rowptr, const int32_t *colidx, const T *val, T alpha,
                const T *B, int K, T beta, T *C) {
    std::vector<T> psum((size_t)N);
    for (int i = 0; i < M; ++i) {
        std::fill(psum.begin(), psum.end(), T(0));
        for (int32_t j = rowptr[i]; j < rowptr[i + 1]; ++j) {
            const T a = val[j];
            const T *bcol = B + colidx[j];
            for (int nn = 0; nn < N; ++nn) psum[nn] += a * bcol[(size_t)K * nn];
        }
        for (int nn = 0; nn < N; ++nn) {
            T &c = C[i + (size_t)M * nn];
            c = alpha * psum[nn] + beta * c;
        }
    }
}

struct Options {
    std::vector<std::string> pos;
    bool f64 = false, fast = false, strict_exit = false, json = false;
    int gpus = 1;
};

int die(const char *what, int rc) {
    std::fprintf(stderr, "sextans: %s: %s [%s]\n", what, sx_last_error(), sx_status_name(rc));
    return EXIT_FAILURE;
}

template <typename T> struct Api;
template <> struct Api<float> {
    static int load(const char *p, int *M, int *K, int64_t *nnz, int32_t **rp, int32_t **ci, float **v) { return sx_load_mtx_f32(p, M, K, nnz, rp, ci, v); }
    static int upload(sx_ctx *c, int M, int K, int64_t nnz, const int32_t *rp, const int32_t *ci, const float *v) { return sx_upload_csr_f32(c, M, K, nnz, rp, ci, v); }
    static int spmm(sx_ctx *c, int N, float a, const float *B, float b, float *C, int rp, double *ns) { return sx_spmm_f32(c, N, a, B, b, C, rp, ns); }
    static int stage_B(sx_ctx *c, int N, const float *B) { return sx_stage_B_f32(c, N, B); }
    static int stage_C(sx_ctx *c, int N, const float *C) { return sx_stage_C_f32(c, N, C); }
    static int launch(sx_ctx *c, float a, float b, int rp, double *ns) { return sx_launch_f32(c, a, b, rp, ns); }
    static int fetch_C(sx_ctx *c, float *C) { return sx_fetch_C_f32(c, C); }
    static constexpr ncclDataType_t nccl_type = ncclFloat;
};
template <> struct Api<double> {
    static int load(const char *p, int *M, int *K, int64_t *nnz, int32_t **rp, int32_t **ci, double **v) { return sx_load_mtx_f64(p, M, K, nnz, rp, ci, v); }
    static int upload(sx_ctx *c, int M, int K, int64_t nnz, const int32_t *rp, const int32_t *ci, const double *v) { return sx_upload_csr_f64(c, M, K, nnz, rp, ci, v); }
    static int spmm(sx_ctx *c, int N, double a, const double *B, double b, double *C, int rp, double *ns) { return sx_spmm_f64(c, N, a, B, b, C, rp, ns); }
    static int stage_B(sx_ctx *c, int N, const double *B) { return sx_stage_B_f64(c, N, B); }
    static int stage_C(sx_ctx *c, int N, const double *C) { return sx_stage_C_f64(c, N, C); }
    static int launch(sx_ctx *c, double a, double b, int rp, double *ns) { return sx_launch_f64(c, a, b, rp, ns); }
    static int fetch_C(sx_ctx *c, double *C) { return sx_fetch_C_f64(c, C); }
    static constexpr ncclDataType_t nccl_type = ncclDouble;
};

template <typename T>
int run(const Options &opt, const char *filename_A, int N, int rp_time, float ALPHA, float BETA) {
    int M = 0, K = 0;
    int64_t nnz = 0;
    int32_t *rowptr = nullptr, *colidx = nullptr;
    T *val = nullptr;

    cout << "Reading sparse A matrix...";
    int rc = Api<T>::load(filename_A, &M, &K, &nnz, &rowptr, &colidx, &val);
    if (rc) {
        // the reference prints the reason and exits 1 (src/sparse_helper.h:100-109,181-191)
        cout << sx_last_error() << "\n";
        return 1;
    }
    cout << "done\n";

    cout << "Matrix size: \n";
    cout << "A: sparse matrix, " << M << " x " << K << ". NNZ = " << nnz << "\n";
    cout << "B: dense matrix, " << K << " x " << N << "\n";
    cout << "C: dense matrix, " << M << " x " << N << "\n";

    // dense operands in page-locked memory (stands in for tapa::aligned_allocator)
    T *mat_B = nullptr, *mat_C_cpu = nullptr, *mat_C_dev = nullptr;
    const size_t nB = (size_t)K * N, nC = (size_t)M * N;
    if ((rc = sx_host_alloc(nB * sizeof(T), (void **)&mat_B))) return die("sx_host_alloc", rc);
    if ((rc = sx_host_alloc(nC * sizeof(T), (void **)&mat_C_dev))) return die("sx_host_alloc", rc);
    mat_C_cpu = (T *)std::malloc(std::max<size_t>(nC, 1) * sizeof(T));

    cout << "Generating dense matirx B ...";
    for (int nn = 0; nn < N; ++nn)
        for (int kk = 0; kk < K; ++kk) mat_B[kk + (size_t)K * nn] = T(1.0);
    cout << "Generating dense matirx C ...";
    for (int nn = 0; nn < N; ++nn)
        for (int mm = 0; mm < M; ++mm) {
            // the reference stores the double expression into a float (host.cpp:109);
            // the f64 run sees that same float-rounded value
            const float c = (float)(1.0 * (mm + 1) * (nn + 1) / M / N);
            mat_C_cpu[mm + (size_t)M * nn] = (T)c;
            mat_C_dev[mm + (size_t)M * nn] = (T)c;
        }
    cout << "done\n";

    cout << "Preparing sparse A for GPU ...";
    const int G = opt.gpus;
    std::vector<int32_t> bounds((size_t)G + 1);
    if ((rc = sx_partition_rows(M, rowptr, G, bounds.data()))) return die("sx_partition_rows", rc);
    std::vector<sx_ctx *> ctx((size_t)G, nullptr);
    std::vector<std::vector<int32_t>> sub_rowptr((size_t)G);
    for (int g = 0; g < G; ++g) {
        if ((rc = sx_create(g, &ctx[g]))) return die("sx_create", rc);
        if (opt.fast && (rc = sx_set_option(ctx[g], SX_OPT_ARITH, 1))) return die("sx_set_option", rc);
        const int r0 = bounds[g], r1 = bounds[g + 1];
        const int32_t base = rowptr[r0];
        sub_rowptr[g].resize((size_t)(r1 - r0) + 1);
        for (int i = r0; i <= r1; ++i) sub_rowptr[g][i - r0] = rowptr[i] - base;
        if ((rc = Api<T>::upload(ctx[g], r1 - r0, K, rowptr[r1] - base, sub_rowptr[g].data(), colidx + base, val + base)))
            return die("sx_upload_csr", rc);
    }
    cout << "done\n";

    cout << "Run spmm on cpu...";
    auto start_cpu = std::chrono::steady_clock::now();
    check_spmm<T>(M, N, rowptr, colidx, val, (T)ALPHA, mat_B, K, (T)BETA, mat_C_cpu);
    auto end_cpu = std::chrono::steady_clock::now();
    double time_cpu = std::chrono::duration_cast<std::chrono::nanoseconds>(end_cpu - start_cpu).count() * 1e-9;
    cout << "done (" << time_cpu * 1000 << " msec)\n";
    cout << "CPU GFLOPS: " << 2.0f * (nnz + M) * N / 1000000000 / time_cpu << "\n";

    cout << "launch kernel\n";
    // G == 1: the one host-facing call.  G > 1: B goes to GPU 0 once and reaches the other
    // GPUs by ONE ncclBroadcast over NVLink (the daisy chain of src/sextans.cpp:909-941);
    // then one host thread per GPU runs its row block.  A row block of the column-major C
    // is a strided slice, so it is packed into its own column-major buffer around the call.
    std::vector<double> ns((size_t)G, 0.0);
    std::vector<int> status((size_t)G, 0);
    std::vector<std::string> errtext((size_t)G);
    std::vector<std::vector<T>> blockC((size_t)G);
    double bcast_ms = 0.0;
    if (G == 1) {
        status[0] = Api<T>::spmm(ctx[0], N, (T)ALPHA, mat_B, (T)BETA, mat_C_dev, rp_time, &ns[0]);
        if (status[0]) errtext[0] = sx_last_error();
    } else {
        std::vector<ncclComm_t> comm((size_t)G);
        std::vector<int> devs((size_t)G);
        std::vector<cudaStream_t> streams((size_t)G);
        std::vector<void *> dB((size_t)G);
        size_t bytesB = 0;
        for (int g = 0; g < G; ++g) devs[g] = g;
        if (ncclCommInitAll(comm.data(), G, devs.data()) != ncclSuccess) {
            std::fprintf(stderr, "sextans: ncclCommInitAll failed\n");
            return EXIT_FAILURE;
        }
        for (int g = 0; g < G; ++g) {
            cudaSetDevice(g);
            cudaStreamCreateWithFlags(&streams[g], cudaStreamNonBlocking);
            if ((rc = sx_set_stream(ctx[g], streams[g]))) return die("sx_set_stream", rc);
            if (g == 0 && (rc = Api<T>::stage_B(ctx[0], N, mat_B))) return die("sx_stage_B", rc);
            if ((rc = sx_device_B(ctx[g], N, &dB[g], &bytesB))) return die("sx_device_B", rc);
        }
        auto t0 = std::chrono::steady_clock::now();
        ncclGroupStart();
        for (int g = 0; g < G; ++g)
            ncclBroadcast(dB[0], dB[g], bytesB / sizeof(T), Api<T>::nccl_type, 0, comm[g], streams[g]);
        ncclGroupEnd();
        for (int g = 0; g < G; ++g) { cudaSetDevice(g); cudaStreamSynchronize(streams[g]); }
        bcast_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        auto worker = [&](int g) {
            const int r0 = bounds[g], r1 = bounds[g + 1], mb = r1 - r0;
            blockC[g].resize((size_t)mb * N);
            for (int nn = 0; nn < N; ++nn)
                std::memcpy(blockC[g].data() + (size_t)mb * nn, mat_C_dev + r0 + (size_t)M * nn, (size_t)mb * sizeof(T));
            int st = Api<T>::stage_C(ctx[g], N, blockC[g].data());
            if (!st) st = Api<T>::launch(ctx[g], (T)ALPHA, (T)BETA, rp_time, &ns[g]);
            if (!st) st = Api<T>::fetch_C(ctx[g], blockC[g].data());
            status[g] = st;
            if (st) errtext[g] = sx_last_error();
        };
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(worker, g);
        for (auto &t : th) t.join();
        for (int g = 0; g < G; ++g) {
            sx_set_stream(ctx[g], nullptr);
            cudaSetDevice(g);
            cudaStreamDestroy(streams[g]);
            ncclCommDestroy(comm[g]);
        }
        std::printf("B broadcast over NCCL: %.3f ms (%zu bytes to %d GPUs)\n", bcast_ms, bytesB, G - 1);
    }
    for (int g = 0; g < G; ++g)
        if (status[g]) {
            std::fprintf(stderr, "sextans: device %d: %s [%s]\n", g, errtext[g].c_str(), sx_status_name(status[g]));
            return EXIT_FAILURE;
        }
    if (G > 1)
        for (int g = 0; g < G; ++g) {
            const int r0 = bounds[g], mb = bounds[g + 1] - r0;
            for (int nn = 0; nn < N; ++nn)
                std::memcpy(mat_C_dev + r0 + (size_t)M * nn, blockC[g].data() + (size_t)mb * nn, (size_t)mb * sizeof(T));
        }
    double time_taken = *std::max_element(ns.begin(), ns.end());  // the GPUs run side by side
    time_taken *= (1e-9 / rp_time);
    std::printf("Kernel time is %f ms\n", time_taken * 1000);
    float gflops = 2.0 * N * (nnz + M) / 1e9 / time_taken;
    std::printf("GFLOPS:%f \n", gflops);

    // verification, criterion of host.cpp:262-289 (in the arithmetic type of the run)
    int mismatch_cnt = 0;
    double max_rel = 0.0;
    for (int nn = 0; nn < N; ++nn)
        for (int mm = 0; mm < M; ++mm) {
            const T v_cpu = mat_C_cpu[mm + (size_t)nn * M], v_dev = mat_C_dev[mm + (size_t)nn * M];
            const float dff = std::fabs((float)v_cpu - (float)v_dev);
            const float x = std::min(std::fabs((float)v_cpu), std::fabs((float)v_dev)) + 1e-4;
            if (!(dff / x <= 1e-4)) mismatch_cnt++;
            const double rel = std::fabs((double)v_cpu - (double)v_dev) / std::max(std::fabs((double)v_cpu), 1e-30);
            if (rel > max_rel) max_rel = rel;
        }
    float diffpercent = (M && N) ? 100.0 * mismatch_cnt / M / N : 0.f;
    bool pass = diffpercent < 2.0;
    if (pass) cout << "Success!\n";
    else cout << "Failed.\n";
    std::printf("num_mismatch = %d, percent = %.2f%%\n", mismatch_cnt, diffpercent);
    std::printf("max_rel_err = %.3e\n", max_rel);
    if (opt.json) {
        int64_t launches = 0;
        for (int g = 0; g < G; ++g) {
            int64_t l = 0;
            sx_get_info(ctx[g], SX_INFO_LAUNCHES, &l);
            launches += l;
        }
        std::printf("{\"M\": %d, \"K\": %d, \"nnz\": %lld, \"N\": %d, \"dtype\": \"%s\", \"gpus\": %d, \"rp_time\": %d, "
                    "\"kernel_ms\": %.6f, \"gflops_2nnzN\": %.3f, \"gflops_ref_formula\": %.3f, \"cpu_ms\": %.3f, "
                    "\"mismatch\": %d, \"max_rel_err\": %.3e, \"pass\": %s, \"gpu_launches\": %lld}\n",
                    M, K, (long long)nnz, N, sizeof(T) == 8 ? "f64" : "f32", G, rp_time, time_taken * 1000,
                    2.0 * N * nnz / 1e9 / time_taken, (double)gflops, time_cpu * 1000, mismatch_cnt, max_rel,
                    pass ? "true" : "false", (long long)launches);
    }
    for (int g = 0; g < G; ++g) sx_destroy(ctx[g]);
    sx_host_free(mat_B);
    sx_host_free(mat_C_dev);
    std::free(mat_C_cpu);
    sx_free(rowptr);
    sx_free(colidx);
    sx_free(val);
    return (opt.strict_exit && !pass) ? EXIT_FAILURE : EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char **argv) {
    std::printf("start host\n");

    Options opt;
    if (const char *e = std::getenv("SEXTANS_GPUS")) opt.gpus = std::atoi(e);
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--dtype" && i + 1 < argc) opt.f64 = std::string(argv[++i]) == "f64";
        else if (a == "--gpus" && i + 1 < argc) opt.gpus = std::atoi(argv[++i]);
        else if (a == "--fast") opt.fast = true;
        else if (a == "--strict-exit") opt.strict_exit = true;
        else if (a == "--json") opt.json = true;
        else opt.pos.push_back(a);
    }
    float ALPHA = 0.85;
    float BETA = -2.06;
    int rp_time = 1;
    const size_t np = opt.pos.size() + 1;  // the reference's argc
    if (np == 6) {
        rp_time = std::atoi(opt.pos[2].c_str());
        ALPHA = std::atof(opt.pos[3].c_str());
        BETA = std::atof(opt.pos[4].c_str());
    } else if (np == 5) {
        ALPHA = std::atof(opt.pos[2].c_str());
        BETA = std::atof(opt.pos[3].c_str());
    } else if (np == 4) {
        rp_time = std::atoi(opt.pos[2].c_str());
    } else if (np != 3) {
        cout << "Usage: " << argv[0] << " [matrix A file] [N] [rp_time] [alpha] [beta]" << std::endl;
        return EXIT_FAILURE;
    }
    if (opt.gpus < 1) opt.gpus = 1;
    if (rp_time < 1) rp_time = 1;  // the kernel treats 0 as 1 (src/sextans.cpp:52-54)

    const char *filename_A = opt.pos[0].c_str();
    int N = (std::atoi(opt.pos[1].c_str()) + 7) / 8 * 8;  // tapa::round_up<8> (host.cpp:51)
    if (N < 8) N = 8;

    cout << "N = " << N << "\n";
    cout << "alpha = " << ALPHA << "\n";
    cout << "beta = " << BETA << "\n";

    return opt.f64 ? run<double>(opt, filename_A, N, rp_time, ALPHA, BETA)
                   : run<float>(opt, filename_A, N, rp_time, ALPHA, BETA);
}
