// gather_bench.cu -- how fast can an SM pull random B rows?  (design evidence, not product)
//
// SpMM moves nnz * N * sizeof(T) bytes of B rows through L2 no matter how A is stored,
// so the row-gather mechanism sets the ceiling.  This measures, for a table of `rows`
// rows of RB bytes (L2-resident or DRAM-sized) and a random index stream:
//   ldg    : per-lane LDG.128 into registers, U rows in flight per lane group, batch-synchronous
//   ldg2   : same with two batches in flight (software pipeline)
//   ldgsts : cp.async 16 B per lane into a per-warp shared-memory ring of D rows
//   bulk   : cp.async.bulk (TMA, UBLKCP) one instruction per row into a per-warp ring, mbarrier per slot
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- ldg: G lanes per row, U rows per batch ----------------------------------------
template <int G, int U, int NB>
__global__ void __launch_bounds__(256) k_ldg(const float4 *__restrict__ tab, const int *__restrict__ idx, int per_group, float4 *out, int rowvec) {
    const int lane = threadIdx.x & 31, lg = lane & (G - 1);
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int *my = idx + grp * per_group;
    float4 acc = make_float4(0, 0, 0, 0);
    float4 b[NB][U];
    // prologue
#pragma unroll
    for (int n = 0; n < NB - 1; ++n)
#pragma unroll
        for (int u = 0; u < U; ++u) b[n][u] = __ldg(tab + (int64_t)__ldg(my + n * U + u) * rowvec + lg);
    for (int j = 0; j < per_group; j += U * NB) {
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int nxt = j + (n + NB - 1) * U;   // batch issued now lands in buffer (n+NB-1)%NB
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = nxt + u < per_group ? nxt + u : per_group - 1;
                b[(n + NB - 1) % NB][u] = __ldg(tab + (int64_t)__ldg(my + p) * rowvec + lg);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc = add4(acc, b[n][u]);
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- ldgsts: per-warp ring of D slots of RB bytes ----------------------------------
template <int G, int D>
__global__ void __launch_bounds__(256) k_ldgsts(const float4 *__restrict__ tab, const int *__restrict__ idx, int per_group, float4 *out, int rowvec) {
    extern __shared__ float4 smem[];
    const int lane = threadIdx.x & 31, lg = lane & (G - 1);
    const int gid_in_cta = threadIdx.x / G;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int *my = idx + grp * per_group;
    float4 *ring = smem + (size_t)gid_in_cta * D * G;   // D slots of G float4
    float4 acc = make_float4(0, 0, 0, 0);
    auto issue = [&](int j) {
        const int p = j < per_group ? j : per_group - 1;
        const float4 *src = tab + (int64_t)__ldg(my + p) * rowvec + lg;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (j % D) * G + lg);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        asm volatile("cp.async.commit_group;");
    };
#pragma unroll 1
    for (int j = 0; j < D - 1; ++j) issue(j);
#pragma unroll 4
    for (int j = 0; j < per_group; ++j) {
        issue(j + D - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
        acc = add4(acc, ring[(j % D) * G + lg]);   // own lane's 16 B: no cross-lane sync needed
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- bulk: cp.async.bulk per row, one mbarrier per slot ------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// one warp walks its index slice; lane 0 issues; rows of RB = G*16 bytes; slots grouped in batches of BT
template <int G, int D, int BT>
__global__ void __launch_bounds__(256) k_bulk(const float4 *__restrict__ tab, const int *__restrict__ idx, int per_warp, float4 *out, int rowvec) {
    extern __shared__ float4 smem[];
    __shared__ uint64_t bars[8][D / BT];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int *my = idx + warp * per_warp;
    float4 *ring = smem + (size_t)w * D * G;
    constexpr int NBAT = D / BT;
    if (lane == 0)
        for (int i = 0; i < NBAT; ++i) mbar_init(&bars[w][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    float4 acc = make_float4(0, 0, 0, 0);
    const int nbatch = per_warp / BT;
    auto issue = [&](int bt) {   // lanes 0..BT-1 each issue one row copy of batch bt
        if (bt >= nbatch) return;
        const int slot = bt % NBAT;
        if (lane == 0) mbar_expect_tx(&bars[w][slot], BT * G * 16);
        __syncwarp();
        if (lane < BT) {
            const int p = bt * BT + lane;
            bulk_g2s(ring + (slot * BT + lane) * G, tab + (int64_t)__ldg(my + p) * rowvec, G * 16, &bars[w][slot]);
        }
    };
    for (int bt = 0; bt < NBAT - 1; ++bt) issue(bt);
    for (int bt = 0; bt < nbatch; ++bt) {
        issue(bt + NBAT - 1);
        const int slot = bt % NBAT;
        mbar_wait(&bars[w][slot], (bt / NBAT) & 1);
        // consume BT rows of G float4: lanes cover (32/G) rows per step
#pragma unroll
        for (int r = 0; r < BT; r += 32 / G) acc = add4(acc, ring[(slot * BT + r) * G + lane]);
        __syncwarp();
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_it(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int G>
void run_case(const char *label, int64_t rows, int64_t ngather) {
    const int rowvec = G;   // float4 per row
    float4 *tab; int *idx; float4 *out;
    CK(cudaMalloc(&tab, rows * rowvec * 16));
    CK(cudaMemset(tab, 0, rows * rowvec * 16));
    std::vector<int> h(ngather);
    uint64_t s = 88172645463325252ull;
    for (auto &x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (int)(s % (uint64_t)rows); }
    CK(cudaMalloc(&idx, ngather * 4 + 4096));
    CK(cudaMemcpy(idx, h.data(), ngather * 4, cudaMemcpyHostToDevice));
    const int threads = 256;
    const int64_t bytes = ngather * rowvec * 16;
    printf("%s: table %.0f MB, %lld gathers of %d B (%.2f GB)\n", label, rows * rowvec * 16 / 1e6, (long long)ngather, rowvec * 16, bytes / 1e9);
    CK(cudaMalloc(&out, (size_t)148 * 64 * 2048 * 16));
    auto report = [&](const char *name, float ms) { printf("   %-28s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    // per-group slice lengths chosen so that the grid is ~148 SMs x 8 CTAs x 4 waves
    {
        const int per_group = 256;
        const int64_t groups = ngather / per_group;
        const int grid = (int)(groups * G / threads);
        report("ldg U=8", time_it([&] { k_ldg<G, 8, 1><<<grid, threads>>>(tab, idx, per_group, out, rowvec); }));
        report("ldg U=4 x2 batches", time_it([&] { k_ldg<G, 4, 2><<<grid, threads>>>(tab, idx, per_group, out, rowvec); }));
        report("ldg U=8 x2 batches", time_it([&] { k_ldg<G, 8, 2><<<grid, threads>>>(tab, idx, per_group, out, rowvec); }));
        report("ldg U=4 x4 batches", time_it([&] { k_ldg<G, 4, 4><<<grid, threads>>>(tab, idx, per_group, out, rowvec); }));
        {
            constexpr int D = 16;
            const int sm = (threads / G) * D * G * 16;
            CK(cudaFuncSetAttribute(k_ldgsts<G, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            report("ldgsts ring D=16", time_it([&] { k_ldgsts<G, D><<<grid, threads, sm>>>(tab, idx, per_group, out, rowvec); }));
        }
        {
            constexpr int D = 8;
            const int sm = (threads / G) * D * G * 16;
            CK(cudaFuncSetAttribute(k_ldgsts<G, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            report("ldgsts ring D=8", time_it([&] { k_ldgsts<G, D><<<grid, threads, sm>>>(tab, idx, per_group, out, rowvec); }));
        }
    }
    {
        const int per_warp = 1024;
        const int64_t warps = ngather / per_warp;
        const int grid = (int)(warps * 32 / threads);
        {
            constexpr int D = 32, BT = 8;
            const int sm = 8 * D * G * 16;
            if (sm <= 200 * 1024) {
                CK(cudaFuncSetAttribute(k_bulk<G, D, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                report("bulk(TMA) D=32 batch 8", time_it([&] { k_bulk<G, D, BT><<<grid, threads, sm>>>(tab, idx, per_warp, out, rowvec); }));
            }
        }
        {
            constexpr int D = 16, BT = 4;
            const int sm = 8 * D * G * 16;
            CK(cudaFuncSetAttribute(k_bulk<G, D, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            report("bulk(TMA) D=16 batch 4", time_it([&] { k_bulk<G, D, BT><<<grid, threads, sm>>>(tab, idx, per_warp, out, rowvec); }));
        }
        {
            constexpr int D = 64, BT = 16;
            const int sm = 8 * D * G * 16;
            if (sm <= 200 * 1024) {
                CK(cudaFuncSetAttribute(k_bulk<G, D, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                report("bulk(TMA) D=64 batch 16", time_it([&] { k_bulk<G, D, BT><<<grid, threads, sm>>>(tab, idx, per_warp, out, rowvec); }));
            }
        }
    }
    CK(cudaFree(tab)); CK(cudaFree(idx)); CK(cudaFree(out));
}

__global__ void k_empty() {}

int main() {
    printf("empty kernel, event to event: %.2f us\n", time_it([] { k_empty<<<148, 256>>>(); }, 20) * 1e3);
    // 128 B rows (N=16 fp64): L2-resident 64 MB table and a 1 GB table
    run_case<8>("rows of 128 B, L2-resident", 500000, 1 << 25);
    run_case<8>("rows of 128 B, DRAM", 8000000, 1 << 25);
    // 512 B rows (N=128 fp32)
    run_case<32>("rows of 512 B, L2-resident", 125000, 1 << 23);
    run_case<32>("rows of 512 B, DRAM", 2000000, 1 << 23);
    // 64 B rows (N=16 fp32)
    run_case<4>("rows of 64 B, L2-resident", 1000000, 1 << 25);
    return 0;
}
