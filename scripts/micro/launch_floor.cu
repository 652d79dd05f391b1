// launch_floor.cu -- what a chain of small dependent kernels costs on this GPU, per graph node:
// empty kernels, with and without programmatic dependent launch (PDL), with dependent chains of
// COLD global loads of depth 0..3 per block (the record -> run list -> data chain of the SpMM
// kernels), with a TMA bulk copy of a cold 16 / 64 KB piece, with and without a large dynamic
// shared-memory footprint.  Tells which part of a ~4 us SpMM step on nasa4704 is launch
// machinery and which is memory latency.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// DEPTH dependent loads (pointer chase through `next`), then optionally a TMA bulk copy of `tma_bytes`
// from `data` at an offset derived from the chase; everything after griddepcontrol.wait when PDL.
template <bool PDL>
__global__ void chain_kernel(const uint32_t *__restrict__ next, const unsigned char *__restrict__ data, int depth,
                             uint32_t tma_bytes, uint32_t start_salt, uint32_t nslots, uint32_t *sink, int prologue_depth) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    uint32_t idx = (blockIdx.x * 2654435761u + start_salt) % nslots;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // loads that do not depend on the previous kernel (the A side of an SpMM): before the wait
    for (int d = 0; d < prologue_depth; ++d) idx = __ldg(next + (size_t)idx * 32);
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int d = 0; d < depth; ++d) idx = __ldg(next + (size_t)idx * 32);   // one 128-byte line per hop
    if (tma_bytes && threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(tma_bytes) : "memory");
        const unsigned char *src = data + (size_t)(idx % (nslots / 1024)) * 65536;
        for (uint32_t o = 0; o < tma_bytes; o += 16384u)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + o)), "l"(src + o), "r"(min(16384u, tma_bytes - o)), "r"(smem_u32(&bar)) : "memory");
    }
    if (tma_bytes) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                     ::"r"(smem_u32(&bar)) : "memory");
        idx += smem[threadIdx.x];
    }
    if (idx == 0xffffffffu) *sink = idx;   // never true; keeps the loads alive
}

struct Result { float us; };

template <bool PDL>
float run(cudaStream_t st, const uint32_t *next, const unsigned char *data, uint32_t nslots, uint32_t *sink, int grid, int block,
          size_t smem, int depth, uint32_t tma_bytes, int prologue_depth, int nodes, int reps) {
    auto kern = chain_kernel<PDL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < nodes; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = PDL ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, kern, next, data, depth, tma_bytes, (uint32_t)(i * 7919u + 13u), nslots, sink, prologue_depth));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGraphExecDestroy(ge));
    CK(cudaGraphDestroy(g));
    return ms * 1000.f / (reps * nodes);
}

int main() {
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    // 1 GiB chase table (one 128-byte line per slot) and 1 GiB of data: nothing stays in the 126 MB L2
    const uint32_t nslots = 8u << 20;
    uint32_t *next, *sink;
    unsigned char *data;
    CK(cudaMalloc(&next, (size_t)nslots * 128));
    CK(cudaMalloc(&data, (size_t)(nslots / 1024) * 65536 + 65536));
    CK(cudaMalloc(&sink, 4));
    std::vector<uint32_t> h((size_t)nslots * 32, 0);
    uint64_t x = 88172645463325252ull;
    for (uint32_t i = 0; i < nslots; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[(size_t)i * 32] = (uint32_t)(x % nslots); }
    CK(cudaMemcpy(next, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(data, 1, (size_t)(nslots / 1024) * 65536 + 65536));
    printf("%-64s %8s %8s\n", "per graph node, us (148 blocks x 256 threads)", "no PDL", "PDL");
    struct Case { const char *name; size_t smem; int depth; uint32_t tma; int pro; int grid; };
    const Case cases[] = {
        {"empty", 0, 0, 0, 0, 148},
        {"empty, 100 KB dynamic smem", 100 * 1024, 0, 0, 0, 148},
        {"1 cold load", 0, 1, 0, 0, 148},
        {"2 dependent cold loads", 0, 2, 0, 0, 148},
        {"3 dependent cold loads", 0, 3, 0, 0, 148},
        {"2 cold loads in the prologue (before the wait), 0 after", 0, 0, 0, 2, 148},
        {"2 in the prologue, 1 after", 0, 1, 0, 2, 148},
        {"TMA 16 KB cold", 64 * 1024, 0, 16384, 0, 148},
        {"TMA 64 KB cold", 64 * 1024, 0, 65536, 0, 148},
        {"1 cold load -> TMA 16 KB cold", 64 * 1024, 1, 16384, 0, 148},
        {"2 cold loads -> TMA 16 KB cold", 64 * 1024, 2, 16384, 0, 148},
        {"2 in the prologue -> TMA 16 KB cold after the wait", 64 * 1024, 0, 16384, 2, 148},
        {"1 cold load -> TMA 64 KB cold, 100 KB smem (1 block/SM)", 100 * 1024, 1, 65536, 0, 148},
        {"empty, 296 blocks", 0, 0, 0, 0, 296},
        {"empty, 592 blocks", 0, 0, 0, 0, 592},
        {"2 cold loads -> TMA 16 KB, 444 blocks (3/SM)", 40 * 1024, 2, 16384, 0, 444},
    };
    for (const Case &c : cases) {
        const float a = run<false>(st, next, data, nslots, sink, c.grid, 256, c.smem, c.depth, c.tma, c.pro, 50, 40);
        const float b = run<true>(st, next, data, nslots, sink, c.grid, 256, c.smem, c.depth, c.tma, c.pro, 50, 40);
        printf("%-64s %8.2f %8.2f\n", c.name, a, b);
    }
    return 0;
}
