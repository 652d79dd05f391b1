// pcie_floor.cu -- what the pieces of a ~60 us host-facing SpMM call cost on this box, as the HOST sees them
// (std::chrono around launch ... cudaStreamSynchronize, median of many): an empty kernel, SM-driven reads of
// page-locked host memory (1.2 MB: B and C_in of nasa4704 N=16 fp64), SM-driven writes (0.6 MB: C), the same
// through the copy engines, reads and writes at the same time (is the link full duplex for SM-driven traffic?),
// and large transfers for the asymptotic rates.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void empty_kernel() {}

// every thread moves 16-byte units, grid-stride, consecutive threads on consecutive addresses
__global__ void __launch_bounds__(256) read_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, int64_t n16) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// UNR independent loads in flight per thread
template <int UNR>
__global__ void __launch_bounds__(256) read_unrolled_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, int64_t n16) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * UNR) {
        int4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) if (i + u * stride < n16) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < UNR; ++u) if (i + u * stride < n16) dst[i + u * stride] = v[u];
    }
}

// the access pattern of the SpMM call: column-major operands, a block owns `rows` consecutive rows of all `ncol`
// columns -- runs of rows * 8 bytes at a stride of ld * 8 bytes; a warp per column, lanes along the rows
__global__ void __launch_bounds__(256) tile_read_kernel(const double *__restrict__ src, double *__restrict__ dst, int ld, int ncol, int rows) {
    const int r0 = blockIdx.x * rows, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < ncol; c += 8)
        for (int r = lane; r < rows && r0 + r < ld; r += 32) dst[(size_t)c * ld + r0 + r] = src[(size_t)c * ld + r0 + r];
}
// the same with 8-byte cp.async into shared memory (LDGSTS over PCIe), then out to dst
__global__ void __launch_bounds__(256) tile_read_cpasync_kernel(const double *__restrict__ src, double *__restrict__ dst, int ld, int ncol, int rows) {
    __shared__ double buf[16 * 64];
    const int r0 = blockIdx.x * rows, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < ncol; c += 8)
        for (int r = lane; r < rows && r0 + r < ld; r += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(buf + c * rows + r)), "l"(src + (size_t)c * ld + r0 + r) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int c = warp; c < ncol; c += 8)
        for (int r = lane; r < rows && r0 + r < ld; r += 32) dst[(size_t)c * ld + r0 + r] = buf[c * rows + r];
}

static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename F>
static double median_us(F f, int reps = 200, int warm = 20) {
    for (int i = 0; i < warm; ++i) f();
    std::vector<double> t(reps);
    for (int i = 0; i < reps; ++i) {
        const double t0 = now_us();
        f();
        t[i] = now_us() - t0;
    }
    std::sort(t.begin(), t.end());
    return t[reps / 2];
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs, PCIe domain:bus %04x:%02x\n", prop.name, prop.multiProcessorCount, prop.pciDomainID, prop.pciBusID);
    const size_t big = (size_t)256 << 20;
    int4 *h_in, *h_out, *d_a, *d_b;
    CK(cudaHostAlloc((void **)&h_in, big, cudaHostAllocMapped));
    CK(cudaHostAlloc((void **)&h_out, big, cudaHostAllocMapped));
    CK(cudaMalloc((void **)&d_a, big));
    CK(cudaMalloc((void **)&d_b, big));
    for (size_t i = 0; i < big / 16; ++i) h_in[i] = make_int4((int)i, 1, 2, 3);
    cudaStream_t s0, s1;
    CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    const int64_t in16 = 1204224 / 16, out16 = 602112 / 16;   // nasa4704 N=16 fp64: B + C_in, C

    printf("\n-- host-visible cost of one call: launch(es) + cudaStreamSynchronize, median of 200, us --\n");
    printf("empty kernel                                  %8.2f\n", median_us([&] { empty_kernel<<<1, 32, 0, s0>>>(); cudaStreamSynchronize(s0); }));
    printf("two empty kernels                             %8.2f\n", median_us([&] { empty_kernel<<<148, 256, 0, s0>>>(); empty_kernel<<<148, 256, 0, s0>>>(); cudaStreamSynchronize(s0); }));
    for (int grid : {74, 148, 296, 592}) {
        printf("SM read  1.2 MB host->device, %3d blocks       %8.2f\n", grid, median_us([&] { read_kernel<<<grid, 256, 0, s0>>>(h_in, d_a, in16); cudaStreamSynchronize(s0); }));
    }
    printf("SM read  1.2 MB, 148 blocks, 4 loads in flight %8.2f\n", median_us([&] { read_unrolled_kernel<4><<<148, 128, 0, s0>>>(h_in, d_a, in16); cudaStreamSynchronize(s0); }));
    printf("SM read  0.6 MB host->device, 148 blocks       %8.2f\n", median_us([&] { read_kernel<<<148, 256, 0, s0>>>(h_in, d_a, out16); cudaStreamSynchronize(s0); }));
    for (int grid : {74, 148, 296}) {
        printf("SM write 0.6 MB device->host, %3d blocks       %8.2f\n", grid, median_us([&] { read_kernel<<<grid, 256, 0, s0>>>(d_b, h_out, out16); cudaStreamSynchronize(s0); }));
    }
    {
        const double *hs = reinterpret_cast<const double *>(h_in);
        double *hd = reinterpret_cast<double *>(h_out), *dd = reinterpret_cast<double *>(d_a);
        const double *ds = reinterpret_cast<const double *>(d_b);
        printf("tile read  0.6 MB (147 blocks x 32 rows x 16 col) %6.2f\n", median_us([&] { tile_read_kernel<<<147, 256, 0, s0>>>(hs, dd, 4704, 16, 32); cudaStreamSynchronize(s0); }));
        printf("tile read  0.6 MB, 8-byte cp.async              %7.2f\n", median_us([&] { tile_read_cpasync_kernel<<<147, 256, 0, s0>>>(hs, dd, 4704, 16, 32); cudaStreamSynchronize(s0); }));
        printf("tile read  1.2 MB (as 32 columns)               %7.2f\n", median_us([&] { tile_read_kernel<<<147, 256, 0, s0>>>(hs, dd, 4704, 32, 32); cudaStreamSynchronize(s0); }));
        printf("tile read  0.6 MB (74 blocks x 64 rows)         %7.2f\n", median_us([&] { tile_read_kernel<<<74, 256, 0, s0>>>(hs, dd, 4704, 16, 64); cudaStreamSynchronize(s0); }));
        printf("tile write 0.6 MB (147 blocks x 32 rows x 16 col) %6.2f\n", median_us([&] { tile_read_kernel<<<147, 256, 0, s0>>>(ds, hd, 4704, 16, 32); cudaStreamSynchronize(s0); }));
        printf("tile write 0.6 MB (74 blocks x 64 rows)         %7.2f\n", median_us([&] { tile_read_kernel<<<74, 256, 0, s0>>>(ds, hd, 4704, 16, 64); cudaStreamSynchronize(s0); }));
    }
    printf("SM read 1.2 MB then SM write 0.6 MB (2 kernels) %7.2f\n", median_us([&] {
        read_kernel<<<148, 256, 0, s0>>>(h_in, d_a, in16); read_kernel<<<148, 256, 0, s0>>>(d_b, h_out, out16); cudaStreamSynchronize(s0); }));
    printf("SM read 1.2 MB || SM write 0.6 MB (2 streams)  %8.2f\n", median_us([&] {
        read_kernel<<<148, 256, 0, s0>>>(h_in, d_a, in16); read_kernel<<<148, 256, 0, s1>>>(d_b, h_out, out16); cudaStreamSynchronize(s0); cudaStreamSynchronize(s1); }));
    printf("memcpy H2D 1.2 MB                              %8.2f\n", median_us([&] { cudaMemcpyAsync(d_a, h_in, 1204224, cudaMemcpyHostToDevice, s0); cudaStreamSynchronize(s0); }));
    printf("memcpy H2D 0.6 MB x2                           %8.2f\n", median_us([&] {
        cudaMemcpyAsync(d_a, h_in, 602112, cudaMemcpyHostToDevice, s0); cudaMemcpyAsync(d_a + out16, h_in + out16, 602112, cudaMemcpyHostToDevice, s0); cudaStreamSynchronize(s0); }));
    printf("memcpy D2H 0.6 MB                              %8.2f\n", median_us([&] { cudaMemcpyAsync(h_out, d_b, 602112, cudaMemcpyDeviceToHost, s0); cudaStreamSynchronize(s0); }));
    printf("memcpy H2D 1.2 MB, kernel, memcpy D2H 0.6 MB   %8.2f\n", median_us([&] {
        cudaMemcpyAsync(d_a, h_in, 1204224, cudaMemcpyHostToDevice, s0); empty_kernel<<<148, 256, 0, s0>>>();
        cudaMemcpyAsync(h_out, d_b, 602112, cudaMemcpyDeviceToHost, s0); cudaStreamSynchronize(s0); }));
    printf("memcpy H2D 1.2 MB || memcpy D2H 0.6 MB         %8.2f\n", median_us([&] {
        cudaMemcpyAsync(d_a, h_in, 1204224, cudaMemcpyHostToDevice, s0); cudaMemcpyAsync(h_out, d_b, 602112, cudaMemcpyDeviceToHost, s1);
        cudaStreamSynchronize(s0); cudaStreamSynchronize(s1); }));
    printf("SM read 1.2 MB, then memcpy D2H 0.6 MB         %8.2f\n", median_us([&] {
        read_kernel<<<148, 256, 0, s0>>>(h_in, d_a, in16); cudaMemcpyAsync(h_out, d_b, 602112, cudaMemcpyDeviceToHost, s0); cudaStreamSynchronize(s0); }));

    printf("\n-- asymptotic rates, 256 MB, device time (events), GB/s --\n");
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto rate = [&](auto f) {
        f();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s0));
        f();
        CK(cudaEventRecord(e1, s0));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        return ms;
    };
    const int64_t b16 = big / 16;
    float ms;
    ms = rate([&] { read_kernel<<<592, 256, 0, s0>>>(h_in, d_a, b16); });
    printf("SM read                    %6.1f\n", big / ms / 1e6);
    ms = rate([&] { read_unrolled_kernel<4><<<592, 256, 0, s0>>>(h_in, d_a, b16); });
    printf("SM read, 4 in flight       %6.1f\n", big / ms / 1e6);
    ms = rate([&] { read_kernel<<<592, 256, 0, s0>>>(d_b, h_out, b16); });
    printf("SM write                   %6.1f\n", big / ms / 1e6);
    ms = rate([&] { cudaMemcpyAsync(d_a, h_in, big, cudaMemcpyHostToDevice, s0); });
    printf("memcpy H2D                 %6.1f\n", big / ms / 1e6);
    ms = rate([&] { cudaMemcpyAsync(h_out, d_b, big, cudaMemcpyDeviceToHost, s0); });
    printf("memcpy D2H                 %6.1f\n", big / ms / 1e6);
    // both directions at once: the write kernel on s1 while the read is timed on s0 (and the reverse)
    ms = rate([&] { read_kernel<<<592, 256, 0, s1>>>(d_b, h_out, b16); read_kernel<<<592, 256, 0, s0>>>(h_in, d_a, b16); });
    printf("SM read while SM write     %6.1f (read side)\n", big / ms / 1e6);
    ms = rate([&] { read_kernel<<<592, 256, 0, s1>>>(h_in, d_a, b16); read_kernel<<<592, 256, 0, s0>>>(d_b, h_out, b16); });
    printf("SM write while SM read     %6.1f (write side)\n", big / ms / 1e6);
    ms = rate([&] { cudaMemcpyAsync(h_out, d_b, big, cudaMemcpyDeviceToHost, s1); cudaMemcpyAsync(d_a, h_in, big, cudaMemcpyHostToDevice, s0); });
    printf("memcpy H2D while D2H       %6.1f (H2D side)\n", big / ms / 1e6);
    ms = rate([&] { cudaMemcpyAsync(h_out, d_b, big, cudaMemcpyDeviceToHost, s1); read_kernel<<<592, 256, 0, s0>>>(h_in, d_a, b16); });
    printf("SM read while memcpy D2H   %6.1f (read side)\n", big / ms / 1e6);
    return 0;
}
