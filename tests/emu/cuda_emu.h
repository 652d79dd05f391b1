// cuda_emu.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the
// block-level kernels of sextans_b200/csrc/spmm_kernels.cuh on the CPU, so that their
// index arithmetic, shared-memory staging and barrier structure can be checked without a
// GPU (tests/test_kernel_emulation_cpu.py).  One OS thread per CUDA thread, one block at a
// time; __syncthreads is a real barrier; TMA bulk copies are memcpys that complete on an
// emulated mbarrier.  No timing, no memory model, no warp shuffles -- kernels that need
// those (the lane-group kernels) are not emulated; their parity is established on the GPU.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static

struct uint3e { unsigned x = 0, y = 0, z = 0; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline thread_local uint3e threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
inline double2 make_double2(double a, double b) { return {a, b}; }
inline int2 make_int2(int a, int b) { return {a, b}; }

using std::max;
using std::min;

namespace sx_emu {
inline std::barrier<> *cur_barrier = nullptr;
inline unsigned char *cur_dyn_smem = nullptr;
inline unsigned char *dyn_smem() { return cur_dyn_smem; }
// __syncwarp(mask): a real barrier among the lanes named by the mask (the lane groups of the
// staged kernel), one per (warp, mask), created by the first lane that gets there
inline std::mutex warp_mu;
inline std::map<uint64_t, std::unique_ptr<std::barrier<>>> warp_barriers;
inline std::mutex mbar_mu;  // the emulated mbarrier operations are made atomic with one lock

// Run `body` as a grid of `grid` blocks of `block` threads with `smem` bytes of dynamic
// shared memory per block (blocks one after the other).
inline void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body) {
    gridDim = dim3(grid);
    blockDim = dim3(block);
    // exactly `smem` bytes (rounded up to the 128-byte alignment unit), so that an address
    // sanitizer build catches a kernel that runs off the end of its dynamic shared memory
    const size_t bytes = (std::max<size_t>(smem, 1) + 127) / 128 * 128;
    unsigned char *base = static_cast<unsigned char *>(std::aligned_alloc(128, bytes));
    struct Free { unsigned char *p; ~Free() { std::free(p); } } guard{base};
    for (unsigned b = 0; b < grid; ++b) {
        std::barrier<> bar((std::ptrdiff_t)block);
        cur_barrier = &bar;
        cur_dyn_smem = base;
        warp_barriers.clear();
        std::memset(base, 0xCD, bytes);  // shared memory starts out as garbage
        std::vector<std::thread> pool;
        pool.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            pool.emplace_back([&, t, b] {
                threadIdx.x = t;
                blockIdx.x = b;
                body();
                bar.arrive_and_drop();  // a thread that has returned no longer takes part in barriers
            });
        for (auto &th : pool) th.join();
    }
}
}  // namespace sx_emu

inline void __syncthreads() { sx_emu::cur_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    const uint64_t key = ((uint64_t)(threadIdx.x / 32) << 32) | mask;
    std::barrier<> *b;
    {
        std::lock_guard<std::mutex> lk(sx_emu::warp_mu);
        auto &slot = sx_emu::warp_barriers[key];
        if (!slot) slot = std::make_unique<std::barrier<>>((std::ptrdiff_t)__builtin_popcount(mask));
        b = slot.get();
    }
    b->arrive_and_wait();
}
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcv(const T *p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
// warp shuffles: every lane named by the mask posts its value in its warp's mailbox, the
// lanes meet (the same per-(warp, mask) barrier as __syncwarp), read the slot they were asked
// for, and meet again before the mailbox may be overwritten
namespace sx_emu {
struct alignas(16) Slot { unsigned char b[16]; };
inline Slot mailbox[64][32];  // [warp of the block][lane]
template <typename T> inline T exchange(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 16, "shuffle of at most 16 bytes");
    const unsigned warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    std::memcpy(mailbox[warp][lane].b, &v, sizeof(T));
    __syncwarp(mask);
    T out;
    std::memcpy(&out, mailbox[warp][src_lane & 31].b, sizeof(T));
    __syncwarp(mask);
    return out;
}
}  // namespace sx_emu
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = (int)(threadIdx.x % 32);
    return sx_emu::exchange(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const int lane = (int)(threadIdx.x % 32);
    const int src = lane ^ lanemask;
    return sx_emu::exchange(mask, v, (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane);
}
inline long long clock64() { return 0; }
inline void __nanosleep(unsigned) {}
inline void __threadfence() {}
inline void __threadfence_system() {}
inline unsigned atomicAdd(unsigned *p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline unsigned atomicExch(unsigned *p, unsigned v) { return std::atomic_ref<unsigned>(*p).exchange(v); }
inline size_t __cvta_generic_to_shared(const void *p) { return reinterpret_cast<size_t>(p); }
