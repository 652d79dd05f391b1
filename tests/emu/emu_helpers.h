// emu_helpers.h -- TEST INFRASTRUCTURE: CPU stand-ins for the PTX helper section of
// spmm_kernels.cuh (cache-policy loads/stores, mbarrier, TMA bulk copy).  The test pastes
// this in place of that section when it builds the emulated kernels.
__device__ __forceinline__ uint64_t policy_evict_first() { return 0; }
__device__ __forceinline__ int ld_stream(const int *p, uint64_t) { return *p; }
__device__ __forceinline__ float ld_stream(const float *p, uint64_t) { return *p; }
__device__ __forceinline__ double ld_stream(const double *p, uint64_t) { return *p; }
__device__ __forceinline__ float4 ld_once(const float4 *p, uint64_t) { return *p; }
__device__ __forceinline__ double2 ld_once(const double2 *p, uint64_t) { return *p; }
__device__ __forceinline__ void st_once(float4 *p, const float4 &v, uint64_t) { *p = v; }
__device__ __forceinline__ void st_once(double2 *p, const double2 &v, uint64_t) { *p = v; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)reinterpret_cast<uintptr_t>(p); }
// mbarrier word: bit 63 = armed (the expect_tx arrival has happened), low 32 bits = bytes
// still outstanding.  One phase only, which is all the block-level kernels use.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int) { std::atomic_ref<uint64_t>(*bar).store(0); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    std::atomic_ref<uint64_t> w(*bar);
    w.fetch_add(bytes);
    w.fetch_or(1ull << 63);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t) {
    std::atomic_ref<uint64_t> w(*bar);
    for (;;) {
        const uint64_t v = w.load();
        if ((v >> 63) && (uint32_t)v == 0) return;
        std::this_thread::yield();
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t) {
    if (bytes % 16 || reinterpret_cast<uintptr_t>(dst) % 16 || reinterpret_cast<uintptr_t>(src) % 16) {
        std::fprintf(stderr, "emu: TMA bulk copy not 16-byte aligned/sized (%u bytes)\n", bytes);
        std::abort();
    }
    std::memcpy(dst, src, bytes);
    std::atomic_ref<uint64_t>(*bar).fetch_sub(bytes);
}
