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
// mbarrier word: bit 63 = parity of the phase in progress, bit 62 = armed (the expect_tx
// arrival of this phase has happened), low 32 bits = bytes still outstanding.  A phase
// completes -- the parity flips -- when it is armed and no bytes are outstanding;
// mbar_wait(P) returns once the phase of parity P has completed (try_wait.parity semantics).
__device__ __forceinline__ void mbar_init(uint64_t *bar, int) {
    std::lock_guard<std::mutex> lk(sx_emu::mbar_mu);
    *bar = 0;
}
__device__ __forceinline__ void sx_emu_mbar_settle(uint64_t *bar) {
    if ((*bar >> 62 & 1) && (uint32_t)*bar == 0) *bar = (*bar ^ (1ull << 63)) & ~(1ull << 62);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    std::lock_guard<std::mutex> lk(sx_emu::mbar_mu);
    *bar = ((*bar & ~0xFFFFFFFFull) | (uint32_t)((uint32_t)*bar + bytes)) | (1ull << 62);
    sx_emu_mbar_settle(bar);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (;;) {
        {
            std::lock_guard<std::mutex> lk(sx_emu::mbar_mu);
            if ((uint32_t)(*bar >> 63) != (parity & 1u)) return;
        }
        std::this_thread::yield();
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t) {
    if (bytes % 16 || reinterpret_cast<uintptr_t>(dst) % 16 || reinterpret_cast<uintptr_t>(src) % 16) {
        std::fprintf(stderr, "emu: TMA bulk copy not 16-byte aligned/sized (%u bytes)\n", bytes);
        std::abort();
    }
    std::memcpy(dst, src, bytes);
    std::lock_guard<std::mutex> lk(sx_emu::mbar_mu);
    *bar = (*bar & ~0xFFFFFFFFull) | (uint32_t)((uint32_t)*bar - bytes);
    sx_emu_mbar_settle(bar);
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    if (bytes % 16 || reinterpret_cast<uintptr_t>(src) % 16) {
        std::fprintf(stderr, "emu: L2 bulk prefetch not 16-byte aligned/sized (%u bytes)\n", bytes);
        std::abort();
    }
    // touch the first and last byte so that an address-sanitizer build sees a prefetch of memory the product does not own
    volatile unsigned char sink = static_cast<const unsigned char *>(src)[0];
    if (bytes) sink = static_cast<const unsigned char *>(src)[bytes - 1];
    (void)sink;
}
__device__ __forceinline__ uint64_t policy_evict_last() { return 0; }
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { return std::atomic_ref<const uint32_t>(*p).load(); }
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { std::atomic_ref<uint32_t>(*p).store(v); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc) {
    if (reinterpret_cast<uintptr_t>(smem_dst) % 16 || reinterpret_cast<uintptr_t>(gsrc) % 16) {
        std::fprintf(stderr, "emu: cp.async of 16 bytes not 16-byte aligned\n");
        std::abort();
    }
    std::memcpy(smem_dst, gsrc, 16);
}
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) { return std::atomic_ref<const uint32_t>(*p).load(); }
__device__ __forceinline__ void st_relaxed_sys(uint32_t *p, uint32_t v) { std::atomic_ref<uint32_t>(*p).store(v); }
__device__ __forceinline__ void fence_acq_rel_sys() {}
__device__ __forceinline__ void cp_async_elem(float *smem_dst, const float *gsrc) { *smem_dst = *gsrc; }
__device__ __forceinline__ void cp_async_elem(double *smem_dst, const double *gsrc) { *smem_dst = *gsrc; }
__device__ __forceinline__ void cp_async_commit() {}
__device__ __forceinline__ void cp_async_wait_pending(int n) {
    if (n < 0 || n > 7) { std::fprintf(stderr, "emu: cp.async.wait_group %d\n", n); std::abort(); }
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) { return std::atomic_ref<const uint32_t>(*p).load(); }
__device__ __forceinline__ float4 ld_l2(const float4 *p) { return *p; }
__device__ __forceinline__ double2 ld_l2(const double2 *p) { return *p; }
