/*
 * tapa.h -- source-compatibility stand-in for the TAPA host API, exactly as far as the
 * UNMODIFIED Sextans host program uses it:
 *
 *   tapa::aligned_allocator<T>                           src/sextans-host.cpp:24,
 *                                                        src/sparse_helper.h:409
 *   tapa::round_up<8>(n)                                 src/sextans-host.cpp:51
 *   tapa::vec_t<float, 16>, tapa::mmap<T>,               src/sextans.h:17-26
 *   tapa::mmaps<T, S>
 *   tapa::read_only_mmap<T>(vector),                     src/sextans-host.cpp:239-243
 *   tapa::read_only_mmaps<T, S>(vector-of-vectors).reinterpret<U>(),
 *   tapa::write_only_mmaps<T, S>(...).reinterpret<U>()
 *   tapa::invoke(Sextans, bitstream, args...) -> elapsed nanoseconds
 *                                                        src/sextans-host.cpp:237-251
 *
 * With this directory on the include path and integration/sextans_kernel_b200.cpp in
 * place of src/sextans.cpp, `g++ src/sextans-host.cpp` builds the reference host program
 * as it is and its one device call lands in libsextans_b200.so.
 * Nothing here is TAPA code; it is a from-scratch shim of the few names above.
 */
#ifndef SEXTANS_B200_COMPAT_TAPA_H
#define SEXTANS_B200_COMPAT_TAPA_H

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

extern "C" double sx_sextans_last_kernel_ns(void); /* libsextans_b200.so */

namespace tapa {

/* page-aligned host memory, like the allocator TAPA hands to XRT */
template <typename T>
struct aligned_allocator {
    using value_type = T;
    aligned_allocator() = default;
    template <typename U>
    aligned_allocator(const aligned_allocator<U> &) {}
    T *allocate(std::size_t n) {
        void *p = nullptr;
        if (n > SIZE_MAX / sizeof(T) || posix_memalign(&p, 4096, n * sizeof(T) ? n * sizeof(T) : 1) != 0)
            throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, std::size_t) { std::free(p); }
    template <typename U>
    bool operator==(const aligned_allocator<U> &) const { return true; }
    template <typename U>
    bool operator!=(const aligned_allocator<U> &) const { return false; }
};

template <int N, typename T>
constexpr T round_up(T x) {
    return (x + (N - 1)) / N * N;
}

template <typename T, int N>
struct vec_t {
    T v[N];
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
};

/* a caller-owned host buffer handed to the kernel: pointer + element count */
template <typename T>
class mmap {
  public:
    mmap() = default;
    mmap(T *ptr, std::size_t size) : ptr_(ptr), size_(size) {}
    template <typename Container,
              typename = typename std::enable_if<!std::is_base_of<mmap<T>, Container>::value>::type>
    explicit mmap(Container &c) : ptr_(c.data()), size_(c.size()) {}
    T *data() const { return ptr_; }
    std::size_t size() const { return size_; }
    template <typename U>
    mmap<U> reinterpret() const {
        return mmap<U>(reinterpret_cast<U *>(ptr_), size_ * sizeof(T) / sizeof(U));
    }

  private:
    T *ptr_ = nullptr;
    std::size_t size_ = 0;
};

template <typename T, int S>
class mmaps {
  public:
    mmaps() = default;
    template <typename Containers,
              typename = typename std::enable_if<!std::is_base_of<mmaps<T, S>, Containers>::value>::type>
    explicit mmaps(Containers &cs) {
        for (int i = 0; i < S; ++i) ch_[i] = mmap<T>(cs[i].data(), cs[i].size());
    }
    mmap<T> &operator[](int i) { return ch_[i]; }
    const mmap<T> &operator[](int i) const { return ch_[i]; }
    template <typename U>
    mmaps<U, S> reinterpret() const {
        mmaps<U, S> out;
        for (int i = 0; i < S; ++i) out[i] = ch_[i].template reinterpret<U>();
        return out;
    }

  private:
    mmap<T> ch_[S];
};

/* direction tags only matter to XRT's DMA; here every buffer is plain host memory */
template <typename T> struct read_only_mmap : mmap<T> { using mmap<T>::mmap; };
template <typename T> struct write_only_mmap : mmap<T> { using mmap<T>::mmap; };
template <typename T> struct read_write_mmap : mmap<T> { using mmap<T>::mmap; };
template <typename T, int S> struct read_only_mmaps : mmaps<T, S> { using mmaps<T, S>::mmaps; };
template <typename T, int S> struct write_only_mmaps : mmaps<T, S> { using mmaps<T, S>::mmaps; };
template <typename T, int S> struct read_write_mmaps : mmaps<T, S> { using mmaps<T, S>::mmaps; };

/* Runs the top-level task and returns the KERNEL time in nanoseconds (for all rp_time
 * repeats), which is what the reference divides by rp_time (src/sextans-host.cpp:252).
 * `bitstream` (env TAPAB) selected FPGA vs software simulation; the B200 engine has one
 * backend, so it is accepted and ignored. */
template <typename Func, typename... Args>
inline int64_t invoke(Func &&f, const std::string &bitstream, Args &&...args) {
    (void)bitstream;
    std::forward<Func>(f)(std::forward<Args>(args)...);
    return static_cast<int64_t>(sx_sextans_last_kernel_ns());
}

}  // namespace tapa

#endif
