// oracle/ref_shim.cpp -- C entry points around the UNMODIFIED reference header.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spmm_oracle.c).  This file contains no
// reference code: it includes /root/reference/src/sparse_helper.h (and through it
// mmio.h) where they lie, via -I on the compile line (oracle/Makefile), and the
// output goes to oracle/_ref/libsextans_ref.so, which is git-ignored.
//
// The only thing the header needs from TAPA is the name
// tapa::aligned_allocator in one parameter type (src/sparse_helper.h:409); the
// alias below supplies it.  The FPGA kernel itself (src/sextans.cpp) needs
// <tapa.h>/<ap_int.h>, which are not in this image, so it is not built
// (DESIGN.md "what is unbuildable").
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace tapa {
template <class T>
using aligned_allocator = std::allocator<T>;
}

#include "sparse_helper.h"

namespace {
template <class T>
T *dup(const std::vector<T> &v) {
    T *p = static_cast<T *>(std::malloc(sizeof(T) * (v.size() ? v.size() : 1)));
    if (p && !v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}
}  // namespace

extern "C" {

// read_suitsparse_matrix(..., CSC) followed by CSC_2_CSR, the exact sequence of
// src/sextans-host.cpp:67-84.  Loader errors exit(1) inside the reference.
int sxref_load_csr(const char *path, int *M, int *K, int *nnz, int **rowptr,
                   int **colidx, float **val) {
    std::vector<int> cptr, ridx, rptr, cidx;
    std::vector<float> cval, rval;
    read_suitsparse_matrix(const_cast<char *>(path), cptr, ridx, cval, *M, *K, *nnz, CSC);
    CSC_2_CSR(*M, *K, *nnz, cptr, ridx, cval, rptr, cidx, rval);
    *rowptr = dup(rptr);
    *colidx = dup(cidx);
    *val = dup(rval);
    return 0;
}

// cpu_spmm_CSR on caller arrays (copied into the std::vectors the reference
// signature wants).  Returns the nanoseconds spent inside cpu_spmm_CSR only,
// measured like src/sextans-host.cpp:207-217.
double sxref_cpu_spmm_csr(int M, int N, int K, int nnz, float alpha, const int *rowptr,
                          const int *colidx, const float *val, const float *B, float beta,
                          float *C) {
    std::vector<int> rp(rowptr, rowptr + M + 1), ci(colidx, colidx + nnz);
    std::vector<float> v(val, val + nnz), b(B, B + (size_t)K * N), c(C, C + (size_t)M * N);
    auto t0 = std::chrono::steady_clock::now();
    cpu_spmm_CSR(M, N, K, nnz, alpha, rp, ci, v, b, beta, c);
    auto t1 = std::chrono::steady_clock::now();
    std::memcpy(C, c.data(), sizeof(float) * c.size());
    return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
}

// The FPGA preprocessing, exposed only so tests can assert the property the
// survey relies on: per-row column order survives the hazard scheduler
// (src/sparse_helper.h:345-403).  Returns ptr.back() (padded slots per PE).
int sxref_edge_list_slots(int M, int K, int nnz, const int *rowptr_csr, const int *colidx_csr,
                          const float *val_csr, int num_pe, int window, int dep_dist) {
    // rebuild CSC from CSR (rows ascending within a column, as the loader gives)
    std::vector<int> cptr(K + 1, 0), ridx(nnz);
    std::vector<float> cval(nnz);
    for (int j = 0; j < nnz; ++j) cptr[colidx_csr[j] + 1]++;
    for (int k = 0; k < K; ++k) cptr[k + 1] += cptr[k];
    std::vector<int> fill(K, 0);
    for (int i = 0; i < M; ++i)
        for (int j = rowptr_csr[i]; j < rowptr_csr[i + 1]; ++j) {
            int k = colidx_csr[j];
            int pos = cptr[k] + fill[k]++;
            ridx[pos] = i;
            cval[pos] = val_csr[j];
        }
    std::vector<std::vector<edge>> pes;
    std::vector<int> ptr;
    generate_edge_list_for_all_PEs(cptr, ridx, cval, num_pe, M, K, window, pes, ptr, dep_dist);
    return ptr.empty() ? 0 : ptr.back();
}

// The complete A side of "Preparing sparse A for FPGA" (src/sextans-host.cpp:117-146):
// generate_edge_list_for_all_PEs + edge_list_64bit with the shipped constants
// (src/sextans.h:7-12), from a CSR whose rows hold ascending columns.  Outputs are
// malloc'ed: ptr (ptr_len ints, unpadded), 8 channel images of image_len 64-bit words.
// Lets the tests feed the engine's image path (sx_sextans_invoke) the reference's own images.
int sxref_build_images(int M, int K, int nnz, const int *rowptr_csr, const int *colidx_csr,
                       const float *val_csr, int **ptr, int *ptr_len, unsigned long **images /*[8]*/,
                       long *image_len) {
    std::vector<int> cptr(K + 1, 0), ridx(nnz);
    std::vector<float> cval(nnz);
    for (int j = 0; j < nnz; ++j) cptr[colidx_csr[j] + 1]++;
    for (int k = 0; k < K; ++k) cptr[k + 1] += cptr[k];
    std::vector<int> fill(K, 0);
    for (int i = 0; i < M; ++i)
        for (int j = rowptr_csr[i]; j < rowptr_csr[i + 1]; ++j) {
            int k = colidx_csr[j];
            int pos = cptr[k] + fill[k]++;
            ridx[pos] = i;
            cval[pos] = val_csr[j];
        }
    std::vector<std::vector<edge>> pes;
    std::vector<int> eptr;
    generate_edge_list_for_all_PEs(cptr, ridx, cval, 8 * 8, M, K, 4096, pes, eptr, 10);
    std::vector<std::vector<unsigned long, tapa::aligned_allocator<unsigned long>>> img(8);
    edge_list_64bit(pes, eptr, img, 8);
    *ptr = dup(eptr);
    *ptr_len = (int)eptr.size();
    *image_len = (long)img[0].size();
    for (int c = 0; c < 8; ++c) images[c] = dup(img[c]);
    return eptr.empty() ? 0 : eptr.back();
}

void sxref_free(void *p) { std::free(p); }

}  // extern "C"
