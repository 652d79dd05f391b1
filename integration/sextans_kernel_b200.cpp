// sextans_kernel_b200.cpp -- takes the place of src/sextans.cpp in the reference tree.
//
// Defines the reference's own top-level task
//     void Sextans(tapa::mmap<int> edge_list_ptr, tapa::mmaps<ap_uint<512>, 8> edge_list_ch,
//                  tapa::mmaps<float_v16, 4> mat_B_ch, tapa::mmaps<float_v16, 8> mat_C_ch_in,
//                  tapa::mmaps<float_v16, 8> mat_C_ch, int NUM_ITE, int NUM_A_LEN, int M,
//                  int K, int P_N, int alpha_u, int beta_u)            (src/sextans.h:20-26)
// on top of libsextans_b200.so, so that the UNMODIFIED host program src/sextans-host.cpp
// -- its loader, its FPGA preprocessing, its B/C channel repacking, its verification --
// runs against a B200 instead of the FPGA / the TAPA software simulation:
//
//   g++ -O2 -std=c++17 -I<repo>/include/tapa_compat -I<repo>/include -I<ref>/src
//       <ref>/src/sextans-host.cpp <repo>/integration/sextans_kernel_b200.cpp
//       -L<repo>/sextans_b200 -lsextans_b200 -Wl,-rpath,<repo>/sextans_b200 -o sextans
//
// (oracle/Makefile target `ref_host` does exactly this into oracle/_ref/.)  It includes
// the reference's sextans.h for the prototype, so it only compiles next to the reference.
// Failures end the process, as a failing XRT/TAPA launch does in the reference
// (SURVEY.md section 8(b) "Return / errors").  Device: env SEXTANS_DEVICE (default 0).
#include <cstdio>
#include <cstdlib>

#include "sextans.h"       // the reference's prototype and channel counts
#include "sextans_b200.h"  // the engine's C ABI

static_assert(NUM_CH_SPARSE == SX_IMAGES_A_CHANNELS && NUM_CH_B == SX_IMAGES_B_CHANNELS &&
                  NUM_CH_C == SX_IMAGES_C_CHANNELS && WINDOW_SIZE == SX_IMAGES_WINDOW,
              "the image decoder is written for the shipped configuration (src/sextans.h:7-11)");

namespace {
[[noreturn]] void die(const char *what, int rc) {
    std::fprintf(stderr, "Sextans (B200): %s: %s: %s\n", what, sx_status_name(rc), sx_last_error());
    std::exit(1);
}

struct Device {
    sx_ctx *ctx = nullptr;
    Device() {
        const char *e = std::getenv("SEXTANS_DEVICE");
        const int rc = sx_create(e ? std::atoi(e) : 0, &ctx);
        if (rc) die("sx_create", rc);
    }
    ~Device() { sx_destroy(ctx); }
};
}  // namespace

void Sextans(tapa::mmap<int> edge_list_ptr, tapa::mmaps<ap_uint<512>, NUM_CH_SPARSE> edge_list_ch,
             tapa::mmaps<float_v16, NUM_CH_B> mat_B_ch, tapa::mmaps<float_v16, NUM_CH_C> mat_C_ch_in,
             tapa::mmaps<float_v16, NUM_CH_C> mat_C_ch, const int NUM_ITE, const int NUM_A_LEN, const int M,
             const int K, const int P_N, const int alpha_u, int beta_u) {
    static Device dev;  // one context per process, like the one bitstream the reference loads
    const int N = P_N & 0xFFFF;
    // the buffers are caller-owned; refuse any that is smaller than what the scalars imply
    // instead of reading past its end (the FPGA would)
    if ((long long)edge_list_ptr.size() < (long long)NUM_ITE + 1) die("edge_list_ptr too short", SX_ERR_INVALID);
    const uint64_t *a[NUM_CH_SPARSE];
    const float *b[NUM_CH_B], *cin[NUM_CH_C];
    float *cout[NUM_CH_C];
    for (int c = 0; c < NUM_CH_SPARSE; ++c) {
        if ((long long)edge_list_ch[c].size() * 8 < sx_images_A_words(NUM_A_LEN)) die("A channel image too short", SX_ERR_INVALID);
        a[c] = reinterpret_cast<const uint64_t *>(edge_list_ch[c].data());
    }
    for (int c = 0; c < NUM_CH_B; ++c) {
        if ((long long)mat_B_ch[c].size() * 16 < sx_images_B_floats(K, N)) die("B channel image too short", SX_ERR_INVALID);
        b[c] = reinterpret_cast<const float *>(mat_B_ch[c].data());
    }
    for (int c = 0; c < NUM_CH_C; ++c) {
        if ((long long)mat_C_ch_in[c].size() * 16 < sx_images_C_floats(M, N) ||
            (long long)mat_C_ch[c].size() * 16 < sx_images_C_floats(M, N))
            die("C channel image too short", SX_ERR_INVALID);
        cin[c] = reinterpret_cast<const float *>(mat_C_ch_in[c].data());
        cout[c] = reinterpret_cast<float *>(mat_C_ch[c].data());
    }
    const int rc = sx_sextans_invoke(dev.ctx, edge_list_ptr.data(), a, b, cin, cout, NUM_ITE, NUM_A_LEN, M, K, P_N,
                                     alpha_u, beta_u, nullptr);
    if (rc) die("sx_sextans_invoke", rc);
}
