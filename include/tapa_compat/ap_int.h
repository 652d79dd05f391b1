/*
 * ap_int.h -- source-compatibility stand-in for the Xilinx HLS header, host side only.
 *
 * The UNMODIFIED Sextans host program (src/sextans-host.cpp:8,240) only NAMES the type
 * ap_uint<512>, as the element type its 64-bit edge words are reinterpreted to before
 * they go to the kernel; it never touches a bit of it.  An opaque W-bit blob is all the
 * host needs.  (The FPGA kernel source src/sextans.cpp, which does use the bit-slicing
 * operators, is what libsextans_b200.so replaces -- it is not compiled.)
 */
#ifndef SEXTANS_B200_COMPAT_AP_INT_H
#define SEXTANS_B200_COMPAT_AP_INT_H

template <int W>
struct ap_uint {
    static_assert(W > 0, "width");
    unsigned char bytes[(W + 7) / 8];
};

#endif
