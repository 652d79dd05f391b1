// stands in for <cuda_runtime.h> when spmm_kernels.cuh is compiled for the CPU emulation
#pragma once
#include "../cuda_emu.h"
