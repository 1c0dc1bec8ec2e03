// tests/hostsim/idct10_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// The product's 10-bit simple IDCT (libav_b200/csrc/idct10.cu: kernel, batched entry point, table slots) compiled UNCHANGED as host C++.
#include "shim/cuda_runtime.h"
#include "../../libav_b200/csrc/idct10.cu"

// ff_idctdsp_init_cuda() itself lives in capi_idct.cu (GPU-only kernels); this is the branch it takes for bits_per_raw_sample == 10
extern "C" void hostsim_idctdsp_init10(IDCTDSPContext *c) { avb::idct10_install(c); }
