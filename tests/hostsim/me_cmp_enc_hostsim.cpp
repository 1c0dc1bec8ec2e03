// tests/hostsim/me_cmp_enc_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// The product's quant_psnr / bit / rd metrics (libav_b200/csrc/me_cmp_enc.cu: kernel, batched entry point, state objects, table slots)
// compiled UNCHANGED as host C++: one thread per record, no shared memory, no collectives.
#include "shim/cuda_runtime.h"
#include "../../libav_b200/csrc/me_cmp_enc.cu"
