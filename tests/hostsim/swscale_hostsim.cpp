// tests/hostsim/swscale_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// libav_b200/csrc/swscale.cu compiled for the host: every kernel of the scaler except the two shared-memory tile kernels is
// thread-independent (one thread = one sample / pixel pair / pixel group), so the whole frame path -- context decisions, staging, the fused
// same-size kernels, the unscaled converters, the source readers, the two-pass general scaler with its range conversion, the packed / nv12 /
// 32-bit output stages -- can run one thread after the other on the CPU.  The source is the product file with its <<<>>> launches rewritten
// mechanically into AVB_LAUNCH(...) calls (gen_launches.py -> _gen/swscale_gen.cu) and the PTX one-liners of common.cuh / swscale.cu replaced
// by the plain definitions those headers carry under AVB_HOSTSIM.  The tile kernels synchronise their threads and abort if launched here:
// the tests force the general scaler onto its two-pass path (avb200_set_tuning("sws_general_variant", 1), a knob the product has for
// profiling); 16-bit destinations, which only exist in the tile kernel, stay GPU-only.
#include "shim/cuda_runtime.h"
namespace avb { int32_t gt_smem[4]; }          // the tile kernels' dynamic shared memory symbol (never used here)
#include "_gen/swscale_gen.cu"
