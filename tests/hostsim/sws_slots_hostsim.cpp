// tests/hostsim/sws_slots_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// (the contexts the slots belong to come from the host-compiled swscale.cu: swscale_hostsim.cpp)
//
// The product's per-line SwsContext slots (libav_b200/csrc/sws_slots.cu + sws_dev.cuh) compiled UNCHANGED as host C++ against
// shim/cuda_runtime.h (see slots_hostsim.cpp).  The one thing the slots need from the rest of the product -- the SwsSlotView of a
// context made by sws_getContext_cuda() -- cannot come from swscale.cu here (its frame kernels are GPU-only); the test obtains it
// from the product library's host-only sws_debug_slot_view_cuda() and hands it over through hostsim_sws_context().
#include "shim/cuda_runtime.h"
#include "../../libav_b200/csrc/sws_slots.cu"
#include "../../libav_b200/csrc/sws_filter.cu"     // host-only code (filter design, colour constants)

extern "C" {
// the product's own host code for sws_setColorspaceDetails' constants (libav_b200/csrc/sws_filter.cu rgb_constants, linked from the
// product's object file is not possible here -- it lives in a .cu with the filter design -- so the translation unit is included)
int hostsim_rgb_constants(int32_t out[19], const int inv_table[4], int full_range, int brightness, int contrast, int saturation)
{
    avb::RgbConstants k;
    avb::rgb_constants(k, inv_table, full_range, brightness, contrast, saturation);
    memcpy(out, &k, sizeof(k));
    return (int)(sizeof(k) / 4);
}
}
