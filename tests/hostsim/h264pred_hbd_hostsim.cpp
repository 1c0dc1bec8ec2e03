// tests/hostsim/h264pred_hbd_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// The product's 9 / 10-bit H264PredContext slots (libav_b200/csrc/h264pred_hbd.cu) compiled UNCHANGED as host C++.
#include "shim/cuda_runtime.h"
#include "../../libav_b200/csrc/h264pred_hbd.cu"

// ff_h264_pred_init_cuda() itself lives in h264pred.cu (its 8-bit kernels use shared memory); this is the branch it takes for 9 / 10 bit
extern "C" void hostsim_h264_pred_init_hbd(H264PredContext *h, int bits) { avb::h264pred_init_hbd(h, bits); }
extern "C" void hostsim_h264_pred_install_422(H264PredContext *h, int bits) { avb::h264pred_install_422(h, bits); }
