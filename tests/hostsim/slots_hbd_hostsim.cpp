// tests/hostsim/slots_hbd_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// The product's 9 / 10-bit H.264 slots (libav_b200/csrc/slots_hbd.cu + h264dsp_hbd.cuh) compiled UNCHANGED as host C++ (see slots_hostsim.cpp).
#include "shim/cuda_runtime.h"
#include "../../libav_b200/csrc/slots_hbd.cu"
