// tests/hostsim/h264_hbd_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// The batched 9 / 10-bit / 4:2:2 residual and motion-compensation kernels (libav_b200/csrc/h264_hbd_batch.cu; their lanes never communicate)
// compiled UNCHANGED as host C++.  The deblocking wavefront of that file synchronises its warps and stays GPU-only.
#include "shim/cuda_runtime.h"
#include "../../libav_b200/csrc/h264_hbd_batch.cu"
