// libav_b200/csrc/idct_dq.h -- tables of the inverse quantisers fused in front of the simple IDCT (idctdsp.cu), built on
// the host by capi_idct.cu from FFMpegDequantTables and passed to the kernels by value (constant bank).
#pragma once
#include <stdint.h>

namespace avb {

struct DqTables {
    uint16_t intra[64], inter[64];
    uint8_t rank[64];            // scan position of raster index j (inverse of ScanTable.permutated)
    uint8_t raster_end[64];      // ScanTable.raster_end
    int alternate_scan, h263_aic;
};

}  // namespace avb
