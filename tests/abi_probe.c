/* prints the same facts as ref_abi_info() (oracle/refbuild/refapi.c) for include/avdsp_b200_tables.h */
#include <stdio.h>
#include <stddef.h>
#include "avdsp_b200.h"
int main(void)
{
    long v[] = {
        sizeof(IDCTDSPContext), offsetof(IDCTDSPContext, idct), offsetof(IDCTDSPContext, idct_permutation), offsetof(IDCTDSPContext, perm_type),
        sizeof(FDCTDSPContext), sizeof(BlockDSPContext), offsetof(BlockDSPContext, fill_block_tab),
        sizeof(MECmpContext), offsetof(MECmpContext, sad), offsetof(MECmpContext, nsse), offsetof(MECmpContext, pix_abs),
        sizeof(H264DSPContext), offsetof(H264DSPContext, h264_v_loop_filter_luma), offsetof(H264DSPContext, h264_idct_add),
        offsetof(H264DSPContext, h264_idct_add16), offsetof(H264DSPContext, h264_add_pixels8_clear), offsetof(H264DSPContext, startcode_find_candidate),
        sizeof(H264QpelContext), offsetof(H264QpelContext, avg_h264_qpel_pixels_tab),
        sizeof(H264ChromaContext), sizeof(HpelDSPContext), offsetof(HpelDSPContext, put_no_rnd_pixels_tab), offsetof(HpelDSPContext, avg_no_rnd_pixels_tab),
        sizeof(H264PredContext), offsetof(H264PredContext, pred8x8l), offsetof(H264PredContext, pred16x16), offsetof(H264PredContext, pred8x8l_filter_add),
        offsetof(H264PredContext, pred16x16_add),
        sizeof(PixblockDSPContext), offsetof(PixblockDSPContext, diff_pixels),
        sizeof(QpelDSPContext), offsetof(QpelDSPContext, put_no_rnd_qpel_pixels_tab),
    };
    for (unsigned i = 0; i < sizeof(v) / sizeof(v[0]); i++) printf("%ld\n", v[i]);
    return 0;
}
