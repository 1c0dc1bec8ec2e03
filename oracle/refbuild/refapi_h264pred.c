/*
 * oracle/refbuild/refapi_h264pred.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 * ref_h264_pred / ref_h264_pred_add: the reference's intra predictors through its own table, filled by
 * ff_h264_pred_init(&hpc, AV_CODEC_ID_H264, 8, 1) (libavcodec/h264pred.c:411-585).  No arithmetic of its own.
 */
#define ORC_PREFIX ref_
#include "../oracle_api.h"

#include <pthread.h>
#include <string.h>

#include "libavutil/cpu.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/h264pred.h"

static pthread_once_t once = PTHREAD_ONCE_INIT;
static H264PredContext hpc;
static void init(void) { av_set_cpu_flags_mask(0); ff_h264_pred_init(&hpc, AV_CODEC_ID_H264, 8, 1); }

void ref_h264_pred(int tab, int mode, uint8_t *src, const uint8_t *topright, int has_topleft, int has_topright, ptrdiff_t stride)
{
    pthread_once(&once, init);
    switch (tab) {
    case 0: hpc.pred4x4[mode](src, topright, stride); break;
    case 1: hpc.pred8x8l[mode](src, has_topleft, has_topright, stride); break;
    case 2: hpc.pred8x8[mode](src, stride); break;
    default: hpc.pred16x16[mode](src, stride); break;
    }
}

void ref_h264_pred_add(int tab, int mode, uint8_t *pix, const int *block_offset, int16_t *block, int has_topleft, int has_topright,
                       ptrdiff_t stride)
{
    pthread_once(&once, init);
    switch (tab) {
    case 0: hpc.pred4x4_add[mode](pix, block, stride); break;
    case 1: hpc.pred8x8l_add[mode](pix, block, stride); break;
    case 2: hpc.pred8x8l_filter_add[mode](pix, block, has_topleft, has_topright, stride); break;
    case 3: hpc.pred8x8_add[mode ? HOR_PRED8x8 : VERT_PRED8x8](pix, block_offset, block, stride); break;
    default: hpc.pred16x16_add[mode ? HOR_PRED8x8 : VERT_PRED8x8](pix, block_offset, block, stride); break;
    }
}

/* which entries ff_h264_pred_init(h, AV_CODEC_ID_H264, bit_depth, chroma_format_idc) fills (see ref_table_fill in refapi.c) */
int ref_pred_table_fill(int bit_depth, int chroma_format_idc, uint8_t *out, int cap)
{
    union { H264PredContext h; void *w[256]; } u;
    memset(&u, 0, sizeof(u));
    av_set_cpu_flags_mask(0);
    ff_h264_pred_init(&u.h, AV_CODEC_ID_H264, bit_depth, chroma_format_idc);
    const int n = (int)(sizeof(u.h) / sizeof(void *));
    for (int i = 0; i < n && i < cap; i++) out[i] = u.w[i] != NULL;
    return n;
}
