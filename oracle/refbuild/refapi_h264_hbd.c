/*
 * oracle/refbuild/refapi_h264_hbd.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 * ref_h264_hbd_*: the 9 / 10-bit entries of the tables the UNMODIFIED reference fills in ff_h264dsp_init(c, bits, chroma_format_idc),
 * ff_h264qpel_init(c, bits), ff_h264chroma_init(c, bits) (oracle/oracle_api.h).
 */
#include <pthread.h>
#include <stdint.h>
#include <stddef.h>
#include "libavutil/cpu.h"
#include "libavcodec/h264dsp.h"
#include "libavcodec/h264qpel.h"
#include "libavcodec/h264chroma.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/h264pred.h"
#define ORC_PREFIX ref_
#include "../oracle_api.h"

static pthread_once_t once = PTHREAD_ONCE_INIT;
static H264DSPContext dsp[2][2];          /* [bits 9 / 10][chroma_format_idc 1 / 2] */
static H264QpelContext qpel[2];
static H264ChromaContext chroma[2];
static H264PredContext pred[2], pred422[3];      /* pred422: chroma_format_idc 2 at 8 / 9 / 10 bit */
static void init_all(void)
{
    av_set_cpu_flags_mask(0);
    ff_h264_pred_init(&pred422[0], AV_CODEC_ID_H264, 8, 2);
    for (int b = 0; b < 2; b++) {
        ff_h264dsp_init(&dsp[b][0], 9 + b, 1); ff_h264dsp_init(&dsp[b][1], 9 + b, 2);
        ff_h264qpel_init(&qpel[b], 9 + b); ff_h264chroma_init(&chroma[b], 9 + b);
        ff_h264_pred_init(&pred[b], AV_CODEC_ID_H264, 9 + b, 1);
        ff_h264_pred_init(&pred422[1 + b], AV_CODEC_ID_H264, 9 + b, 2);
    }
}
#define D(bits, idc2) (pthread_once(&once, init_all), &dsp[(bits) - 9][idc2])

void ref_h264_hbd_idct(int bits, int which, uint8_t *dst, int32_t *block, int stride)
{
    H264DSPContext *c = D(bits, 0); int16_t *b = (int16_t *)block;
    switch (which) {
    case 0: c->h264_idct_add(dst, b, stride); break;
    case 1: c->h264_idct8_add(dst, b, stride); break;
    case 2: c->h264_idct_dc_add(dst, b, stride); break;
    default: c->h264_idct8_dc_add(dst, b, stride); break;
    }
}
void ref_h264_hbd_idct_mb(int bits, int which, uint8_t *dst, uint8_t **dst2, const int *bo, int32_t *block, int stride, const uint8_t *nnzc)
{
    H264DSPContext *c = D(bits, which == 4); int16_t *b = (int16_t *)block;
    switch (which) {
    case 0: c->h264_idct_add16(dst, bo, b, stride, nnzc); break;
    case 1: c->h264_idct_add16intra(dst, bo, b, stride, nnzc); break;
    case 2: c->h264_idct8_add4(dst, bo, b, stride, nnzc); break;
    default: c->h264_idct_add8(dst2, bo, b, stride, nnzc); break;
    }
}
void ref_h264_hbd_dc_dequant(int bits, int kind, int32_t *out, int32_t *in, int qmul)
{
    if (kind == 0) D(bits, 0)->h264_luma_dc_dequant_idct((int16_t *)out, (int16_t *)in, qmul);
    else D(bits, kind == 2)->h264_chroma_dc_dequant_idct((int16_t *)out, qmul);
}
void ref_h264_hbd_add_pixels_clear(int bits, int w8, uint8_t *dst, int32_t *block, int stride)
{
    H264DSPContext *c = D(bits, 0);
    if (w8) c->h264_add_pixels8_clear(dst, (int16_t *)block, stride); else c->h264_add_pixels4_clear(dst, (int16_t *)block, stride);
}
void ref_h264_hbd_weight(int bits, int widx, uint8_t *b, int stride, int h, int ld, int w, int off) { D(bits, 0)->weight_h264_pixels_tab[widx](b, stride, h, ld, w, off); }
void ref_h264_hbd_biweight(int bits, int widx, uint8_t *d, uint8_t *s, int stride, int h, int ld, int wd, int ws, int off)
{ D(bits, 0)->biweight_h264_pixels_tab[widx](d, s, stride, h, ld, wd, ws, off); }
void ref_h264_hbd_loop_filter(int bits, int which, uint8_t *pix, int stride, int alpha, int beta, const int8_t *tc0)
{
    H264DSPContext *c = D(bits, which >= 12); int8_t *t = (int8_t *)tc0;
    switch (which) {
    case 0: c->h264_v_loop_filter_luma(pix, stride, alpha, beta, t); break;
    case 1: c->h264_h_loop_filter_luma(pix, stride, alpha, beta, t); break;
    case 2: c->h264_v_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 3: c->h264_h_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 4: c->h264_v_loop_filter_chroma(pix, stride, alpha, beta, t); break;
    case 5: case 12: c->h264_h_loop_filter_chroma(pix, stride, alpha, beta, t); break;
    case 6: c->h264_v_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 7: case 13: c->h264_h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 8: c->h264_h_loop_filter_luma_mbaff(pix, stride, alpha, beta, t); break;
    case 9: c->h264_h_loop_filter_luma_mbaff_intra(pix, stride, alpha, beta); break;
    case 10: case 14: c->h264_h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, t); break;
    default: c->h264_h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    }
}
void ref_h264_hbd_qpel(int bits, int avg, int sidx, int mc, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    pthread_once(&once, init_all);
    (avg ? qpel[bits - 9].avg_h264_qpel_pixels_tab : qpel[bits - 9].put_h264_qpel_pixels_tab)[sidx][mc](dst, src, stride);
}
void ref_h264_hbd_chroma(int bits, int avg, int widx, uint8_t *dst, uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    pthread_once(&once, init_all);
    (avg ? chroma[bits - 9].avg_h264_chroma_pixels_tab : chroma[bits - 9].put_h264_chroma_pixels_tab)[widx](dst, src, stride, h, x, y);
}

void ref_h264_hbd_pred(int bits, int tab, int mode, uint8_t *src, const uint8_t *topright, int has_topleft, int has_topright, ptrdiff_t stride)
{
    pthread_once(&once, init_all);
    H264PredContext *h = &pred[bits - 9];
    switch (tab) {
    case 0: h->pred4x4[mode](src, topright, stride); break;
    case 1: h->pred8x8l[mode](src, has_topleft, has_topright, stride); break;
    case 2: h->pred8x8[mode](src, stride); break;
    default: h->pred16x16[mode](src, stride); break;
    }
}
void ref_h264_hbd_pred_add(int bits, int tab, int mode, uint8_t *pix, const int *block_offset, int32_t *block, int has_topleft, int has_topright, ptrdiff_t stride)
{
    pthread_once(&once, init_all);
    H264PredContext *h = &pred[bits - 9];
    int16_t *b = (int16_t *)block;
    switch (tab) {
    case 0: h->pred4x4_add[mode](pix, b, stride); break;
    case 1: h->pred8x8l_add[mode](pix, b, stride); break;
    case 2: h->pred8x8l_filter_add[mode](pix, b, has_topleft, has_topright, stride); break;
    case 3: h->pred8x8_add[mode ? HOR_PRED8x8 : VERT_PRED8x8](pix, block_offset, b, stride); break;
    default: h->pred16x16_add[mode ? HOR_PRED8x8 : VERT_PRED8x8](pix, block_offset, b, stride); break;
    }
}

void ref_h264_pred422(int bits, int mode, uint8_t *src, ptrdiff_t stride)
{
    pthread_once(&once, init_all);
    pred422[bits - 8].pred8x8[mode](src, stride);
}
void ref_h264_pred422_add(int bits, int add_mode, uint8_t *pix, const int *block_offset, void *block, ptrdiff_t stride)
{
    pthread_once(&once, init_all);
    pred422[bits - 8].pred8x8_add[add_mode ? HOR_PRED8x8 : VERT_PRED8x8](pix, block_offset, (int16_t *)block, stride);
}
