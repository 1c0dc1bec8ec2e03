/*
 * oracle/refbuild/refapi.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Implements oracle/oracle_api.h with the prefix ref_ by calling the UNMODIFIED reference
 * (compiled from /root/reference by build_ref.sh into oracle/_ref/libavref.so) through
 * the reference's own init functions and function-pointer tables: every call goes
 * ff_*_init() -> table slot, exactly what a codec does.  No arithmetic of its own.
 */
#define ORC_PREFIX ref_
#include "../oracle_api.h"

#include <pthread.h>
#include <string.h>
#include <stdlib.h>

#include "libavutil/mem.h"
#include "libavutil/cpu.h"
#include "libavutil/pixfmt.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/idctdsp.h"
#include "libavcodec/fdctdsp.h"
#include "libavcodec/blockdsp.h"
#include "libavcodec/me_cmp.h"
#include "libavcodec/h264dsp.h"
#include "libavcodec/h264qpel.h"
#include "libavcodec/h264chroma.h"
#include "libavcodec/hpeldsp.h"
#include "libavcodec/h264pred.h"
#include "libavcodec/pixblockdsp.h"
#include "libavcodec/qpeldsp.h"
#include "libavcodec/fft.h"
#include "libavcodec/dct.h"
#include "libswscale/swscale.h"
#include "libswscale/swscale_internal.h"

static pthread_once_t once = PTHREAD_ONCE_INIT;
static IDCTDSPContext idsp, idsp10;      /* bits_per_raw_sample 8 and 10 */
static FDCTDSPContext fdsp_islow, fdsp_ifast, fdsp10;
static BlockDSPContext bdsp;
static MECmpContext mecc;
static H264DSPContext h264, h264_422;     /* chroma_format_idc 1 and 2 */
static H264QpelContext qpel;
static H264ChromaContext chroma;
static HpelDSPContext hpel;
static PixblockDSPContext pixb;
static QpelDSPContext mqpel;

static void init_all(void)
{
    AVCodecContext *avctx = calloc(1, sizeof(*avctx));
#ifdef REF_SIMD                           /* oracle/_ref/libavref_simd.so, TIMING ONLY: the tables a default-configured decoder gets on this host */
    avctx->idct_algo           = FF_IDCT_AUTO;
#else
    av_set_cpu_flags_mask(0);             /* what FATE's -cpuflags 0 does: portable C only */
    avctx->idct_algo           = FF_IDCT_SIMPLE;
#endif
    avctx->bits_per_raw_sample = 8;
#ifndef REF_SIMD
    avctx->flags               = AV_CODEC_FLAG_BITEXACT;
#endif
    avctx->dct_algo            = FF_DCT_INT;
    ff_idctdsp_init(&idsp, avctx);
    avctx->bits_per_raw_sample = 10;
    ff_idctdsp_init(&idsp10, avctx);
    ff_fdctdsp_init(&fdsp10, avctx);
    avctx->bits_per_raw_sample = 8;
    ff_fdctdsp_init(&fdsp_islow, avctx);
    avctx->dct_algo = FF_DCT_FASTINT;
    ff_fdctdsp_init(&fdsp_ifast, avctx);
    ff_blockdsp_init(&bdsp);
    ff_me_cmp_init_static();
    ff_me_cmp_init(&mecc, avctx);
    ff_h264dsp_init(&h264, 8, 1);
    ff_h264dsp_init(&h264_422, 8, 2);
    ff_h264qpel_init(&qpel, 8);
    ff_h264chroma_init(&chroma, 8);
    ff_pixblockdsp_init(&pixb, avctx);
    ff_qpeldsp_init(&mqpel);
    ff_hpeldsp_init(&hpel, AV_CODEC_FLAG_BITEXACT);
    free(avctx);
}
#define INIT() pthread_once(&once, init_all)

void ref_simple_idct_put(uint8_t *d, ptrdiff_t s, int16_t *b) { INIT(); idsp.idct_put(d, s, b); }
void ref_simple_idct_add(uint8_t *d, ptrdiff_t s, int16_t *b) { INIT(); idsp.idct_add(d, s, b); }
void ref_simple_idct(int16_t *b) { INIT(); idsp.idct(b); }
void ref_simple_idct10(int mode, uint8_t *d, ptrdiff_t s, int16_t *b)
{ INIT(); if (mode == 0) idsp10.idct_put(d, s, b); else if (mode == 1) idsp10.idct_add(d, s, b); else idsp10.idct(b); }
void ref_put_pixels_clamped(const int16_t *b, uint8_t *p, ptrdiff_t s) { INIT(); idsp.put_pixels_clamped(b, p, s); }
void ref_put_signed_pixels_clamped(const int16_t *b, uint8_t *p, ptrdiff_t s) { INIT(); idsp.put_signed_pixels_clamped(b, p, s); }
void ref_add_pixels_clamped(const int16_t *b, uint8_t *p, ptrdiff_t s) { INIT(); idsp.add_pixels_clamped(b, p, s); }
void ref_clear_block(int16_t *b) { INIT(); bdsp.clear_block(b); }
void ref_clear_blocks(int16_t *b) { INIT(); bdsp.clear_blocks(b); }
void ref_fill_block(int w16, uint8_t *b, uint8_t v, ptrdiff_t s, int h) { INIT(); bdsp.fill_block_tab[w16 ? 0 : 1](b, v, s, h); }

struct idct_job { int mode; int16_t *blocks; uint8_t *frame; const uint32_t *off; ptrdiff_t stride; size_t lo, hi; };
static void *idct_worker(void *p)
{
    struct idct_job *j = p;
    for (size_t i = j->lo; i < j->hi; i++) {
        int16_t *b = j->blocks + 64 * i;
        if (j->mode == 0)      idsp.idct_put(j->frame + j->off[i], j->stride, b);
        else if (j->mode == 1) idsp.idct_add(j->frame + j->off[i], j->stride, b);
        else                   idsp.idct(b);
    }
    return NULL;
}
void ref_idct_batch(int mode, int16_t *blocks, uint8_t *frame, const uint32_t *off, ptrdiff_t stride,
                    size_t n, int nthreads)
{
    INIT();
    if (nthreads < 1) nthreads = 1;
    pthread_t th[256]; struct idct_job jobs[256];
    if (nthreads > 256) nthreads = 256;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (struct idct_job){ mode, blocks, frame, off, stride, n * t / nthreads, n * (t + 1) / nthreads };
        if (nthreads == 1) idct_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, idct_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

void ref_fdct(int which, int16_t *block)
{
    INIT();
    switch (which) {
    case 0: fdsp_islow.fdct(block); break;
    case 1: fdsp_islow.fdct248(block); break;
    case 2: fdsp_ifast.fdct(block); break;
    case 3: fdsp_ifast.fdct248(block); break;
    case 4: fdsp10.fdct(block); break;            /* ff_jpeg_fdct_islow_10, fdctdsp.c:31-33 */
    default: fdsp10.fdct248(block); break;
    }
}

/* ff_set_cmp() (libavcodec/me_cmp.c:365-417), the reference's own selection of the six compare functions of a kind out of ANY MECmpContext --
 * e.g. the one ff_me_cmp_init_cuda() filled (tests/test_me_cmp_select_cpu.py).  Only in _ref. */
void ref_me_cmp_select(void *table, int type, void *out[6]) { INIT(); ff_set_cmp((MECmpContext *)table, (me_cmp_func *)out, type); }

/* Which entries does the reference's own init fill?  kind: 0 H264DSPContext (a = bit_depth, b = chroma_format_idc), 1 H264QpelContext (a = bit_depth),
 * 2 H264ChromaContext (a), 3 H264PredContext for AV_CODEC_ID_H264 (a, b), 4 HpelDSPContext (a = flags), 5 MECmpContext, 6 QpelDSPContext,
 * 7 BlockDSPContext, 8 FDCTDSPContext (a = bits_per_raw_sample, b = dct_algo), 9 PixblockDSPContext (a = bits), 10 IDCTDSPContext (a = bits, b = idct_algo;
 * only its six function pointers).  out[i] = 1 when pointer-sized word i of the table is non-zero.  Returns the word count.  Only in _ref. */
int ref_table_fill(int kind, int a, int b, uint8_t *out, int cap)
{
    INIT();
    union { H264DSPContext h; H264QpelContext q; H264ChromaContext c; HpelDSPContext hp; MECmpContext m; QpelDSPContext mq; BlockDSPContext bd;
            FDCTDSPContext f; PixblockDSPContext p; IDCTDSPContext i; void *w[4096]; } u;
    AVCodecContext *avctx = calloc(1, sizeof(*avctx));
    size_t bytes = 0;
    memset(&u, 0, sizeof(u));
    avctx->bits_per_raw_sample = a; avctx->flags = AV_CODEC_FLAG_BITEXACT; avctx->idct_algo = b; avctx->dct_algo = b;
    switch (kind) {
    case 0: ff_h264dsp_init(&u.h, a, b); bytes = sizeof(u.h); break;
    case 1: ff_h264qpel_init(&u.q, a); bytes = sizeof(u.q); break;
    case 2: ff_h264chroma_init(&u.c, a); bytes = sizeof(u.c); break;
    case 4: ff_hpeldsp_init(&u.hp, a); bytes = sizeof(u.hp); break;
    case 5: ff_me_cmp_init(&u.m, avctx); bytes = sizeof(u.m); break;
    case 6: ff_qpeldsp_init(&u.mq); bytes = sizeof(u.mq); break;
    case 7: ff_blockdsp_init(&u.bd); bytes = sizeof(u.bd); break;
    case 8: ff_fdctdsp_init(&u.f, avctx); bytes = sizeof(u.f); break;
    case 9: ff_pixblockdsp_init(&u.p, avctx); bytes = sizeof(u.p); break;
    case 10: ff_idctdsp_init(&u.i, avctx); bytes = 6 * sizeof(void *); break;
    default: free(avctx); return -1;
    }
    free(avctx);
    const int n = (int)(bytes / sizeof(void *));
    for (int i = 0; i < n && i < cap; i++) out[i] = u.w[i] != NULL;
    return n;
}

void ref_h264_idct(int which, uint8_t *dst, int16_t *block, int stride)
{
    INIT();
    switch (which) {
    case 0: h264.h264_idct_add(dst, block, stride); break;
    case 1: h264.h264_idct8_add(dst, block, stride); break;
    case 2: h264.h264_idct_dc_add(dst, block, stride); break;
    default: h264.h264_idct8_dc_add(dst, block, stride); break;
    }
}
void ref_h264_idct_mb(int which, uint8_t *dst, uint8_t **dst2, const int *bo, int16_t *block, int stride,
                      const uint8_t *nnzc)
{
    INIT();
    switch (which) {
    case 0: h264.h264_idct_add16(dst, bo, block, stride, nnzc); break;
    case 1: h264.h264_idct_add16intra(dst, bo, block, stride, nnzc); break;
    case 2: h264.h264_idct8_add4(dst, bo, block, stride, nnzc); break;
    case 3: h264.h264_idct_add8(dst2, bo, block, stride, nnzc); break;
    default: h264_422.h264_idct_add8(dst2, bo, block, stride, nnzc); break;     /* ff_h264_idct_add8_422 */
    }
}
void ref_h264_luma_dc_dequant_idct(int16_t *o, int16_t *i, int q) { INIT(); h264.h264_luma_dc_dequant_idct(o, i, q); }
void ref_h264_chroma_dc_dequant_idct(int16_t *b, int q) { INIT(); h264.h264_chroma_dc_dequant_idct(b, q); }
void ref_h264_chroma422_dc_dequant_idct(int16_t *b, int q) { INIT(); h264_422.h264_chroma_dc_dequant_idct(b, q); }
void ref_h264_loop_filter(int which, uint8_t *pix, int stride, int alpha, int beta, const int8_t *tc0)
{
    INIT();
    int8_t *t = (int8_t *)tc0;
    switch (which) {
    case 0: h264.h264_v_loop_filter_luma(pix, stride, alpha, beta, t); break;
    case 1: h264.h264_h_loop_filter_luma(pix, stride, alpha, beta, t); break;
    case 2: h264.h264_v_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 3: h264.h264_h_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 4: h264.h264_v_loop_filter_chroma(pix, stride, alpha, beta, t); break;
    case 5: h264.h264_h_loop_filter_chroma(pix, stride, alpha, beta, t); break;
    case 6: h264.h264_v_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 7: h264.h264_h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 8: h264.h264_h_loop_filter_luma_mbaff(pix, stride, alpha, beta, t); break;
    case 9: h264.h264_h_loop_filter_luma_mbaff_intra(pix, stride, alpha, beta); break;
    case 10: h264.h264_h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, t); break;
    case 11: h264.h264_h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    case 12: h264_422.h264_h_loop_filter_chroma(pix, stride, alpha, beta, t); break;              /* chroma422 */
    case 13: h264_422.h264_h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 14: h264_422.h264_h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, t); break;        /* chroma422_mbaff */
    default: h264_422.h264_h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    }
}
void ref_h264_weight(int widx, uint8_t *b, int stride, int h, int ld, int w, int off)
{ INIT(); h264.weight_h264_pixels_tab[widx](b, stride, h, ld, w, off); }
void ref_h264_biweight(int widx, uint8_t *d, uint8_t *s, int stride, int h, int ld, int wd, int ws, int off)
{ INIT(); h264.biweight_h264_pixels_tab[widx](d, s, stride, h, ld, wd, ws, off); }
void ref_h264_add_pixels_clear(int w8, uint8_t *dst, int16_t *block, int stride)
{ INIT(); if (w8) h264.h264_add_pixels8_clear(dst, block, stride); else h264.h264_add_pixels4_clear(dst, block, stride); }

void ref_h264_qpel(int avg, int sidx, int mc, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    INIT();
    (avg ? qpel.avg_h264_qpel_pixels_tab : qpel.put_h264_qpel_pixels_tab)[sidx][mc](dst, src, stride);
}
void ref_h264_chroma(int avg, int widx, uint8_t *dst, uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    INIT();
    (avg ? chroma.avg_h264_chroma_pixels_tab : chroma.put_h264_chroma_pixels_tab)[widx](dst, src, stride, h, x, y);
}

int ref_hpel(int tab, int sidx, int dxy, uint8_t *block, const uint8_t *pixels, ptrdiff_t ls, int h)
{
    INIT();
    op_pixels_func f = NULL;
    switch (tab) {
    case 0: f = hpel.put_pixels_tab[sidx][dxy]; break;
    case 1: f = hpel.avg_pixels_tab[sidx][dxy]; break;
    case 2: f = hpel.put_no_rnd_pixels_tab[sidx][dxy]; break;
    default: f = sidx == 0 ? hpel.avg_no_rnd_pixels_tab[dxy] : NULL; break;
    }
    if (!f) return -1;
    f(block, pixels, ls, h);
    return 0;
}

int ref_me_cmp(int kind, int sidx, int dxy, const uint8_t *b1, const uint8_t *b2, ptrdiff_t stride, int h)
{
    INIT();
    me_cmp_func f = NULL;
    uint8_t *p1 = (uint8_t *)b1, *p2 = (uint8_t *)b2;
    switch (kind) {
    case 0: f = mecc.pix_abs[sidx][dxy]; break;
    case 1: f = mecc.sad[sidx]; break;
    case 2: f = mecc.sse[sidx]; break;
    case 3: f = mecc.hadamard8_diff[sidx]; break;
    case 4: f = mecc.vsad[sidx]; break;
    case 5: f = mecc.vsse[sidx]; break;
    case 6: f = mecc.nsse[sidx]; break;
    case 7: f = mecc.hadamard8_diff[4 + sidx]; break;
    case 8: f = mecc.vsad[4 + sidx]; break;
    case 9: f = mecc.vsse[4 + sidx]; break;
    case 10: return mecc.sum_abs_dctelem((int16_t *)p1);
#ifndef REF_SIMD      /* (the timing-only flavour does not link the encoder-context shim) */
    case 11: case 12: case 13: { extern int ref_me_cmp_enc(int, int, int, uint8_t *, uint8_t *, ptrdiff_t, int); return ref_me_cmp_enc(kind, sidx, dxy, p1, p2, stride, h); }
#endif
    }
    if (!f) return -1;
    return f(NULL, p1, p2, stride, h);
}

struct fs_job { const uint8_t *cur, *ref; int stride, w, h, range, y0, y1; int32_t *out; };
static void *fs_worker(void *p)
{
    struct fs_job *j = p;
    int mbw = j->w / 16;
    for (int mby = j->y0; mby < j->y1; mby++)
        for (int mbx = 0; mbx < mbw; mbx++) {
            int px = mbx * 16, py = mby * 16;
            int xmin = -px, xmax = j->w - 16 - px, ymin = -py, ymax = j->h - 16 - py;
            if (xmin < -j->range) xmin = -j->range;
            if (ymin < -j->range) ymin = -j->range;
            if (xmax > j->range) xmax = j->range;
            if (ymax > j->range) ymax = j->range;
            uint8_t *c = (uint8_t *)j->cur + py * j->stride + px;
            int best = 1 << 30, bx = 0, by = 0;
            for (int y = ymin; y <= ymax; y++)
                for (int x = xmin; x <= xmax; x++) {
                    int d = mecc.pix_abs[0][0](NULL, c, (uint8_t *)j->ref + (py + y) * j->stride + px + x, j->stride, 16);
                    if (d < best) { best = d; bx = x; by = y; }
                }
            int32_t *o = j->out + 3 * (mby * mbw + mbx);
            o[0] = bx; o[1] = by; o[2] = best;
        }
    return NULL;
}
void ref_full_search(const uint8_t *cur, const uint8_t *ref, int stride, int w, int h, int range,
                     int y0, int y1, int32_t *out, int nthreads)
{
    INIT();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; struct fs_job jobs[256];
    int rows = y1 - y0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (struct fs_job){ cur, ref, stride, w, h, range, y0 + rows * t / nthreads, y0 + rows * (t + 1) / nthreads, out };
        if (nthreads == 1) fs_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, fs_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ---- swscale ---- */
/* SwsFilter state for the contexts created below (oracle_api.h sws_set_filter) */
static SwsVector g_vec[4][2];
static SwsFilter g_srcF, g_dstF;
void ref_sws_set_filter(int which, int side, const double *coeff, int length)
{
    if (which < 0 || which > 3 || side < 0 || side > 1) return;
    g_vec[which][side].coeff = length > 0 ? (double *)coeff : NULL; g_vec[which][side].length = length > 0 ? length : 0;
    SwsFilter *f = side ? &g_dstF : &g_srcF;
    SwsVector *v = g_vec[which][side].length ? &g_vec[which][side] : NULL;
    if (which == 0) f->lumH = v; else if (which == 1) f->lumV = v; else if (which == 2) f->chrH = v; else f->chrV = v;
}
static SwsFilter *filt(int side)
{
    SwsFilter *f = side ? &g_dstF : &g_srcF;
    return (f->lumH || f->lumV || f->chrH || f->chrV) ? f : NULL;
}
static struct SwsContext *mk_sws(int sw, int sh, int dw, int dh, enum AVPixelFormat df, int flags)
{
    INIT();
    return sws_getContext(sw, sh, AV_PIX_FMT_YUV420P, dw, dh, df, flags, filt(0), filt(1), NULL);
}
int ref_sws_yuv420p_to_rgb24(const uint8_t *const src[3], const int ss[3], int sw, int sh, uint8_t *dst,
                             int dstride, int dw, int dh, int flags)
{
    struct SwsContext *c = mk_sws(sw, sh, dw, dh, AV_PIX_FMT_RGB24, flags);
    if (!c) return -1;
    uint8_t *d[4] = { dst, NULL, NULL, NULL };
    int ds[4] = { dstride, 0, 0, 0 };
    const uint8_t *s[4] = { src[0], src[1], src[2], NULL };
    int sst[4] = { ss[0], ss[1], ss[2], 0 };
    int r = sws_scale(c, s, sst, 0, sh, d, ds);
    sws_freeContext(c);
    return r;
}
int ref_sws_yuv420p_to_yuv420p(const uint8_t *const src[3], const int ss[3], int sw, int sh,
                               uint8_t *const dst[3], const int dstr[3], int dw, int dh, int flags)
{
    struct SwsContext *c = mk_sws(sw, sh, dw, dh, AV_PIX_FMT_YUV420P, flags);
    if (!c) return -1;
    uint8_t *d[4] = { dst[0], dst[1], dst[2], NULL };
    int ds[4] = { dstr[0], dstr[1], dstr[2], 0 };
    const uint8_t *s[4] = { src[0], src[1], src[2], NULL };
    int sst[4] = { ss[0], ss[1], ss[2], 0 };
    int r = sws_scale(c, s, sst, 0, sh, d, ds);
    sws_freeContext(c);
    return r;
}
static int g_cs_set, g_cs_inv[4], g_cs_range, g_cs_b, g_cs_c, g_cs_s;
void ref_sws_set_colorspace(const int inv_table[4], int src_range, int brightness, int contrast, int saturation)
{
    g_cs_set = inv_table != NULL;
    if (inv_table) memcpy(g_cs_inv, inv_table, sizeof(g_cs_inv));
    g_cs_range = src_range; g_cs_b = brightness; g_cs_c = contrast; g_cs_s = saturation;
}
int ref_sws_planar(int src_fmt, const uint8_t *const src[3], const int ss[3], int sw, int sh, int dst_fmt,
                   uint8_t *const dst[3], const int dstride[3], int dw, int dh, int flags)
{
    INIT();
    const int packed_src = src_fmt == AV_PIX_FMT_YUYV422 || src_fmt == AV_PIX_FMT_UYVY422 || src_fmt == AV_PIX_FMT_RGB24 || src_fmt == AV_PIX_FMT_BGR24 ||
                           (src_fmt >= AV_PIX_FMT_ARGB && src_fmt <= AV_PIX_FMT_BGRA);
    struct SwsContext *c = sws_getContext(sw, sh, (enum AVPixelFormat)src_fmt, dw, dh, (enum AVPixelFormat)dst_fmt, flags, filt(0), filt(1), NULL);
    if (!c) return -1;
    if (g_cs_set && sws_setColorspaceDetails(c, g_cs_inv, g_cs_range, g_cs_inv, 0, g_cs_b, g_cs_c, g_cs_s) < 0) { sws_freeContext(c); return -2; }
    const int packed_dst = dst_fmt == AV_PIX_FMT_RGB24 || dst_fmt == AV_PIX_FMT_BGR24 || (dst_fmt >= AV_PIX_FMT_ARGB && dst_fmt <= AV_PIX_FMT_BGRA) ||
                           dst_fmt == AV_PIX_FMT_YUYV422 || dst_fmt == AV_PIX_FMT_UYVY422;
    const int nv_dst = dst_fmt == AV_PIX_FMT_NV12 || dst_fmt == AV_PIX_FMT_NV21;
    uint8_t *d[4] = { dst[0], packed_dst ? NULL : dst[1], packed_dst || nv_dst ? NULL : dst[2], NULL };
    int ds[4] = { dstride[0], packed_dst ? 0 : dstride[1], packed_dst || nv_dst ? 0 : dstride[2], 0 };
    /* a yuva420p source needs a fourth plane to pass sws_scale()'s pointer check (swscale_unscaled.c:1196-1210) even where no alpha is read (every
     * destination without alpha, utils.c:1244): the luma plane stands in for it */
    const int alpha_src = src_fmt == AV_PIX_FMT_YUVA420P;
    const uint8_t *s[4] = { src[0], packed_src ? NULL : src[1], packed_src ? NULL : src[2], alpha_src ? src[0] : NULL };
    int sst[4] = { ss[0], packed_src ? 0 : ss[1], packed_src ? 0 : ss[2], alpha_src ? ss[0] : 0 };
    int r = sws_scale(c, s, sst, 0, sh, d, ds);
    sws_freeContext(c);
    return r;
}
int ref_sws_nv12(int nv21, const uint8_t *y, int ystride, const uint8_t *uv, int uvstride, int sw, int sh, int dst_fmt,
                 uint8_t *const dst[3], const int dstride[3], int dw, int dh, int flags)
{
    INIT();
    struct SwsContext *c = sws_getContext(sw, sh, nv21 ? AV_PIX_FMT_NV21 : AV_PIX_FMT_NV12, dw, dh,
                                          dst_fmt == 2 ? AV_PIX_FMT_RGB24 : AV_PIX_FMT_YUV420P, flags, NULL, NULL, NULL);
    if (!c) return -1;
    uint8_t *d[4] = { dst[0], dst_fmt == 2 ? NULL : dst[1], dst_fmt == 2 ? NULL : dst[2], NULL };
    int ds[4] = { dstride[0], dst_fmt == 2 ? 0 : dstride[1], dst_fmt == 2 ? 0 : dstride[2], 0 };
    const uint8_t *s[4] = { y, uv, NULL, NULL };
    int sst[4] = { ystride, uvstride, 0, 0 };
    int r = sws_scale(c, s, sst, 0, sh, d, ds);
    sws_freeContext(c);
    return r;
}
/* ---- per-line slots: the function pointers the reference itself installed in a (scaling) context ---- */
static struct SwsContext *line_ctx(int dst_fmt, int flags)
{
    INIT();
    struct SwsContext *c = sws_getContext(64, 48, AV_PIX_FMT_YUV420P, 96, 80, (enum AVPixelFormat)dst_fmt, flags, NULL, NULL, NULL);
    if (!c) return NULL;
    if (g_cs_set) sws_setColorspaceDetails(c, g_cs_inv, g_cs_range, g_cs_inv, 0, g_cs_b, g_cs_c, g_cs_s);     /* -1 (nothing changed) for yuv destinations, utils.c:821-822 */
    static const uint8_t pb_64[8] = { 64, 64, 64, 64, 64, 64, 64, 64 };
    c->lumDither8 = c->chrDither8 = pb_64;               /* what swscale() sets for every 8-bit source before the first line (swscale.c:445-447) */
    return c;
}
int ref_sws_line_range(int kind, int16_t *dst1, int16_t *dst2, int width)
{
    INIT();
    struct SwsContext *c = sws_getContext(64, 48, kind < 2 ? AV_PIX_FMT_YUVJ420P : AV_PIX_FMT_YUV420P, 96, 80,
                                          kind < 2 ? AV_PIX_FMT_YUV420P : AV_PIX_FMT_YUVJ420P, SWS_BICUBIC, NULL, NULL, NULL);
    if (!c || !c->lumConvertRange || !c->chrConvertRange) { sws_freeContext(c); return -1; }
    if (kind & 1) c->chrConvertRange(dst1, dst2, width); else c->lumConvertRange(dst1, width);
    sws_freeContext(c);
    return 0;
}
int ref_sws_line_hscale(int dst_fmt, int flags, void *dst, int dstW, const uint8_t *src, const int16_t *filter, const int32_t *filterPos, int filterSize)
{
    struct SwsContext *c = line_ctx(dst_fmt, flags);
    if (!c) return -1;
    c->hyScale(c, dst, dstW, src, filter, filterPos, filterSize);
    int same = c->hcScale == c->hyScale;
    sws_freeContext(c);
    return same ? 0 : -1;
}
int ref_sws_line_hfast(int chroma, int16_t *dst1, int16_t *dst2, int dstW, const uint8_t *src1, const uint8_t *src2, int srcW, int xInc)
{
    struct SwsContext *c = line_ctx(AV_PIX_FMT_YUV420P, SWS_FAST_BILINEAR);
    if (!c || !c->hyscale_fast || !c->hcscale_fast) { sws_freeContext(c); return -1; }
    if (chroma) c->hcscale_fast(c, dst1, dst2, dstW, src1, src2, srcW, xInc);
    else        c->hyscale_fast(c, dst1, dstW, src1, srcW, xInc);
    sws_freeContext(c);
    return 0;
}
int ref_sws_line_plane(int dst_fmt, const int16_t *filter, int filterSize, const void *const *src, uint8_t *dest, int dstW, const uint8_t *dither, int offset)
{
    struct SwsContext *c = line_ctx(dst_fmt, SWS_BICUBIC);
    if (!c) return -1;
    if (filterSize) c->yuv2planeX(filter, filterSize, (const int16_t **)src, dest, dstW, dither, offset);
    else            c->yuv2plane1((const int16_t *)src[0], dest, dstW, dither, offset);
    sws_freeContext(c);
    return 0;
}
int ref_sws_line_nv12(int dst_fmt, const int16_t *chrFilter, int chrFilterSize, const int16_t *const *chrU, const int16_t *const *chrV, uint8_t *dest, int chrDstW)
{
    struct SwsContext *c = line_ctx(dst_fmt, SWS_BICUBIC);
    if (!c || !c->yuv2nv12cX) { sws_freeContext(c); return -1; }
    c->yuv2nv12cX(c, chrFilter, chrFilterSize, (const int16_t **)chrU, (const int16_t **)chrV, dest, chrDstW);
    sws_freeContext(c);
    return 0;
}
int ref_sws_line_packed(int dst_fmt, int flags, int kind, const int16_t *lumFilter, const int16_t *const *lumSrc, int lumFilterSize,
                        const int16_t *chrFilter, const int16_t *const *chrU, const int16_t *const *chrV, int chrFilterSize,
                        uint8_t *dest, int dstW, int yalpha, int uvalpha, int y)
{
    struct SwsContext *c = line_ctx(dst_fmt, flags);
    if (!c) return -1;
    int r = 0;
    if (kind == 1) { if (c->yuv2packed1) c->yuv2packed1(c, lumSrc[0], (const int16_t **)chrU, (const int16_t **)chrV, NULL, dest, dstW, uvalpha, y); else r = -1; }
    else if (kind == 2) { if (c->yuv2packed2) c->yuv2packed2(c, (const int16_t **)lumSrc, (const int16_t **)chrU, (const int16_t **)chrV, NULL, dest, dstW, yalpha, uvalpha, y); else r = -1; }
    else { if (c->yuv2packedX) c->yuv2packedX(c, lumFilter, (const int16_t **)lumSrc, lumFilterSize, chrFilter, (const int16_t **)chrU, (const int16_t **)chrV, chrFilterSize, NULL, dest, dstW, y); else r = -1; }
    sws_freeContext(c);
    return r;
}

/* ---- the reference's own scheduler with foreign per-line slots: sws_getContext() -> the caller replaces the slot fields ->
 * sws_scale().  Used to show that the product's SwsContext slots are a drop-in under the UNMODIFIED swscale() line loop
 * (tests/test_sws_dropin_cpu.py: host simulation; tests/test_zz_gpu_late_slots.py: the product library).  Only in _ref. ---- */
void *ref_sws_open(int src_fmt, int sw, int sh, int dst_fmt, int dw, int dh, int flags)
{
    INIT();
    return sws_getContext(sw, sh, (enum AVPixelFormat)src_fmt, dw, dh, (enum AVPixelFormat)dst_fmt, flags, NULL, NULL, NULL);
}
/* slots: the 12 pointers of SwsLineSlotsCUDA (include/avdsp_b200.h) in its order; NULL entries leave the C function in place.
 * Returns how many fields were replaced, -1 if the table asks for a slot the reference itself does not have for this context. */
int ref_sws_set_slots(void *ctx, void *const slots[12])
{
    struct SwsContext *c = ctx;
    int n = 0;
    if ((slots[2] && !c->hyscale_fast) || (slots[6] && !c->yuv2nv12cX) || (slots[7] && !c->yuv2packed1) || (slots[8] && !c->yuv2packed2) ||
        (slots[9] && !c->yuv2packedX) || (slots[10] && !c->lumConvertRange)) return -1;
    if (slots[0])  { c->hyScale = slots[0]; n++; }
    if (slots[1])  { c->hcScale = slots[1]; n++; }
    if (slots[2])  { c->hyscale_fast = slots[2]; n++; }
    if (slots[3])  { c->hcscale_fast = slots[3]; n++; }
    if (slots[4])  { c->yuv2plane1 = slots[4]; n++; }
    if (slots[5])  { c->yuv2planeX = slots[5]; n++; }
    if (slots[6])  { c->yuv2nv12cX = slots[6]; n++; }
    if (slots[7])  { c->yuv2packed1 = slots[7]; n++; }
    if (slots[8])  { c->yuv2packed2 = slots[8]; n++; }
    if (slots[9])  { c->yuv2packedX = slots[9]; n++; }
    if (slots[10]) { c->lumConvertRange = slots[10]; n++; }
    if (slots[11]) { c->chrConvertRange = slots[11]; n++; }
    /* which slots the reference has installed itself (bit i = slot i non-NULL): the product's table must install the same set */
    return n;
}
int ref_sws_slot_mask(void *ctx)
{
    struct SwsContext *c = ctx;
    return (c->hyScale ? 1 : 0) | (c->hcScale ? 2 : 0) | (c->hyscale_fast ? 4 : 0) | (c->hcscale_fast ? 8 : 0) | (c->yuv2plane1 ? 16 : 0) |
           (c->yuv2planeX ? 32 : 0) | (c->yuv2nv12cX ? 64 : 0) | (c->yuv2packed1 ? 128 : 0) | (c->yuv2packed2 ? 256 : 0) | (c->yuv2packedX ? 512 : 0) |
           (c->lumConvertRange ? 1024 : 0) | (c->chrConvertRange ? 2048 : 0);
}
int ref_sws_run(void *ctx, const uint8_t *const src[3], const int ss[3], int sh, uint8_t *const dst[3], const int dstride[3])
{
    const uint8_t *s[4] = { src[0], src[1], src[2], NULL };
    int sst[4] = { ss[0], ss[1], ss[2], 0 }, ds[4] = { dstride[0], dstride[1], dstride[2], 0 };
    uint8_t *d[4] = { dst[0], dst[1], dst[2], NULL };
    return sws_scale(ctx, s, sst, 0, sh, d, ds);
}
/* one slice: src[] points at source row y0 (chroma row y0 >> chrSrcVSubSample), dst[] at the top of the picture */
int ref_sws_run_slice(void *ctx, const uint8_t *const src[3], const int ss[3], int y0, int sh, uint8_t *const dst[3], const int dstride[3])
{
    const uint8_t *s[4] = { src[0], src[1], src[2], NULL };
    int sst[4] = { ss[0], ss[1], ss[2], 0 }, ds[4] = { dstride[0], dstride[1], dstride[2], 0 };
    uint8_t *d[4] = { dst[0], dst[1], dst[2], NULL };
    return sws_scale(ctx, s, sst, y0, sh, d, ds);
}
void ref_sws_close(void *ctx) { sws_freeContext(ctx); }

int ref_sws_get_filter(int which, int to_rgb, int sw, int sh, int dw, int dh, int flags, int16_t *filter,
                       int32_t *pos, int cap, int *n_out)
{
    struct SwsContext *c = mk_sws(sw, sh, dw, dh, to_rgb ? AV_PIX_FMT_RGB24 : AV_PIX_FMT_YUV420P, flags);
    if (!c) return -1;
    int16_t *f; int32_t *p; int fs, n;
    switch (which) {
    case 0: f = c->hLumFilter; p = c->hLumFilterPos; fs = c->hLumFilterSize; n = c->dstW; break;
    case 1: f = c->hChrFilter; p = c->hChrFilterPos; fs = c->hChrFilterSize; n = c->chrDstW; break;
    case 2: f = c->vLumFilter; p = c->vLumFilterPos; fs = c->vLumFilterSize; n = c->dstH; break;
    default: f = c->vChrFilter; p = c->vChrFilterPos; fs = c->vChrFilterSize; n = c->chrDstH; break;
    }
    *n_out = n;
    if (!f || !p) { sws_freeContext(c); return -3; }   /* unscaled special converter: no filter banks exist */
    if (n > cap || n * fs > cap) { sws_freeContext(c); return -2; }
    memcpy(filter, f, sizeof(int16_t) * n * fs);
    memcpy(pos, p, sizeof(int32_t) * n);
    sws_freeContext(c);
    return fs;
}
void ref_sws_rgb24_tables(uint8_t *ytab, int32_t *rv, int32_t *gu, int32_t *gv, int32_t *bu)
{
    struct SwsContext *c = mk_sws(16, 16, 16, 16, AV_PIX_FMT_RGB24, SWS_BICUBIC | SWS_ACCURATE_RND | SWS_BITEXACT);
    uint8_t *base = c->yuvTable;
    memcpy(ytab, base, 1024);
    for (int i = 0; i < 256; i++) {
        rv[i] = (int32_t)(c->table_rV[i] - base);
        gu[i] = (int32_t)(c->table_gU[i] - base);
        gv[i] = c->table_gV[i];
        bu[i] = (int32_t)(c->table_bU[i] - base);
    }
    sws_freeContext(c);
}

/* ---- FFT / MDCT ---- */
void ref_fft(int nbits, int inverse, float *z)
{
    FFTContext s;
    INIT();
    ff_fft_init(&s, nbits, inverse);
    s.fft_permute(&s, (FFTComplex *)z);
    s.fft_calc(&s, (FFTComplex *)z);
    ff_fft_end(&s);
}
void ref_imdct_half(int nbits, double scale, float *out, const float *in)
{ FFTContext s; INIT(); ff_mdct_init(&s, nbits, 1, scale); s.imdct_half(&s, out, in); ff_mdct_end(&s); }
void ref_imdct_calc(int nbits, double scale, float *out, const float *in)
{ FFTContext s; INIT(); ff_mdct_init(&s, nbits, 1, scale); s.imdct_calc(&s, out, in); ff_mdct_end(&s); }
void ref_mdct_calc(int nbits, double scale, float *out, const float *in)
{ FFTContext s; INIT(); ff_mdct_init(&s, nbits, 0, scale); s.mdct_calc(&s, out, in); ff_mdct_end(&s); }

/* a live reference FFTContext for the slot tests: ff_fft_init / ff_mdct_init exactly as a codec would call them */
void *ref_fft_ctx_new(int nbits, int inverse, int mdct, double scale)
{
    FFTContext *s = av_mallocz(sizeof(*s));
    INIT();
    if ((mdct ? ff_mdct_init(s, nbits, inverse, scale) : ff_fft_init(s, nbits, inverse)) < 0) { av_free(s); return NULL; }
    return s;
}
void ref_fft_ctx_free(void *p, int mdct) { FFTContext *s = p; if (!s) return; if (mdct) ff_mdct_end(s); else ff_fft_end(s); av_free(s); }
int ref_sizeof_fftcontext(void) { return sizeof(FFTContext); }

/* ---- ABI facts of the reference's tables (sizes and a few offsets), for tests/test_abi_cpu.py ---- */
#include <stddef.h>
void ref_mpeg4_qpel(int kind, int sidx, int mc, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    INIT();
    (kind == 0 ? mqpel.put_qpel_pixels_tab : kind == 1 ? mqpel.put_no_rnd_qpel_pixels_tab : mqpel.avg_qpel_pixels_tab)[sidx][mc](dst, src, stride);
}
void ref_pixblock(int kind, int16_t *block, const uint8_t *s1, const uint8_t *s2, ptrdiff_t stride)
{ INIT(); if (kind) pixb.diff_pixels(block, s1, s2, stride); else pixb.get_pixels(block, s1, stride); }

int ref_abi_info(int32_t *out, int cap)
{
    int32_t v[] = {
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
    int n = sizeof(v) / sizeof(v[0]);
    for (int i = 0; i < n && i < cap; i++) out[i] = v[i];
    return n;
}
