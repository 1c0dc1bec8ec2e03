/*
 * oracle/refbuild/refapi_mpv.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * ref_mpeg_dequant(): the reference's own inverse quantisers.  dct_unquantize_{mpeg1,mpeg2,h263}_{intra,inter}_c are
 * static in libavcodec/mpegvideo.c (:51-270); they are reached exactly as a codec reaches them: ff_mpv_common_init()
 * (with no picture size, so nothing but dct_init() and the picture shells is set up) installs them in the
 * MpegEncContext, ff_mpv_idct_init() builds the scan tables, and the call goes through the installed pointer.
 * mpegvideo.c is compiled unmodified; the handful of decoder entry points it references but this path never calls
 * (picture pool, error resilience, motion compensation, frame threading) are satisfied by the aborting stubs in refapi_mpv_stubs.c.
 */
#define ORC_PREFIX ref_
#include "../oracle_api.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "libavutil/cpu.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/mpegvideo.h"

static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static MpegEncContext *ctx[2][2];                               /* [bitexact][alternate_scan] */

static MpegEncContext *get_ctx(int bitexact, int alternate_scan)
{
    MpegEncContext *s = ctx[bitexact][alternate_scan];
    if (!s) {
        AVCodecContext *avctx = calloc(1, sizeof(*avctx));
        av_set_cpu_flags_mask(0);
        avctx->idct_algo = FF_IDCT_SIMPLE;
        avctx->bits_per_raw_sample = 8;
        avctx->pix_fmt = AV_PIX_FMT_YUV420P;
        avctx->flags = bitexact ? AV_CODEC_FLAG_BITEXACT : 0;
        s = av_mallocz(sizeof(*s));
        s->avctx = avctx;
        s->codec_id = AV_CODEC_ID_MPEG4;
        s->alternate_scan = alternate_scan;
        if (ff_mpv_common_init(s) < 0) abort();
        ff_mpv_idct_init(s);
        ctx[bitexact][alternate_scan] = s;
    }
    return s;
}

void ref_mpeg_dequant(int kind, int16_t *block, int n, int qscale, int last_index, int y_dc_scale, int c_dc_scale,
                      const uint16_t *intra_matrix, const uint16_t *inter_matrix, int alternate_scan, int h263_aic, int ac_pred)
{
    pthread_mutex_lock(&mu);
    MpegEncContext *s = get_ctx(kind == 3, alternate_scan);
    s->block_last_index[n] = last_index;
    s->y_dc_scale = y_dc_scale; s->c_dc_scale = c_dc_scale;
    s->h263_aic = h263_aic; s->ac_pred = ac_pred;
    memcpy(s->intra_matrix, intra_matrix, sizeof(s->intra_matrix));
    memcpy(s->inter_matrix, inter_matrix, sizeof(s->inter_matrix));
    switch (kind) {
    case 0: s->dct_unquantize_mpeg1_intra(s, block, n, qscale); break;
    case 1: s->dct_unquantize_mpeg1_inter(s, block, n, qscale); break;
    case 2: case 3: s->dct_unquantize_mpeg2_intra(s, block, n, qscale); break;
    case 4: s->dct_unquantize_mpeg2_inter(s, block, n, qscale); break;
    case 5: s->dct_unquantize_h263_intra(s, block, n, qscale); break;
    case 6: s->dct_unquantize_h263_inter(s, block, n, qscale); break;
    }
    pthread_mutex_unlock(&mu);
}

/* the scan tables a caller of the batched kernels needs (ScanTable.permutated / raster_end, idctdsp.c:28-47) */
void ref_mpeg_scantables(int alternate_scan, uint8_t *permutated, uint8_t *raster_end)
{
    pthread_mutex_lock(&mu);
    MpegEncContext *s = get_ctx(0, alternate_scan);
    memcpy(permutated, s->intra_scantable.permutated, 64);
    memcpy(raster_end, s->inter_scantable.raster_end, 64);
    pthread_mutex_unlock(&mu);
}


/* Encoder-state me_cmp metrics that only dereference the context's DSP tables (me_cmp.c:538-621): dct_sad, dct_max,
 * dct264_sad.  The MpegEncContext carries pdsp / fdsp / mecc exactly as ff_mpv_encode_init() fills them
 * (mpegvideo_enc.c:742-747); fdct_sel 0 = FF_DCT_AUTO (islow), 2 = FF_DCT_FASTINT (ifast). */
#include "libavcodec/me_cmp.h"
static MpegEncContext *enc_ctx[2];
int ref_me_cmp_enc(int kind, int sidx, int fdct_sel, uint8_t *b1, uint8_t *b2, ptrdiff_t stride, int h)
{
    pthread_mutex_lock(&mu);
    MpegEncContext *s = enc_ctx[fdct_sel == 2];
    if (!s) {
        AVCodecContext *avctx = calloc(1, sizeof(*avctx));
        av_set_cpu_flags_mask(0);
        avctx->bits_per_raw_sample = 8;
        avctx->dct_algo = fdct_sel == 2 ? FF_DCT_FASTINT : FF_DCT_AUTO;
        s = av_mallocz(sizeof(*s));
        s->avctx = avctx;
        ff_pixblockdsp_init(&s->pdsp, avctx);
        ff_fdctdsp_init(&s->fdsp, avctx);
        ff_me_cmp_init_static();
        ff_me_cmp_init(&s->mecc, avctx);
        enc_ctx[fdct_sel == 2] = s;
    }
    me_cmp_func f = kind == 11 ? s->mecc.dct_sad[sidx] : kind == 12 ? s->mecc.dct_max[sidx] : s->mecc.dct264_sad[sidx];
    int r = f ? f(s, b1, b2, stride, h) : -1;
    pthread_mutex_unlock(&mu);
    return r;
}
