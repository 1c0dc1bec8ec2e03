/*
 * oracle/refbuild/refapi_enc.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * ref_me_cmp_quant(): the reference's own quant_psnr / bit / rd comparison functions (static in libavcodec/me_cmp.c:621-782), reached
 * through the table ff_me_cmp_init() fills, on a MpegEncContext set up the way ff_mpv_encode_init() sets up the fields they read:
 * pdsp / fdsp / idsp / mecc (mpegvideo_enc.c:742-747), fast_dct_quantize = ff_dct_quantize_c (:773-777), the inverse quantiser pair by
 * codec family (:1613-1622, installed by ff_mpv_common_init), q_intra_matrix / q_inter_matrix from ff_convert_matrix() (:842-849).
 * libavcodec/mpegvideo_enc.c is compiled unmodified; the encoder entry points it references that this path never reaches (rate control,
 * motion estimation, packet plumbing) are aborting stubs in refapi_mpv_stubs.c.  The VLC length tables are the caller's (a codec's
 * static tables in a real encoder; the functions only index them).
 */
#define ORC_PREFIX ref_
#include "../oracle_api.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "libavutil/cpu.h"
#include "libavutil/mem.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/mpegvideo.h"
#include "libavcodec/me_cmp.h"

static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static MpegEncContext *ctxs[2][4][2];                               /* [fdct ifast][dequant family][alternate_scan] */

static MpegEncContext *get_ctx(const OrcEncState *st)
{
    const int fi = st->fdct_sel == 2, fam = st->dequant & 3, alt = !!st->alternate_scan;
    MpegEncContext *s = ctxs[fi][fam][alt];
    if (!s) {
        AVCodecContext *avctx = calloc(1, sizeof(*avctx));
        av_set_cpu_flags_mask(0);
        avctx->idct_algo = FF_IDCT_SIMPLE;
        avctx->dct_algo = fi ? FF_DCT_FASTINT : FF_DCT_AUTO;
        avctx->bits_per_raw_sample = 8;
        avctx->pix_fmt = AV_PIX_FMT_YUV420P;
        avctx->flags = fam == 2 ? AV_CODEC_FLAG_BITEXACT : 0;
        s = av_mallocz(sizeof(*s));
        s->avctx = avctx;
        s->codec_id = fam == 1 || fam == 2 ? AV_CODEC_ID_MPEG2VIDEO : fam == 3 ? AV_CODEC_ID_MPEG4 : AV_CODEC_ID_MPEG1VIDEO;
        s->out_format = fam == 3 ? FMT_H263 : FMT_MPEG1;
        s->alternate_scan = alt;
        if (ff_mpv_common_init(s) < 0) abort();                    /* dct_init(): the seven dct_unquantize_*_c */
        ff_mpv_idct_init(s);                                       /* idsp + scan tables */
        ff_pixblockdsp_init(&s->pdsp, avctx);
        ff_fdctdsp_init(&s->fdsp, avctx);
        ff_me_cmp_init_static();
        ff_me_cmp_init(&s->mecc, avctx);
        s->fast_dct_quantize = s->dct_quantize = ff_dct_quantize_c;
        if (fam == 1 || fam == 2) { s->dct_unquantize_intra = s->dct_unquantize_mpeg2_intra; s->dct_unquantize_inter = s->dct_unquantize_mpeg2_inter; }
        else if (fam == 3)        { s->dct_unquantize_intra = s->dct_unquantize_h263_intra;  s->dct_unquantize_inter = s->dct_unquantize_h263_inter; }
        else                      { s->dct_unquantize_intra = s->dct_unquantize_mpeg1_intra; s->dct_unquantize_inter = s->dct_unquantize_mpeg1_inter; }
        s->q_intra_matrix = av_mallocz(32 * 64 * sizeof(int));
        s->q_inter_matrix = av_mallocz(32 * 64 * sizeof(int));
        s->q_intra_matrix16 = av_mallocz(32 * 2 * 64 * sizeof(uint16_t));
        s->q_inter_matrix16 = av_mallocz(32 * 2 * 64 * sizeof(uint16_t));
        s->max_qcoeff = 127; s->min_qcoeff = -127;
        ctxs[fi][fam][alt] = s;
    }
    return s;
}

static void load_state(MpegEncContext *s, const OrcEncState *st)
{
    s->qscale = st->qscale; s->mb_intra = st->mb_intra;
    s->y_dc_scale = st->y_dc_scale; s->c_dc_scale = st->c_dc_scale;
    s->h263_aic = st->h263_aic; s->ac_pred = st->ac_pred;
    s->intra_quant_bias = st->intra_quant_bias; s->inter_quant_bias = st->inter_quant_bias;
    s->ac_esc_length = st->ac_esc_length;
    memcpy(s->intra_matrix, st->intra_matrix, sizeof(s->intra_matrix));
    memcpy(s->inter_matrix, st->inter_matrix, sizeof(s->inter_matrix));
    ff_convert_matrix(s, s->q_intra_matrix, s->q_intra_matrix16, s->intra_matrix, s->intra_quant_bias, st->qscale, st->qscale, 1);
    ff_convert_matrix(s, s->q_inter_matrix, s->q_inter_matrix16, s->inter_matrix, s->inter_quant_bias, st->qscale, st->qscale, 0);
    s->intra_ac_vlc_length = (uint8_t *)st->intra_ac_vlc_length; s->intra_ac_vlc_last_length = (uint8_t *)st->intra_ac_vlc_last_length;
    s->inter_ac_vlc_length = (uint8_t *)st->inter_ac_vlc_length; s->inter_ac_vlc_last_length = (uint8_t *)st->inter_ac_vlc_last_length;
    s->luma_dc_vlc_length = (uint8_t *)st->luma_dc_vlc_length;
}

int ref_me_cmp_quant(int kind, int sidx, const OrcEncState *st, uint8_t *b1, uint8_t *b2, ptrdiff_t stride, int h, int32_t *side)
{
    if (st->qscale < 1 || st->qscale > 31 || sidx < 0 || sidx > 1) return -1;
    pthread_mutex_lock(&mu);
    MpegEncContext *s = get_ctx(st);
    load_state(s, st);
    s->block_last_index[0] = -2;
    me_cmp_func f = kind == 14 ? s->mecc.quant_psnr[sidx] : kind == 15 ? s->mecc.bit[sidx] : kind == 16 ? s->mecc.rd[sidx] : NULL;
    int r = f ? f(s, b1, b2, stride, h) : -1;
    if (side) { side[0] = s->block_last_index[0]; side[1] = s->mb_intra; }
    pthread_mutex_unlock(&mu);
    return r;
}

void ref_enc_qmatrices(const OrcEncState *st, int32_t *q_intra, int32_t *q_inter, uint8_t *scantable)
{
    pthread_mutex_lock(&mu);
    MpegEncContext *s = get_ctx(st);
    load_state(s, st);
    memcpy(q_intra, s->q_intra_matrix[st->qscale], 64 * sizeof(int));
    memcpy(q_inter, s->q_inter_matrix[st->qscale], 64 * sizeof(int));
    memcpy(scantable, s->intra_scantable.scantable, 64);
    pthread_mutex_unlock(&mu);
}
