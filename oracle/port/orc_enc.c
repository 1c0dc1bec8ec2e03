/*
 * oracle/port/orc_enc.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Restatement of the three MECmpContext metrics that run the encoder's quantiser (libavcodec/me_cmp.c:621-782):
 *   quant_psnr8x8_c (:621-645)  difference -> quantise as an inter block -> inverse quantise -> simple IDCT -> squared error against the difference
 *   bit8x8_c        (:713-782)  difference -> quantise -> VLC bit count from the codec's run/level length tables
 *   rd8x8_c         (:647-711)  both: bits weighted by qscale^2 * 109 / 128 + squared error of the reconstruction
 * plus the 16-wide wrappers (WRAPPER8_16_SQ, :859-885).  The quantiser is ff_dct_quantize_c (libavcodec/mpegvideo_enc.c:4371-4450) stated
 * per coefficient: the forward DCT, then for the DC of an intra block a rounded division by 8 * dc_scale, for every other coefficient
 * "level = coef * qmat; kept when |level| exceeds (1 << 22) - bias - 1, as (bias + |level|) >> 22 with the sign restored"; the last kept
 * scan position is the return value.  qmat = ff_convert_matrix() (mpegvideo_enc.c:84-160).  Pinned against the compiled reference
 * (oracle/refbuild/refapi_enc.c) in tests/test_oracle_enc_cpu.py.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../oracle_api.h"

enum { QSHIFT = 22, BIAS_SHIFT = 8 };       /* QMAT_SHIFT, QUANT_BIAS_SHIFT (mpegvideo_enc.c:65-68) */

/* AAN scale factors, 14 fractional bits: s(0) = 1, s(k) = cos(k pi / 16) sqrt(2); table[8 u + v] = round(s(u) s(v) 2^14) (aandcttab.c:25-35) */
static const uint16_t aan_row[8] = { 16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520 };
static const uint16_t aan_tab[64] = {
    16384, 22725, 21407, 19266, 16384, 12873,  8867,  4520, 22725, 31521, 29692, 26722, 22725, 17855, 12299,  6270,
    21407, 29692, 27969, 25172, 21407, 16819, 11585,  5906, 19266, 26722, 25172, 22654, 19266, 15137, 10426,  5315,
    16384, 22725, 21407, 19266, 16384, 12873,  8867,  4520, 12873, 17855, 16819, 15137, 12873, 10114,  6967,  3552,
     8867, 12299, 11585, 10426,  8867,  6967,  4799,  2446,  4520,  6270,  5906,  5315,  4520,  3552,  2446,  1247 };

void orc_enc_qmatrices(const OrcEncState *st, int32_t *q_intra, int32_t *q_inter, uint8_t *scantable)
{
    uint8_t rend[64];
    (void)aan_row;
    for (int i = 0; i < 64; i++) {
        const int64_t di = (int64_t)st->qscale * st->intra_matrix[i], dn = (int64_t)st->qscale * st->inter_matrix[i];
        if (st->fdct_sel == 2) {
            q_intra[i] = (int32_t)((UINT64_C(1) << (QSHIFT + 14)) / (aan_tab[i] * di));
            q_inter[i] = (int32_t)((UINT64_C(1) << (QSHIFT + 14)) / (aan_tab[i] * dn));
        } else {
            q_intra[i] = (int32_t)((UINT64_C(1) << QSHIFT) / di);
            q_inter[i] = (int32_t)((UINT64_C(1) << QSHIFT) / dn);
        }
    }
    orc_mpeg_scantables(st->alternate_scan, scantable, rend);
}

/* ff_dct_quantize_c on an already differenced block; returns last_non_zero */
static int quantise(const OrcEncState *st, int16_t *b, int intra, const int32_t *qi, const int32_t *qn, const uint8_t *scan)
{
    orc_fdct(st->fdct_sel == 2 ? 2 : 0, b);
    const int32_t *qmat = intra ? qi : qn;
    const int bias = (intra ? st->intra_quant_bias : st->inter_quant_bias) * (1 << (QSHIFT - BIAS_SHIFT));
    const uint32_t t1 = (1u << QSHIFT) - (uint32_t)bias - 1u, t2 = t1 << 1;
    int last = intra ? 0 : -1;
    if (intra) {
        const int q = (st->h263_aic ? 1 : st->y_dc_scale) << 3;
        b[0] = (int16_t)((b[0] + (q >> 1)) / q);
    }
    for (int i = intra; i < 64; i++) {
        const int j = scan[i];
        const int32_t level = (int32_t)((uint32_t)(int32_t)b[j] * (uint32_t)qmat[j]);
        if ((uint32_t)level + t1 > t2) {
            const int32_t m = level > 0 ? (int32_t)((uint32_t)bias + (uint32_t)level) >> QSHIFT : (int32_t)((uint32_t)bias - (uint32_t)level) >> QSHIFT;
            b[j] = (int16_t)(level > 0 ? m : -m);
            last = i;
        } else b[j] = 0;
    }
    return last;
}

static int count_bits(const OrcEncState *st, const int16_t *b, int intra, int last, const uint8_t *scan)
{
    const uint8_t *len = intra ? st->intra_ac_vlc_length : st->inter_ac_vlc_length;
    const uint8_t *len_last = intra ? st->intra_ac_vlc_last_length : st->inter_ac_vlc_last_length;
    int bits = intra ? st->luma_dc_vlc_length[b[0] + 256] : 0, run = 0;
    for (int i = intra; i <= last; i++) {
        const int level = b[scan[i]];
        if (!level && i < last) { run++; continue; }
        const unsigned idx = (unsigned)(level + 64);
        bits += idx < 128 ? (i == last ? len_last : len)[run * 128 + idx] : st->ac_esc_length;
        run = 0;
    }
    return bits;
}

static void unquantise(const OrcEncState *st, int16_t *b, int intra, int last)
{
    static const int kinds[4][2] = { { 1, 0 }, { 4, 2 }, { 4, 3 }, { 6, 5 } };      /* [family][intra] -> orc_mpeg_dequant kind */
    if (st->dequant == 3 && last < 0) return;     /* (the C function would index raster_end[-1]; every level is zero, nothing changes) */
    orc_mpeg_dequant(kinds[st->dequant & 3][intra], b, 0, st->qscale, last, st->y_dc_scale, st->c_dc_scale, st->intra_matrix, st->inter_matrix,
                     st->alternate_scan, st->h263_aic, st->ac_pred);
}

static int one_block(int kind, const OrcEncState *st, const uint8_t *s1, const uint8_t *s2, ptrdiff_t stride, int32_t *side,
                     const int32_t *qi, const int32_t *qn, const uint8_t *scan)
{
    int16_t t[64], bak[64];
    uint8_t p1[64], p2[64];
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        p1[8 * y + x] = s1[y * stride + x]; p2[8 * y + x] = s2[y * stride + x];
        bak[8 * y + x] = t[8 * y + x] = (int16_t)(p1[8 * y + x] - p2[8 * y + x]);
    }
    const int intra = kind == 14 ? 0 : st->mb_intra != 0;
    const int last = quantise(st, t, intra, qi, qn, scan);
    if (side) { side[0] = last; if (kind == 14) side[1] = 0; }
    if (kind == 14) {
        unquantise(st, t, 0, last);
        orc_simple_idct(t);
        int sum = 0;
        for (int i = 0; i < 64; i++) sum += (t[i] - bak[i]) * (t[i] - bak[i]);
        return sum;
    }
    const int bits = count_bits(st, t, intra, last, scan);
    if (kind == 15) return bits;
    if (last >= 0) unquantise(st, t, intra, last);
    orc_simple_idct_add(p2, 8, t);
    int dist = 0;
    for (int i = 0; i < 64; i++) dist += (p2[i] - p1[i]) * (p2[i] - p1[i]);
    return dist + ((bits * st->qscale * st->qscale * 109 + 64) >> 7);
}

int orc_me_cmp_quant(int kind, int sidx, const OrcEncState *st, uint8_t *b1, uint8_t *b2, ptrdiff_t stride, int h, int32_t *side)
{
    int32_t qi[64], qn[64];
    uint8_t scan[64];
    if (kind < 14 || kind > 16 || sidx < 0 || sidx > 1 || st->qscale < 1 || st->qscale > 31) return -1;
    orc_enc_qmatrices(st, qi, qn, scan);
    if (side) { side[0] = -2; side[1] = st->mb_intra; }
    if (sidx == 1) return one_block(kind, st, b1, b2, stride, side, qi, qn, scan);
    int score = 0;
    for (int k = 0; k < (h == 16 ? 4 : 2); k++)
        score += one_block(kind, st, b1 + 8 * (k & 1) + 8 * (k >> 1) * stride, b2 + 8 * (k & 1) + 8 * (k >> 1) * stride, stride, side, qi, qn, scan);
    return score;
}
