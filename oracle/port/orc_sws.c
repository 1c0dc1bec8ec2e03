/*
 * oracle/port/orc_sws.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of libswscale's scaler for yuv420p sources, 8-bit, as selected by
 * SWS_ACCURATE_RND | SWS_BITEXACT:
 *   filter design           libswscale/utils.c:249-632   (initFilter)
 *   geometry                libswscale/utils.c:887-1198  (sws_init_context)
 *   line scheduling, edges  libswscale/swscale.c:450-682 (swscale)
 *   horizontal pass         libswscale/swscale.c:133-147 (hScale8To15_c)
 *   packed RGB output       libswscale/output.c:936-1110 (yuv2rgb_{X,2,1}_c_template) + :853-866
 *   planar output           libswscale/output.c:242-265  (yuv2planeX_8_c, yuv2plane1_8_c)
 *   yuv->rgb tables         libswscale/yuv2rgb.c:633-658, :671-863 (24 bpp)
 * Unlike the product (which evaluates the table arithmetically) this restatement builds the
 * 1024-entry table and the four 256-entry offset tables and indexes them like the reference.
 * Pinned against oracle/_ref (tests/test_oracle_sws_cpu.py) and the CRC known answers in tests/golden.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../oracle_api.h"

#define F_FAST_BILINEAR 1
#define F_BILINEAR 2
#define F_BICUBIC 4
#define F_X 8
#define F_POINT 0x10
#define F_AREA 0x20
#define F_BICUBLIN 0x40
#define F_GAUSS 0x80
#define F_SINC 0x100
#define F_LANCZOS 0x200
#define F_SPLINE 0x400
#define F_FULL_CHR_H_INT 0x2000
#define F_ACCURATE_RND 0x40000
#define F_BITEXACT 0x80000
#define PARAM_DEFAULT 123456.0

typedef struct { int taps, n; int16_t *coef; int32_t *pos; } bank_t;
typedef struct {
    int srcW, srcH, dstW, dstH, chrSrcW, chrSrcH, chrDstW, chrDstH, flags, rgb, lumXInc, chrXInc;
    bank_t hl, hc, vl, vc;
} sws_t;

static const int64_t ONE = 1LL << 54;
static int64_t ll_abs(int64_t v) { return v < 0 ? -v : v; }

static double spline(double a, double b, double c, double d, double x)
{
    if (x <= 1.0) return ((d * x + c) * x + b) * x + a;
    return spline(0.0, b + 2.0 * c + 3.0 * d, c + 3.0 * d, -b - 3.0 * c - 6.0 * d, x - 1.0);
}

static int64_t tap_weight(int flags, int64_t d, int xinc)
{
    double fd = d * (1.0 / (1 << 30));
    int64_t c;
    if (flags & F_BICUBIC) {
        int64_t B = 0, C = (int64_t)(0.6 * (1 << 24));
        if (d >= 1LL << 31) c = 0;
        else {
            int64_t dd = (d * d) >> 30, ddd = (dd * d) >> 30;
            if (d < 1LL << 30)
                c = (12 * (1 << 24) - 9 * B - 6 * C) * ddd + (-18 * (1 << 24) + 12 * B + 6 * C) * dd + (6 * (1 << 24) - 2 * B) * (1LL << 30);
            else
                c = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd + (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1LL << 30);
        }
        return c * (ONE >> 54);
    } else if (flags & F_X) {
        double v = fd < 1.0 ? cos(fd * M_PI) : -1.0;
        v = v < 0.0 ? -pow(-v, 1.0) : pow(v, 1.0);
        return (int64_t)((v * 0.5 + 0.5) * ONE);
    } else if (flags & F_AREA) {
        int64_t d2 = d - (1 << 29);
        if (d2 * xinc < -(1LL << 45)) c = 1LL << 46;
        else if (d2 * xinc < (1LL << 45)) c = -d2 * xinc + (1LL << 45);
        else c = 0;
        return c * (ONE >> 46);
    } else if (flags & F_GAUSS) {
        return (int64_t)(pow(2.0, -3.0 * fd * fd) * ONE);
    } else if (flags & F_SINC) {
        return (int64_t)((d ? sin(fd * M_PI) / (fd * M_PI) : 1.0) * ONE);
    } else if (flags & F_LANCZOS) {
        c = (int64_t)((d ? sin(fd * M_PI) * sin(fd * M_PI / 3.0) / (fd * fd * M_PI * M_PI / 3.0) : 1.0) * ONE);
        return fd > 3.0 ? 0 : c;
    } else if (flags & F_BILINEAR) {
        c = (1 << 30) - d;
        return (c < 0 ? 0 : c) * (ONE >> 30);
    } else if (flags & F_SPLINE) {
        double p = -2.196152422706632;
        return (int64_t)(spline(1.0, 0.0, p, -p - 1.0, fd) * ONE);
    }
    return 0;
}

/* SwsFilter vectors (libswscale/swscale.h:106-117), set for the calls that follow: which 0 lumH, 1 lumV, 2 chrH, 3 chrV; side 0 = srcFilter,
 * 1 = dstFilter; length 0 clears.  initFilter() convolves the source-side vector into every row and lets the destination-side one only widen
 * the rows (utils.c:444-474); with any vector longer than 1 no unscaled special converter is installed (:974-981,1043). */
static __thread struct { const double *coeff; int length; } g_fv[4][2];
static int uses_filter(void)
{
    for (int w = 0; w < 4; w++) for (int s = 0; s < 2; s++) if (g_fv[w][s].length > 1) return 1;
    return 0;
}
void orc_sws_set_filter(int which, int side, const double *coeff, int length)
{
    if (which < 0 || which > 3 || side < 0 || side > 1) return;
    g_fv[which][side].coeff = length > 0 ? coeff : NULL; g_fv[which][side].length = length > 0 ? length : 0;
}

static int make_bank(bank_t *b, int xinc, int slen, int dlen, int one, int flags, int horiz, int which)
{
    int taps, i, j, k;
    int64_t *w;
    int32_t *pos = calloc(dlen + 1, sizeof(*pos));
    if (abs(xinc - 0x10000) < 10) {
        taps = 1; w = calloc(dlen, sizeof(*w));
        for (i = 0; i < dlen; i++) { w[i] = ONE; pos[i] = i; }
    } else if (flags & F_POINT) {
        int x = xinc / 2 - 0x8000;
        taps = 1; w = calloc(dlen, sizeof(*w));
        for (i = 0; i < dlen; i++) { pos[i] = (x + 0x8000) >> 16; w[i] = ONE; x += xinc; }
    } else if ((xinc <= 65536 && (flags & F_AREA)) || (flags & F_FAST_BILINEAR)) {
        int x = xinc / 2 - 0x8000;
        taps = 2; w = calloc(2 * dlen, sizeof(*w));
        for (i = 0; i < dlen; i++) {
            int xx = x >> 16;
            pos[i] = xx;
            for (j = 0; j < 2; j++) {
                int dist = (int)(((unsigned)xx << 16) - (unsigned)x);
                int64_t c = ONE - (int64_t)abs(dist) * (ONE >> 16);
                w[2 * i + j] = c < 0 ? 0 : c;
                xx++;
            }
            x += xinc;
        }
    } else {
        int sf = (flags & F_BICUBIC) ? 4 : (flags & F_X) ? 8 : (flags & F_AREA) ? 1 : (flags & F_GAUSS) ? 8 :
                 (flags & F_LANCZOS) ? 6 : (flags & F_SINC) ? 20 : (flags & F_SPLINE) ? 20 : (flags & F_BILINEAR) ? 2 : 0;
        int64_t x = xinc - 0x10000;
        if (!sf) { free(pos); return -1; }
        taps = xinc <= 65536 ? 1 + sf : 1 + (sf * slen + dlen - 1) / dlen;
        if (taps > slen - 2) taps = slen - 2;
        if (taps < 1) taps = 1;
        w = calloc((size_t)taps * dlen, sizeof(*w));
        for (i = 0; i < dlen; i++) {
            int xx = (int)((x - ((int64_t)(taps - 2) << 16)) / (1 << 17));
            pos[i] = xx;
            for (j = 0; j < taps; j++) {
                int64_t d = ll_abs(((int64_t)xx << 17) - x) << 13;
                if (xinc > 65536) d = d * dlen / slen;
                w[(size_t)i * taps + j] = tap_weight(flags, d, xinc);
                xx++;
            }
            x += 2 * (int64_t)xinc;
        }
    }
    if (g_fv[which][0].length > 0 || g_fv[which][1].length > 0) {          /* utils.c:444-474 */
        const int ns = g_fv[which][0].length, nd = g_fv[which][1].length;
        const int taps2 = taps + (ns > 0 ? ns - 1 : 0) + (nd > 0 ? nd - 1 : 0);
        int64_t *w2 = calloc((size_t)taps2 * dlen, sizeof(*w2));
        for (i = 0; i < dlen; i++) {
            if (ns > 0) {
                for (k = 0; k < ns; k++)
                    for (j = 0; j < taps; j++) w2[(size_t)i * taps2 + k + j] += g_fv[which][0].coeff[k] * w[(size_t)i * taps + j];
            } else {
                for (j = 0; j < taps; j++) w2[(size_t)i * taps2 + j] = w[(size_t)i * taps + j];
            }
            pos[i] += (taps - 1) / 2 - (taps2 - 1) / 2;
        }
        free(w); w = w2; taps = taps2;
    }
    /* shrink (utils.c:476-520) */
    int minsize = 0;
    for (i = dlen - 1; i >= 0; i--) {
        int64_t *r = w + (size_t)i * taps, cut = 0;
        int m = taps;
        for (j = 0; j < taps; j++) {
            cut += ll_abs(r[0]);
            if (cut > 0.002 * ONE) break;
            if (i < dlen - 1 && pos[i] >= pos[i + 1]) break;
            for (k = 1; k < taps; k++) r[k - 1] = r[k];
            r[k - 1] = 0;
            pos[i]++;
        }
        cut = 0;
        for (j = taps - 1; j > 0; j--) {
            cut += ll_abs(r[j]);
            if (cut > 0.002 * ONE) break;
            m--;
        }
        if (m > minsize) minsize = m;
    }
    if (minsize < 1 || minsize >= 256) { free(w); free(pos); return -1; }
    int64_t *f = calloc((size_t)minsize * dlen, sizeof(*f));
    for (i = 0; i < dlen; i++)
        for (j = 0; j < minsize; j++) f[(size_t)i * minsize + j] = j < taps ? w[(size_t)i * taps + j] : 0;
    free(w);
    if (horiz) {
        for (i = 0; i < dlen; i++) {
            int64_t *r = f + (size_t)i * minsize;
            if (pos[i] < 0) {
                for (j = 1; j < minsize; j++) {
                    int left = j + pos[i] > 0 ? j + pos[i] : 0;
                    r[left] += r[j]; r[j] = 0;
                }
                pos[i] = 0;
            }
            if (pos[i] + minsize > slen) {
                int sh = pos[i] + minsize - slen;
                for (j = minsize - 2; j >= 0; j--) {
                    int right = j + sh < minsize - 1 ? j + sh : minsize - 1;
                    r[right] += r[j]; r[j] = 0;
                }
                pos[i] = slen - minsize;
            }
        }
    }
    b->taps = minsize; b->n = dlen; b->pos = pos;
    b->coef = calloc((size_t)minsize * dlen, sizeof(int16_t));
    for (i = 0; i < dlen; i++) {
        int64_t err = 0, sum = 0;
        for (j = 0; j < minsize; j++) sum += f[(size_t)i * minsize + j];
        sum = (sum + one / 2) / one;
        for (j = 0; j < minsize; j++) {
            int64_t v = f[(size_t)i * minsize + j] + err;
            int q = (int)((v > 0 ? v + (sum >> 1) : v - (sum >> 1)) / sum);
            b->coef[(size_t)i * minsize + j] = (int16_t)q;
            err = v - q * sum;
        }
    }
    free(f);
    return 0;
}

static void free_bank(bank_t *b) { free(b->coef); free(b->pos); memset(b, 0, sizeof(*b)); }
static void sws_close(sws_t *c) { free_bank(&c->hl); free_bank(&c->hc); free_bank(&c->vl); free_bank(&c->vc); }

/* chroma sub-sampling of the planar 8-bit source being converted (log2): 4:2:0 unless orc_sws_planar() says otherwise */
static __thread int g_hs = 1, g_vs = 1;
/* the planar destination: chroma sub-sampling (log2) and sample depth (8, or 9 / 10 in little-endian 16-bit samples) */
static __thread int g_dhs = 1, g_dvs = 1, g_dbits = 8, g_dbe = 0;
static void put16(uint8_t *p, int v) { if (g_dbe) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; } else { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); } }
#define IS_RGB16(f) (((f) >= 36 && (f) <= 43) || ((f) >= 54 && (f) <= 57))     /* rgb565 / 555 / bgr565 / 555 (36-43), rgb444 / bgr444 (54-57), LE and BE */
static __thread int g_rgb16;        /* 15 / 16 / 12-bpp destination of the "rgb" entry point: 1 rgb565 2 bgr565 3 rgb555 4 bgr555 5 rgb444 6 bgr444, + 8 big-endian (output.c:869-902) */
#define IS_RGB48(f) ((f) == 34 || (f) == 35 || (f) == 59 || (f) == 60)             /* rgb48be 34, rgb48le 35, bgr48be 59, bgr48le 60 */
static __thread int g_rgb48;        /* 48-bit destination of the "rgb" entry point: 1 rgb48 2 bgr48, + 8 big-endian (yuv2rgb48_X / _2 / _1, output.c:584-760) */
static __thread int g_src_gray;     /* gray8 source: swscale() never converts chroma lines for it (needs_hcscale, swscale.c:532,768-770), so the vertical stage reads what
                                     * sws_init_context left in the line buffers: bytes of 64 (utils.c:1273), i.e. 0x4040 per 15-bit sample, 0x40404040 per 19-bit sample */
static void gray_lines(int16_t *u, int16_t *v, size_t n) { for (size_t i = 0; g_src_gray && i < n; i++) u[i] = v[i] = 0x4040; }
static void gray_lines19(int32_t *u, int32_t *v, size_t n) { for (size_t i = 0; g_src_gray && i < n; i++) u[i] = v[i] = 0x40404040; }
static __thread int g_pk422;        /* packed 4:2:2 destination of the "rgb" entry point: 1 yuyv422, 2 uyvy422 (yuv2422_X / _2 / _1, output.c:448-576) */
static __thread int g_nospecial;    /* an nv12 / nv21 destination computed through the planar path: no yuv420p-only special converters */
static __thread int g_nocopy;       /* nv12 / nv21 sources never get planarCopyWrapper (swscale_unscaled.c:1158-1170) */

static int sws_open(sws_t *c, int sw, int sh, int dw, int dh, int rgb, int flags)
{
    memset(c, 0, sizeof(*c));
    int algo = flags & (F_POINT | F_AREA | F_BILINEAR | F_FAST_BILINEAR | F_BICUBIC | F_X | F_GAUSS | F_LANCZOS | F_SINC | F_SPLINE | F_BICUBLIN);
    if (!algo) flags |= (dw < sw && dh < sh) ? F_GAUSS : (dw > sw && dh > sh) ? F_SINC : F_LANCZOS;
    else if (algo & (algo - 1)) return -1;
    if (sw < 4 || sh < 1 || dw < 8 || dh < 1 || ((flags & F_FULL_CHR_H_INT) && !rgb)) return -1;
    if (flags & 0x30000) return -1;                          /* SWS_SRC_V_CHR_DROP_MASK: not restated */
    c->srcW = sw; c->srcH = sh; c->dstW = dw; c->dstH = dh; c->flags = flags; c->rgb = rgb;
    c->chrSrcW = -((-sw) >> g_hs); c->chrSrcH = -((-sh) >> g_vs);
    /* packed RGB shares a chroma sample between two pixels unless SWS_FULL_CHR_H_INT asks for one per pixel (utils.c:998-1014) */
    c->chrDstW = rgb ? ((flags & F_FULL_CHR_H_INT) ? dw : (dw + 1) >> 1) : -((-dw) >> g_dhs); c->chrDstH = rgb ? dh : -((-dh) >> g_dvs);
    int lx = (int)((((int64_t)sw << 16) + (dw >> 1)) / dw), ly = (int)((((int64_t)sh << 16) + (dh >> 1)) / dh);
    int cx = (int)((((int64_t)c->chrSrcW << 16) + (c->chrDstW >> 1)) / c->chrDstW);
    int cyi = (int)((((int64_t)c->chrSrcH << 16) + (c->chrDstH >> 1)) / c->chrDstH);
    c->lumXInc = lx; c->chrXInc = cx;
    int lf = (flags & F_BICUBLIN) ? (flags | F_BICUBIC) : flags, cf = (flags & F_BICUBLIN) ? (flags | F_BILINEAR) : flags;
    if (make_bank(&c->hl, lx, sw, dw, 1 << 14, lf, 1, 0) || make_bank(&c->hc, cx, c->chrSrcW, c->chrDstW, 1 << 14, cf, 1, 2) ||
        make_bank(&c->vl, ly, sh, dh, 1 << 12, lf, 0, 1) || make_bank(&c->vc, cyi, c->chrSrcH, c->chrDstH, 1 << 12, cf, 0, 3)) {
        sws_close(c);
        return -1;
    }
    return 0;
}

int orc_sws_get_filter(int which, int to_rgb, int sw, int sh, int dw, int dh, int flags, int16_t *filter,
                       int32_t *pos, int cap, int *n_out)
{
    sws_t c;
    if (sws_open(&c, sw, sh, dw, dh, to_rgb, flags)) return -1;
    bank_t *b = which == 0 ? &c.hl : which == 1 ? &c.hc : which == 2 ? &c.vl : &c.vc;
    int taps = b->taps;
    *n_out = b->n;
    if (b->n > cap || b->n * taps > cap) taps = -2;
    else { memcpy(filter, b->coef, sizeof(int16_t) * b->n * taps); memcpy(pos, b->pos, sizeof(int32_t) * b->n); }
    sws_close(&c);
    return taps;
}

/* the colour settings of sws_setColorspaceDetails (utils.c:807-835); defaults = sws_getContext's (ITU601, limited range, neutral) */
static __thread struct { int inv[4], full_range, brightness, contrast, saturation; } g_cs = { { 104597, 132201, 25675, 53279 }, 0, 0, 1 << 16, 1 << 16 };
static __thread int g_cs_jpeg;       /* a yuvj source format forces full range for one call */

void orc_sws_set_colorspace(const int inv_table[4], int src_range, int brightness, int contrast, int saturation)
{
    static const int itu601[4] = { 104597, 132201, 25675, 53279 };
    memcpy(g_cs.inv, inv_table ? inv_table : itu601, sizeof(g_cs.inv));
    g_cs.full_range = inv_table ? src_range != 0 : 0; g_cs.brightness = inv_table ? brightness : 0;
    g_cs.contrast = inv_table ? contrast : 1 << 16; g_cs.saturation = inv_table ? saturation : 1 << 16;
}

/* ff_yuv2rgb_c_init_tables' scalar part (yuv2rgb.c:671-734): cy, oy and the four chroma coefficients before the division by cy */
static void cs_coeffs(int64_t *cy, int64_t *oy, int64_t *crv, int64_t *cbu, int64_t *cgu, int64_t *cgv, int *yoffs)
{
    const int full = g_cs.full_range || g_cs_jpeg;
    *crv = g_cs.inv[0]; *cbu = g_cs.inv[1]; *cgu = -g_cs.inv[2]; *cgv = -g_cs.inv[3];
    *cy = 1 << 16; *oy = 0; *yoffs = full ? 384 : 326;
    if (!full) { *cy = (*cy * 255) / 219; *oy = 16 << 16; }
    else { *crv = (*crv * 224) / 255; *cbu = (*cbu * 224) / 255; *cgu = (*cgu * 224) / 255; *cgv = (*cgv * 224) / 255; }
    *cy = (*cy * g_cs.contrast) >> 16;
    *crv = (*crv * g_cs.contrast * g_cs.saturation) >> 32; *cbu = (*cbu * g_cs.contrast * g_cs.saturation) >> 32;
    *cgu = (*cgu * g_cs.contrast * g_cs.saturation) >> 32; *cgv = (*cgv * g_cs.contrast * g_cs.saturation) >> 32;
    *oy -= 256 * (int64_t)g_cs.brightness;
}

/* yuv2rgb.c:633-658, :671-863 */
void orc_sws_rgb24_tables(uint8_t *ytab, int32_t *rv, int32_t *gu, int32_t *gv, int32_t *bu)
{
    int yoffs;
    int64_t crv, cbu, cgu, cgv, cy, oy, yb;
    cs_coeffs(&cy, &oy, &crv, &cbu, &cgu, &cgv, &yoffs);
    crv = ((crv << 16) + 0x8000) / cy; cbu = ((cbu << 16) + 0x8000) / cy;
    cgu = ((cgu * 65536) + 0x8000) / cy; cgv = ((cgv * 65536) + 0x8000) / cy;
    yb = -(384 << 16) - oy;
    for (int i = 0; i < 1024; i++) {
        int64_t v = (yb + 0x8000) >> 16;
        ytab[i] = v < 0 ? 0 : v > 255 ? 255 : (uint8_t)v;
        yb += cy;
    }
    int64_t a = 0, b = 0, c = 0, d = 0;
    for (int i = 0; i < 256; i++) {
        rv[i] = yoffs - (int)(crv >> 9) + (int)(a >> 16);
        gu[i] = yoffs - (int)(cgu >> 9) + (int)(b >> 16);
        bu[i] = yoffs - (int)(cbu >> 9) + (int)(c >> 16);
        gv[i] = -(int)(cgv >> 9) + (int)(d >> 16);
        a += crv; b += cgu; c += cbu; d += cgv;
    }
}

static uint8_t u8clip(int v) { return v < 0 ? 0 : v > 255 ? 255 : (uint8_t)v; }

/* hyscale_fast_c / hcscale_fast_c (swscale.c:238-250, :286-299), used instead of the filter bank when
 * SWS_FAST_BILINEAR is set.  The reference reads src[xx + 1] even for the last source pixel, i.e. one byte
 * past the row; here that byte is defined as a copy of the last pixel (tests pad their rows the same way). */
static int16_t *hfast(const uint8_t *src, int stride, int rows, int sw, int dw, int xinc, int chroma, int *pitch)
{
    int p = dw + 2;
    int16_t *out = calloc((size_t)p * rows, sizeof(*out));
    for (int y = 0; y < rows; y++) {
        const uint8_t *s = src + (size_t)y * stride;
        unsigned xpos = 0;
        for (int i = 0; i < dw; i++) {
            unsigned xx = xpos >> 16, xa = (xpos & 0xFFFF) >> 9;
            int a = s[xx], b = s[xx + 1 < (unsigned)sw ? xx + 1 : (unsigned)sw - 1];
            out[(size_t)y * p + i] = (int16_t)(chroma ? a * (int)(xa ^ 127) + b * (int)xa : (a << 7) + (b - a) * (int)xa);
            xpos += xinc;
        }
    }
    *pitch = p;
    return out;
}

/* hScale8To15_c over a whole plane into (rows x (n + 2)) int16, trailing columns zero like the zeroed line buffers */
/* 9 / 10 / 16-bit planar sources (hScale16To15_c, swscale.c:110-131: samples in 16-bit words, shifted down by depth - 1; big-endian formats are
 * byte-swapped by the input stage first, input.c bswap16Y_c / bswap16UV_c) and the ordered dither their 8-bit planar outputs get
 * (should_dither, swscale.c:389-390,553-556): ff_dither_8x8_128[row & 7][(i + offset) & 7], a bit-interleaved 8 x 8 matrix */
static __thread int g_sbits = 8, g_sbe;
static int dither128(int row, int col)
{
    int v = 18;
    if (row & 1) v ^= 32;
    if (row & 2) v ^= 8;
    if (row & 4) v ^= 2;
    if (col & 1) v ^= 48;
    if (col & 2) v ^= 12;
    if (col & 4) v ^= 3;
    return 2 * v;
}
static int sample_at(const uint8_t *src, size_t byte_row, int x)
{
    if (g_sbits == 8) return src[byte_row + x];
    const uint8_t *p = src + byte_row + 2 * (size_t)x;
    return g_sbe ? (p[0] << 8) | p[1] : p[0] | (p[1] << 8);
}
static int16_t *hpass(const uint8_t *src, int stride, int rows, const bank_t *b, int *pitch)
{
    int p = b->n + 2;
    int16_t *out = calloc((size_t)p * rows, sizeof(*out));
    for (int y = 0; y < rows; y++)
        for (int i = 0; i < b->n; i++) {
            int v = 0;
            for (int j = 0; j < b->taps; j++) v += sample_at(src, (size_t)y * stride, b->pos[i] + j) * b->coef[(size_t)i * b->taps + j];
            v >>= g_sbits == 8 ? 7 : g_sbits - 1;
            out[(size_t)y * p + i] = (int16_t)(v > 32767 ? 32767 : v);
        }
    *pitch = p;
    return out;
}

/* hScale8To19_c (swscale.c:62-80): the lines a 16-bit destination is filtered from */
static int32_t *hpass19(const uint8_t *src, int stride, int rows, const bank_t *b, int *pitch)
{
    int p = b->n + 2;
    int32_t *out = calloc((size_t)p * rows, sizeof(*out));
    for (int y = 0; y < rows; y++)
        for (int i = 0; i < b->n; i++) {
            int v = 0;
            for (int j = 0; j < b->taps; j++) v += src[(size_t)y * stride + b->pos[i] + j] * b->coef[(size_t)i * b->taps + j];
            v >>= 3;
            out[(size_t)y * p + i] = v > (1 << 19) - 1 ? (1 << 19) - 1 : v;
        }
    *pitch = p;
    return out;
}

static int rowsel(int first, int j, int h) { int r = first + j; return r < 0 ? 0 : r > h - 1 ? h - 1 : r; }

/* Range conversion between the horizontal and the vertical pass (c->lumConvertRange / chrConvertRange, swscale.c:166-197 installed at
 * :748-765 when srcRange != dstRange and the destination is not rgb): 1 = full (yuvj) -> limited, 2 = limited -> full, 15-bit lines */
static __thread int g_range;
static int16_t range_sample(int v, int kind, int chroma)
{
    if (kind == 1) return (int16_t)(chroma ? (v * 1799 + 4081085) >> 11 : (v * 14071 + 33561947) >> 14);
    if (chroma) { v = v < 30775 ? v : 30775; return (int16_t)((v * 4663 - 9289992) >> 12); }
    v = v < 30189 ? v : 30189;
    return (int16_t)((v * 19077 - 39057361) >> 14);
}
static void range_lines(int16_t *p, int pitch, int rows, int w, int chroma)
{
    if (!g_range) return;
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < w; x++) p[(size_t)y * pitch + x] = range_sample(p[(size_t)y * pitch + x], g_range, chroma);
}

/* 48-bit rgb destinations behind swscale(): dstBpc 16, so the lines are hScale8To19_c's (swscale.c:728-741; no fast-bilinear line functions at
 * that depth) and the packed output stage is yuv2rgb48_X / _2 / _1_c_template (output.c:593-760) with swscale()'s X / 2 / 1 selection
 * (swscale.c:658-683).  int arithmetic wraps like the compiled C does (the X accumulators start at -2^30 for that reason). */
static uint32_t clip30(uint32_t a) { return (a & 0xC0000000u) ? ((a >> 31) ? 0 : 0x3FFFFFFFu) : a; }
static void put48(uint8_t *d, int k, uint32_t v30)
{
    const unsigned v = clip30(v30) >> 14;
    if (g_rgb48 & 8) { d[2 * k] = (uint8_t)(v >> 8); d[2 * k + 1] = (uint8_t)v; } else { d[2 * k] = (uint8_t)v; d[2 * k + 1] = (uint8_t)(v >> 8); }
}
static int rgb48_scaled(sws_t *c, const uint8_t *const src[3], const int ss[3], int sh, uint8_t *dst, int dstride, int dw, int dh)
{
    int lp, cp;
    int32_t *L = hpass19(src[0], ss[0], sh, &c->hl, &lp), *U = hpass19(src[1], ss[1], c->chrSrcH, &c->hc, &cp), *V = hpass19(src[2], ss[2], c->chrSrcH, &c->hc, &cp);
    gray_lines19(U, V, (size_t)cp * c->chrSrcH);
    int64_t kcy, koy, kcrv, kcbu, kcgu, kcgv; int kyoffs;
    cs_coeffs(&kcy, &koy, &kcrv, &kcbu, &kcgu, &kcgv, &kyoffs);
#define R16(f) ((int16_t)({ int r_ = (int)(((int64_t)(f) + (1 << 15)) >> 16); r_ < -0x7FFF ? -0x8000 : r_ > 0x7FFF ? 0x7FFF : r_; }))
    const int y_coeff = R16(kcy << 13), y_offset = R16(koy << 9), v2r = R16(kcrv << 13), v2g = R16(kcgv << 13), u2g = R16(kcgu << 13), u2b = R16(kcbu << 13);
#undef R16
    const int fl = c->vl.taps, fc = c->vc.taps, bgr = (g_rgb48 & 7) == 2;
    for (int y = 0; y < dh; y++) {
        const int firstL = c->vl.pos[y] > 1 - fl ? c->vl.pos[y] : 1 - fl, firstC = c->vc.pos[y] > 1 - fc ? c->vc.pos[y] : 1 - fc;
        const int16_t *lf = c->vl.coef + (size_t)y * fl, *cf = c->vc.coef + (size_t)y * fc;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int i = 0; i < (dw + 1) >> 1; i++) {
#define L19(j, x) L[(size_t)rowsel(firstL, j, sh) * lp + (x)]
#define U19(j) U[(size_t)rowsel(firstC, j, c->chrSrcH) * cp + i]
#define V19(j) V[(size_t)rowsel(firstC, j, c->chrSrcH) * cp + i]
            int32_t Y1, Y2, Uv, Vv;
            if (fl == 1 && fc <= 2) {
                const int ua = fc == 1 ? 0 : cf[1];
                Y1 = L19(0, 2 * i) >> 2; Y2 = L19(0, 2 * i + 1) >> 2;
                if (ua < 2048) { Uv = (U19(0) + (-128 * (1 << 11))) >> 2; Vv = (V19(0) + (-128 * (1 << 11))) >> 2; }
                else { Uv = (U19(0) + U19(1) + (-128 * (1 << 12))) >> 3; Vv = (V19(0) + V19(1) + (-128 * (1 << 12))) >> 3; }
            } else if (fl == 2 && fc == 2) {
                const int ya = lf[1], ua = cf[1];
                Y1 = (int32_t)((uint32_t)L19(0, 2 * i) * (uint32_t)(4096 - ya) + (uint32_t)L19(1, 2 * i) * (uint32_t)ya) >> 14;
                Y2 = (int32_t)((uint32_t)L19(0, 2 * i + 1) * (uint32_t)(4096 - ya) + (uint32_t)L19(1, 2 * i + 1) * (uint32_t)ya) >> 14;
                Uv = (int32_t)((uint32_t)U19(0) * (uint32_t)(4096 - ua) + (uint32_t)U19(1) * (uint32_t)ua + (uint32_t)(-128 * (1 << 23))) >> 14;
                Vv = (int32_t)((uint32_t)V19(0) * (uint32_t)(4096 - ua) + (uint32_t)V19(1) * (uint32_t)ua + (uint32_t)(-128 * (1 << 23))) >> 14;
            } else {
                uint32_t a1 = (uint32_t)-0x40000000, a2 = a1, au = (uint32_t)(-128 * (1 << 23)), av = au;
                for (int j = 0; j < fl; j++) { a1 += (uint32_t)L19(j, 2 * i) * (uint32_t)(int32_t)lf[j]; a2 += (uint32_t)L19(j, 2 * i + 1) * (uint32_t)(int32_t)lf[j]; }
                for (int j = 0; j < fc; j++) { au += (uint32_t)U19(j) * (uint32_t)(int32_t)cf[j]; av += (uint32_t)V19(j) * (uint32_t)(int32_t)cf[j]; }
                Y1 = ((int32_t)a1 >> 14) + 0x10000; Y2 = ((int32_t)a2 >> 14) + 0x10000; Uv = (int32_t)au >> 14; Vv = (int32_t)av >> 14;
            }
            const uint32_t y1 = (uint32_t)(Y1 - y_offset) * (uint32_t)y_coeff + (1u << 13), y2 = (uint32_t)(Y2 - y_offset) * (uint32_t)y_coeff + (1u << 13);
            const uint32_t R = (uint32_t)Vv * (uint32_t)v2r, G = (uint32_t)Vv * (uint32_t)v2g + (uint32_t)Uv * (uint32_t)u2g, B = (uint32_t)Uv * (uint32_t)u2b;
            uint8_t *q = d + 12 * i;
            put48(q, 0, (bgr ? B : R) + y1); put48(q, 1, G + y1); put48(q, 2, (bgr ? R : B) + y1);
            if (2 * i + 1 < dw || dstride >= 6 * (dw + 1)) { put48(q, 3, (bgr ? B : R) + y2); put48(q, 4, G + y2); put48(q, 5, (bgr ? R : B) + y2); }
#undef L19
#undef U19
#undef V19
        }
    }
    free(L); free(U); free(V);
    return dh;
}

int orc_sws_yuv420p_to_rgb24(const uint8_t *const src[3], const int ss[3], int sw, int sh, uint8_t *dst, int dstride,
                             int dw, int dh, int flags)
{
    sws_t c;
    if (sws_open(&c, sw, sh, dw, dh, 1, flags)) return -1;
    uint8_t ytab[1024]; int32_t rv[256], gu[256], gv[256], bu[256];
    orc_sws_rgb24_tables(ytab, rv, gu, gv, bu);
    if (g_rgb16 && sw == dw && sh == dh && !(flags & F_ACCURATE_RND) && !(dh & 1) && g_hs == 1 && g_vs <= 1 && !uses_filter() && g_sbits == 8) {
        sws_close(&c);          /* the ordered-dither table converters yuv2rgb_c_16 / _15 / _12_ordered_dither (yuv2rgb.c:377-573): not restated */
        return -1;
    }
    if (g_rgb48 && !(sw == dw && sh == dh && !(flags & F_ACCURATE_RND) && !(dh & 1) && g_hs == 1 && g_vs <= 1 && !uses_filter() && g_sbits == 8)) {
        int r48 = g_sbits == 8 ? rgb48_scaled(&c, src, ss, sh, dst, dstride, dw, dh) : -1;
        sws_close(&c);
        return r48;
    }
    if (sw == dw && sh == dh && !(flags & F_ACCURATE_RND) && !(dh & 1) && g_hs == 1 && g_vs <= 1 && !g_pk422 && !uses_filter() && g_sbits == 8) {
        /* (4:2:0 and 4:2:2 sources only, swscale_unscaled.c:1051; a 4:2:2 source has its chroma pitch doubled, yuv2rgb.c:133-136,
         * i.e. both rows of a pair read the even chroma line)
         * unscaled table converter yuv2rgb_c_24_rgb (yuv2rgb.c:126-175, :335-372; chosen at swscale_unscaled.c:1051-1055):
         * nearest chroma, dstW & ~1 pixels per row */
        for (int y = 0; y < dh; y++)
            for (int i = 0; i < dw >> 1; i++) {
                const int crow = (y >> 1) << (1 - g_vs);
                int Uv = src[1][(size_t)crow * ss[1] + i], Vv = src[2][(size_t)crow * ss[2] + i];
                const uint8_t *r = ytab + rv[Vv], *g = ytab + gu[Uv] + gv[Vv], *b = ytab + bu[Uv];
                const uint8_t *py = src[0] + (size_t)y * ss[0] + 2 * i;
                uint8_t *d = dst + (size_t)y * dstride + 6 * i;
                if (g_rgb48) {      /* yuv2rgb_c_48 / yuv2rgb_c_bgr48 (yuv2rgb.c:106-124,177-236): the 8-bit table value in both bytes of a component */
                    const uint8_t *t0 = (g_rgb48 & 7) == 2 ? b : r, *t2 = (g_rgb48 & 7) == 2 ? r : b;
                    d += 6 * i;
                    for (int e = 0; e < 2; e++) {
                        d[6 * e + 0] = d[6 * e + 1] = t0[py[e]]; d[6 * e + 2] = d[6 * e + 3] = g[py[e]]; d[6 * e + 4] = d[6 * e + 5] = t2[py[e]];
                    }
                    continue;
                }
                d[0] = r[py[0]]; d[1] = g[py[0]]; d[2] = b[py[0]]; d[3] = r[py[1]]; d[4] = g[py[1]]; d[5] = b[py[1]];
            }
        sws_close(&c);
        return dh;
    }
    int lp, cp;
    const int fast = (c.flags & F_FAST_BILINEAR) && g_sbits == 8;      /* the fast line functions only exist for 8-bit sources (swscale.c:733-739) */
    int16_t *L = fast ? hfast(src[0], ss[0], sh, sw, dw, c.lumXInc, 0, &lp) : hpass(src[0], ss[0], sh, &c.hl, &lp);
    int16_t *U = fast ? hfast(src[1], ss[1], c.chrSrcH, c.chrSrcW, c.chrDstW, c.chrXInc, 1, &cp) : hpass(src[1], ss[1], c.chrSrcH, &c.hc, &cp);
    int16_t *V = fast ? hfast(src[2], ss[2], c.chrSrcH, c.chrSrcW, c.chrDstW, c.chrXInc, 1, &cp) : hpass(src[2], ss[2], c.chrSrcH, &c.hc, &cp);
    gray_lines(U, V, (size_t)cp * c.chrSrcH);
    if (g_pk422) { range_lines(L, lp, sh, dw, 0); range_lines(U, cp, c.chrSrcH, c.chrDstW, 1); range_lines(V, cp, c.chrSrcH, c.chrDstW, 1); }    /* (a yuv destination: swscale.c:748-765) */
    const int fl = c.vl.taps, fc = c.vc.taps;
    if (flags & F_FULL_CHR_H_INT) {
        /* yuv2rgb24_full_X_c (output.c:1165-1240): one chroma sample per pixel, 30-bit fixed point colour matrix with the
         * coefficients of ff_yuv2rgb_c_init_tables (yuv2rgb.c:735-740); always the X variant (output.c:1392-1460) */
        int64_t kcy, koy, kcrv, kcbu, kcgu, kcgv; int kyoffs;
        cs_coeffs(&kcy, &koy, &kcrv, &kcbu, &kcgu, &kcgv, &kyoffs);
#define R16(f) ((int16_t)({ int r_ = (int)(((int64_t)(f) + (1 << 15)) >> 16); r_ < -0x7FFF ? -0x8000 : r_ > 0x7FFF ? 0x7FFF : r_; }))
        const int y_coeff = R16(kcy << 13), y_offset = R16(koy << 9), v2r = R16(kcrv << 13), v2g = R16(kcgv << 13),
                  u2g = R16(kcgu << 13), u2b = R16(kcbu << 13);
#undef R16
        for (int y = 0; y < dh; y++) {
            int firstL = c.vl.pos[y] > 1 - fl ? c.vl.pos[y] : 1 - fl;
            int firstC = c.vc.pos[y] > 1 - fc ? c.vc.pos[y] : 1 - fc;
            const int16_t *lf = c.vl.coef + (size_t)y * fl, *cf = c.vc.coef + (size_t)y * fc;
            uint8_t *d = dst + (size_t)y * dstride;
            for (int i = 0; i < dw; i++) {
                int Y = 0, Uv = -128 * (1 << 19), Vv = -128 * (1 << 19);
                for (int j = 0; j < fl; j++) Y += L[(size_t)rowsel(firstL, j, sh) * lp + i] * lf[j];
                for (int j = 0; j < fc; j++) {
                    Uv += U[(size_t)rowsel(firstC, j, c.chrSrcH) * cp + i] * cf[j];
                    Vv += V[(size_t)rowsel(firstC, j, c.chrSrcH) * cp + i] * cf[j];
                }
                Y >>= 10; Uv >>= 10; Vv >>= 10;
                Y = (Y - y_offset) * y_coeff + (1 << 21);
                int R = Y + Vv * v2r, G = Y + Vv * v2g + Uv * u2g, B = Y + Uv * u2b;
                if ((R | G | B) & 0xC0000000) {
                    R = R < 0 ? 0 : R > 0x3FFFFFFF ? 0x3FFFFFFF : R; G = G < 0 ? 0 : G > 0x3FFFFFFF ? 0x3FFFFFFF : G; B = B < 0 ? 0 : B > 0x3FFFFFFF ? 0x3FFFFFFF : B;
                }
                d[3 * i] = (uint8_t)(R >> 22); d[3 * i + 1] = (uint8_t)(G >> 22); d[3 * i + 2] = (uint8_t)(B >> 22);
            }
        }
        free(L); free(U); free(V);
        sws_close(&c);
        return dh;
    }
    for (int y = 0; y < dh; y++) {
        int firstL = c.vl.pos[y] > 1 - fl ? c.vl.pos[y] : 1 - fl;
        int firstC = c.vc.pos[y] > 1 - fc ? c.vc.pos[y] : 1 - fc;
        const int16_t *lf = c.vl.coef + (size_t)y * fl, *cf = c.vc.coef + (size_t)y * fc;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int i = 0; i < (dw + 1) >> 1; i++) {
            int Y1, Y2, Uv, Vv, j;
#define LUM(j, x) L[(size_t)rowsel(firstL, j, sh) * lp + (x)]
#define CHU(j) U[(size_t)rowsel(firstC, j, c.chrSrcH) * cp + i]
#define CHV(j) V[(size_t)rowsel(firstC, j, c.chrSrcH) * cp + i]
            if (fl == 1 && fc <= 2) {
                int ua = fc == 1 ? 0 : cf[1];
                Y1 = LUM(0, 2 * i) >> 7; Y2 = LUM(0, 2 * i + 1) >> 7;
                if (ua < 2048) { Uv = CHU(0) >> 7; Vv = CHV(0) >> 7; }
                else { Uv = (CHU(0) + CHU(1)) >> 8; Vv = (CHV(0) + CHV(1)) >> 8; }
                Y1 = u8clip(Y1); Y2 = u8clip(Y2); Uv = u8clip(Uv); Vv = u8clip(Vv);
            } else if (fl == 2 && fc == 2) {
                int ya = lf[1], ua = cf[1];
                Y1 = (LUM(0, 2 * i) * (4096 - ya) + LUM(1, 2 * i) * ya) >> 19;
                Y2 = (LUM(0, 2 * i + 1) * (4096 - ya) + LUM(1, 2 * i + 1) * ya) >> 19;
                Uv = (CHU(0) * (4096 - ua) + CHU(1) * ua) >> 19;
                Vv = (CHV(0) * (4096 - ua) + CHV(1) * ua) >> 19;
                Y1 = u8clip(Y1); Y2 = u8clip(Y2); Uv = u8clip(Uv); Vv = u8clip(Vv);
            } else {
                Y1 = Y2 = Uv = Vv = 1 << 18;
                for (j = 0; j < fl; j++) { Y1 += LUM(j, 2 * i) * lf[j]; Y2 += LUM(j, 2 * i + 1) * lf[j]; }
                for (j = 0; j < fc; j++) { Uv += CHU(j) * cf[j]; Vv += CHV(j) * cf[j]; }
                Y1 >>= 19; Y2 >>= 19; Uv >>= 19; Vv >>= 19;
                if ((Y1 | Y2 | Uv | Vv) & 0x100) { Y1 = u8clip(Y1); Y2 = u8clip(Y2); Uv = u8clip(Uv); Vv = u8clip(Vv); }
            }
            if (g_pk422) {      /* output_pixels (output.c:448-467): the four values are stored, not converted */
                const int room = 2 * i + 1 < dw || dstride >= 2 * (dw + 1);
                if (g_pk422 == 1) { d[4 * i] = (uint8_t)Y1; d[4 * i + 1] = (uint8_t)Uv; if (room) { d[4 * i + 2] = (uint8_t)Y2; d[4 * i + 3] = (uint8_t)Vv; } }
                else              { d[4 * i] = (uint8_t)Uv; d[4 * i + 1] = (uint8_t)Y1; if (room) { d[4 * i + 2] = (uint8_t)Vv; d[4 * i + 3] = (uint8_t)Y2; } }
                continue;
            }
            const uint8_t *r = ytab + rv[Vv], *g = ytab + gu[Uv] + gv[Vv], *b = ytab + bu[Uv];
            if (g_rgb16) {
                /* yuv2rgb_write (output.c:869-902): the channel tables of these depths hold the 8-bit value cut down to its field
                 * (yuv2rgb.c:806-844) and are indexed with an ordered dither added: 2 x 2 for 565 / 555, 4 x 4 for 444 */
                static const uint8_t d2x2_4[2][2] = { { 1, 3 }, { 2, 0 } }, d2x2_8[2][2] = { { 6, 2 }, { 0, 4 } };
                static const uint8_t d4x4[4][2] = { { 8, 4 }, { 2, 14 }, { 10, 6 }, { 0, 12 } };
                const int kind = (g_rgb16 & 7) - 1, fmt = kind >> 1, isbgr = kind & 1;
                int dr[2], dg[2], db[2];
                for (int k = 0; k < 2; k++) {
                    if (fmt == 0)      { dr[k] = d2x2_8[y & 1][k]; dg[k] = d2x2_4[y & 1][k];     db[k] = d2x2_8[(y & 1) ^ 1][k]; }
                    else if (fmt == 1) { dr[k] = d2x2_8[y & 1][k]; dg[k] = d2x2_8[y & 1][k ^ 1]; db[k] = d2x2_8[(y & 1) ^ 1][k]; }
                    else               { dr[k] = d4x4[y & 3][k];   dg[k] = d4x4[y & 3][k ^ 1];   db[k] = d4x4[(y & 3) ^ 3][k]; }
                }
                const int rs = fmt == 2 ? 4 : 3, gs = fmt == 0 ? 2 : fmt == 1 ? 3 : 4, gpos = fmt == 2 ? 4 : 5, hi = fmt == 0 ? 11 : fmt == 1 ? 10 : 8;
                const int Yk[2] = { Y1, Y2 };
                for (int k = 0; k < 2; k++) {
                    if (k && !(2 * i + 1 < dw || dstride >= 2 * (dw + 1))) break;
                    unsigned v = (unsigned)(r[Yk[k] + dr[k]] >> rs) << (isbgr ? 0 : hi) | (unsigned)(g[Yk[k] + dg[k]] >> gs) << gpos | (unsigned)(b[Yk[k] + db[k]] >> rs) << (isbgr ? hi : 0);
                    if (g_rgb16 & 8) v = ((v >> 8) | (v << 8)) & 0xffff;
                    d[4 * i + 2 * k] = (uint8_t)v; d[4 * i + 2 * k + 1] = (uint8_t)(v >> 8);
                }
                continue;
            }
            d[6 * i + 0] = r[Y1]; d[6 * i + 1] = g[Y1]; d[6 * i + 2] = b[Y1];
            if (2 * i + 1 < dw || dstride >= 3 * (dw + 1)) { d[6 * i + 3] = r[Y2]; d[6 * i + 4] = g[Y2]; d[6 * i + 5] = b[Y2]; }
        }
    }
    free(L); free(U); free(V);
    sws_close(&c);
    return dh;
}

/* yuv2plane1_8_c / yuv2planeX_8_c (output.c:242-265, dither 64) and yuv2plane1_10_c / yuv2planeX_10_c (:183-213) in one recipe:
 * the rounding constants and shifts only depend on the output depth */
static void vplane(const int16_t *s, int pitch, int sh, const bank_t *b, uint8_t *dst, int dstride, int w, int h, int doff)
{
    const int bits = g_dbits, top = (1 << bits) - 1;
    for (int y = 0; y < h; y++) {
        int fs = b->taps, first = b->pos[y] > 1 - fs ? b->pos[y] : 1 - fs;
        for (int i = 0; i < w; i++) {
            int v;
            const int dth = (bits == 8 && g_sbits > 8) ? dither128(y & 7, (i + doff) & 7) : 1 << (14 - bits);      /* 64 for 8-bit output without dithering */
            if (fs == 1) v = (s[(size_t)rowsel(first, 0, sh) * pitch + i] + dth) >> (15 - bits);
            else {
                v = bits == 8 ? dth << 12 : 1 << (26 - bits);
                for (int j = 0; j < fs; j++) v += s[(size_t)rowsel(first, j, sh) * pitch + i] * b->coef[(size_t)y * fs + j];
                v >>= 27 - bits;
            }
            v = v < 0 ? 0 : v > top ? top : v;
            if (bits == 8) dst[(size_t)y * dstride + i] = (uint8_t)v;
            else put16(dst + (size_t)y * dstride + 2 * i, v);
        }
    }
}

/* yuv2plane1_16_c / yuv2planeX_16_c (output.c:136-172), little-endian */
static void vplane16(const int32_t *s, int pitch, int sh, const bank_t *b, uint8_t *dst, int dstride, int w, int h)
{
    for (int y = 0; y < h; y++) {
        int fs = b->taps, first = b->pos[y] > 1 - fs ? b->pos[y] : 1 - fs;
        for (int i = 0; i < w; i++) {
            int v;
            if (fs == 1) {
                v = (s[(size_t)rowsel(first, 0, sh) * pitch + i] + 4) >> 3;
                v = v < 0 ? 0 : v > 65535 ? 65535 : v;
            } else {
                uint32_t acc = (1u << 14) - 0x40000000u;      /* the reference's bias keeps the sum inside 32 bits */
                for (int j = 0; j < fs; j++) acc += (uint32_t)(s[(size_t)rowsel(first, j, sh) * pitch + i] * b->coef[(size_t)y * fs + j]);
                v = (int32_t)acc >> 15;
                v = (v < -32768 ? -32768 : v > 32767 ? 32767 : v) + 0x8000;
            }
            put16(dst + (size_t)y * dstride + 2 * i, v);
        }
    }
}

int orc_sws_yuv420p_to_yuv420p(const uint8_t *const src[3], const int ss[3], int sw, int sh, uint8_t *const dst[3],
                               const int ds[3], int dw, int dh, int flags)
{
    sws_t c;
    if (sws_open(&c, sw, sh, dw, dh, 0, flags)) return -1;
    if (sw == dw && sh == dh && g_hs == g_dhs && g_vs == g_dvs && !g_nocopy && !g_range && !uses_filter() && g_sbits == 8 && !g_src_gray) {   /* unscaled, same sub-sampling and range: planarCopyWrapper (utils.c:1043-1054,
                                         swscale_unscaled.c:793-1020); 8 -> 9 / 10 bits is a plain shift for limited-range sources (:946-971) */
        for (int p = 0; p < 3; p++) {
            int w = p ? c.chrSrcW : sw, h = p ? c.chrSrcH : sh;
            for (int y = 0; y < h; y++) {
                if (g_dbits == 8) { memcpy(dst[p] + (size_t)y * ds[p], src[p] + (size_t)y * ss[p], w); continue; }
                for (int x = 0; x < w; x++) {
                    const int v = g_dbits == 16 ? src[p][(size_t)y * ss[p] + x] * 257 : src[p][(size_t)y * ss[p] + x] << (g_dbits - 8);
                    put16(dst[p] + (size_t)y * ds[p] + 2 * x, v);
                }
            }
        }
        sws_close(&c);
        return dh;
    }
    if (g_src_gray && sw == dw && sh == dh && !g_range && !uses_filter()) {
        /* isPlanarYUV(dst) && isGray(src): planarCopyWrapper whatever the destination's sub-sampling (swscale_unscaled.c:1156) -- the luma plane is
         * copied, the planes the source does not have are filled with 128 (:812-823).  8-bit destinations (fill_plane9or10 / the 16-bit fill: not restated) */
        if (g_dbits != 8) { sws_close(&c); return -1; }
        for (int y = 0; y < sh; y++) memcpy(dst[0] + (size_t)y * ds[0], src[0] + (size_t)y * ss[0], sw);
        for (int p = 1; p < 3; p++)
            for (int y = 0; y < c.chrDstH; y++) memset(dst[p] + (size_t)y * ds[p], 128, c.chrDstW);
        sws_close(&c);
        return dh;
    }
    int lp, cp;
    if (g_dbits == 16) {                /* 19-bit lines; no fast-bilinear line functions at this depth (swscale.c:728-741) */
        if (g_range) { sws_close(&c); return -1; }      /* (the *Range*16_c variants are not restated) */
        int32_t *L = hpass19(src[0], ss[0], sh, &c.hl, &lp), *U = hpass19(src[1], ss[1], c.chrSrcH, &c.hc, &cp), *V = hpass19(src[2], ss[2], c.chrSrcH, &c.hc, &cp);
        gray_lines19(U, V, (size_t)cp * c.chrSrcH);
        vplane16(L, lp, sh, &c.vl, dst[0], ds[0], dw, dh);
        vplane16(U, cp, c.chrSrcH, &c.vc, dst[1], ds[1], c.chrDstW, c.chrDstH);
        vplane16(V, cp, c.chrSrcH, &c.vc, dst[2], ds[2], c.chrDstW, c.chrDstH);
        free(L); free(U); free(V);
        sws_close(&c);
        return dh;
    }
    const int fast = (c.flags & F_FAST_BILINEAR) && g_sbits == 8;      /* the fast line functions only exist for 8-bit sources (swscale.c:733-739) */
    int16_t *L = fast ? hfast(src[0], ss[0], sh, sw, dw, c.lumXInc, 0, &lp) : hpass(src[0], ss[0], sh, &c.hl, &lp);
    int16_t *U = fast ? hfast(src[1], ss[1], c.chrSrcH, c.chrSrcW, c.chrDstW, c.chrXInc, 1, &cp) : hpass(src[1], ss[1], c.chrSrcH, &c.hc, &cp);
    int16_t *V = fast ? hfast(src[2], ss[2], c.chrSrcH, c.chrSrcW, c.chrDstW, c.chrXInc, 1, &cp) : hpass(src[2], ss[2], c.chrSrcH, &c.hc, &cp);
    gray_lines(U, V, (size_t)cp * c.chrSrcH);
    range_lines(L, lp, sh, dw, 0); range_lines(U, cp, c.chrSrcH, c.chrDstW, 1); range_lines(V, cp, c.chrSrcH, c.chrDstW, 1);
    vplane(L, lp, sh, &c.vl, dst[0], ds[0], dw, dh, 0);
    vplane(U, cp, c.chrSrcH, &c.vc, dst[1], ds[1], c.chrDstW, c.chrDstH, 0);
    vplane(V, cp, c.chrSrcH, &c.vc, dst[2], ds[2], c.chrDstW, c.chrDstH, 3);      /* chrDither8 with offset 3 for V (swscale.c:636-644) */
    free(L); free(U); free(V);
    sws_close(&c);
    return dh;
}


/* Semi-planar sources: nvXXtoUV_c (libswscale/input.c:475-497) in front of the planar path.  A same-size planar
 * destination is the reference's nv12ToPlanarWrapper (swscale_unscaled.c:160-181): luma copy and a split of
 * srcW/2 x srcH/2 chroma samples; an rgb destination always runs swscale() for these sources (the unscaled table
 * converter is only installed for planar yuv, swscale_unscaled.c:1051-1055). */
static int to_rgb_or_bgr(const uint8_t *const src[3], const int ss[3], int sw, int sh, int dst_fmt, uint8_t *dst, int dstride, int dw, int dh, int flags);
int orc_sws_nv12(int nv21, const uint8_t *y, int ystride, const uint8_t *uv, int uvstride, int sw, int sh, int dst_fmt,
                 uint8_t *const dst[3], const int dstride[3], int dw, int dh, int flags)
{
    if (dst_fmt == 0 && sw == dw && sh == dh && !g_range) {
        for (int r = 0; r < sh; r++) memcpy(dst[0] + (size_t)r * dstride[0], y + (size_t)r * ystride, sw);
        for (int r = 0; r < sh / 2; r++)
            for (int x = 0; x < sw / 2; x++) {
                dst[nv21 ? 2 : 1][(size_t)r * dstride[nv21 ? 2 : 1] + x] = uv[(size_t)r * uvstride + 2 * x];
                dst[nv21 ? 1 : 2][(size_t)r * dstride[nv21 ? 1 : 2] + x] = uv[(size_t)r * uvstride + 2 * x + 1];
            }
        return sh;
    }
    const int cw = (sw + 1) >> 1, ch = (sh + 1) >> 1, pitch = cw + 16;
    uint8_t *u = malloc((size_t)pitch * ch * 2), *v = u + (size_t)pitch * ch;
    if (!u) return -1;
    memset(u, 0, (size_t)pitch * ch * 2);
    for (int r = 0; r < ch; r++)
        for (int x = 0; x < cw; x++) {
            (nv21 ? v : u)[(size_t)r * pitch + x] = uv[(size_t)r * uvstride + 2 * x];
            (nv21 ? u : v)[(size_t)r * pitch + x] = uv[(size_t)r * uvstride + 2 * x + 1];
        }
    const uint8_t *src[3] = { y, u, v };
    const int ss[3] = { ystride, pitch, pitch };
    g_nocopy = 1;
    int r = dst_fmt == 2 || dst_fmt == 3 || (dst_fmt >= 25 && dst_fmt <= 28) || dst_fmt == 1 || dst_fmt == 15 || IS_RGB16(dst_fmt) ? to_rgb_or_bgr(src, ss, sw, sh, dst_fmt, dst[0], dstride[0], dw, dh, flags | 0x40000)
                                         : orc_sws_yuv420p_to_yuv420p(src, ss, sw, sh, dst, dstride, dw, dh, flags);
    g_nocopy = 0;
    free(u);
    return r;
}


/* rgb24 output helper for the bgr24 destination: the two only differ in byte order (output.c:1008-1040) */
static int to_rgb_or_bgr(const uint8_t *const src[3], const int ss[3], int sw, int sh, int dst_fmt, uint8_t *dst, int dstride, int dw, int dh, int flags)
{
    if (dst_fmt == 2) return orc_sws_yuv420p_to_rgb24(src, ss, sw, sh, dst, dstride, dw, dh, flags);
    {   /* rgb565 / bgr565 / rgb555 / bgr555 / rgb444 / bgr444, LE and BE (libavutil/pixfmt.h 36-43, 54-57) */
        int k16 = 0;
        switch (dst_fmt) {
        case 37: k16 = 1; break; case 36: k16 = 1 | 8; break; case 41: k16 = 2; break; case 40: k16 = 2 | 8; break;
        case 39: k16 = 3; break; case 38: k16 = 3 | 8; break; case 43: k16 = 4; break; case 42: k16 = 4 | 8; break;
        case 54: k16 = 5; break; case 55: k16 = 5 | 8; break; case 56: k16 = 6; break; case 57: k16 = 6 | 8; break;
        }
        if (k16) {
            if (uses_filter()) return -1;
            g_rgb16 = k16;
            int r16 = orc_sws_yuv420p_to_rgb24(src, ss, sw, sh, dst, dstride, dw, dh, flags & ~F_FULL_CHR_H_INT);
            g_rgb16 = 0;
            return r16;
        }
    }
    if (IS_RGB48(dst_fmt)) {
        if (uses_filter()) return -1;
        g_rgb48 = (dst_fmt == 34 || dst_fmt == 35 ? 1 : 2) | (dst_fmt == 34 || dst_fmt == 59 ? 8 : 0);
        int r48 = orc_sws_yuv420p_to_rgb24(src, ss, sw, sh, dst, dstride, dw, dh, flags & ~F_FULL_CHR_H_INT);
        g_rgb48 = 0;
        return r48;
    }
    if (dst_fmt == 1 || dst_fmt == 15) {          /* yuyv422 / uyvy422: the packed output stage without the colour conversion */
        g_pk422 = dst_fmt == 1 ? 1 : 2;
        int r422 = orc_sws_yuv420p_to_rgb24(src, ss, sw, sh, dst, dstride, dw, dh, flags & ~F_FULL_CHR_H_INT);
        g_pk422 = 0;
        return r422;
    }
    const int pitch = (dw + 1) * 3;
    uint8_t *t = malloc((size_t)pitch * dh);
    if (!t) return -1;
    memset(t, 0, (size_t)pitch * dh);
    int r = orc_sws_yuv420p_to_rgb24(src, ss, sw, sh, t, pitch, dw, dh, flags);
    if (dst_fmt == 3) {
        for (int y = 0; r == dh && y < dh; y++)
            for (int x = 0; x < dw; x++)
                for (int k = 0; k < 3; k++) dst[(size_t)y * dstride + 3 * x + k] = t[(size_t)y * pitch + 3 * x + 2 - k];
    } else {
        /* argb 25, rgba 26, abgr 27, bgra 28: the 32-bit colour tables hold the 24-bit channel values plus alpha 255
         * (yuv2rgb.c:763-800).  Like the 24-bit functions: whole pixel pairs (one pixel past an odd width), single pixels with
         * SWS_FULL_CHR_H_INT, and the unscaled table converter (same size, no SWS_ACCURATE_RND, even height, 4:2:0 / 4:2:2
         * planar source) leaves an odd last column alone */
        static const int order[4][4] = { { 1, 2, 3, 0 }, { 0, 1, 2, 3 }, { 3, 2, 1, 0 }, { 2, 1, 0, 3 } };
        const int *o = order[dst_fmt - 25];
        int w = dw;
        if (sw == dw && sh == dh && !(flags & F_ACCURATE_RND) && !(dh & 1) && g_hs == 1 && g_vs <= 1 && !uses_filter() && g_sbits == 8) w &= ~1;
        else if ((w & 1) && !(flags & F_FULL_CHR_H_INT) && dstride >= 4 * (w + 1)) w++;
        for (int y = 0; r == dh && y < dh; y++)
            for (int x = 0; x < w; x++) {
                uint8_t *d = dst + (size_t)y * dstride + 4 * x;
                const uint8_t *p = t + (size_t)y * pitch + 3 * x;
                d[o[0]] = p[0]; d[o[1]] = p[1]; d[o[2]] = p[2]; d[o[3]] = 255;
            }
    }
    free(t);
    return r;
}

/* input.c:38-47 */
#define RSH 15
#define C_BY  ((int)(0.114 * 219 / 255 * (1 << RSH) + 0.5))
#define C_BV (-(int)(0.081 * 224 / 255 * (1 << RSH) + 0.5))
#define C_BU  ((int)(0.500 * 224 / 255 * (1 << RSH) + 0.5))
#define C_GY  ((int)(0.587 * 219 / 255 * (1 << RSH) + 0.5))
#define C_GV (-(int)(0.419 * 224 / 255 * (1 << RSH) + 0.5))
#define C_GU (-(int)(0.331 * 224 / 255 * (1 << RSH) + 0.5))
#define C_RY  ((int)(0.299 * 219 / 255 * (1 << RSH) + 0.5))
#define C_RV  ((int)(0.500 * 224 / 255 * (1 << RSH) + 0.5))
#define C_RU (-(int)(0.169 * 224 / 255 * (1 << RSH) + 0.5))

/* Packed sources.  src_fmt: AV_PIX_FMT_YUYV422 1, RGB24 2, BGR24 3, UYVY422 15.
 *  - same size: the reference installs special converters (swscale_unscaled.c:1063-1072,1140-1145,1152-1176): rgb24 <-> bgr24
 *    (rgb24tobgr24) and the same-format copy; bgr24 -> yuv420p without SWS_ACCURATE_RND = rgb24toyv12_c (rgb2rgb_template.c:638-693,
 *    8-bit coefficients rgb2rgb.c:111-120); yuyv422 / uyvy422 -> yuv420p = yuyvtoyuv420_c / uyvytoyuv420_c (:854-910);
 *  - otherwise swscale() with the input readers in front (input.c:369-387,456-473,539-625): 8-bit planes, chroma at half width for
 *    rgb unless SWS_FULL_CHR_H_INP or an up-scale asks for every pixel (utils.c:1021-1034), one chroma line per source line. */
static int packed_source(int src_fmt, const uint8_t *src, int stride, int sw, int sh, int dst_fmt, uint8_t *const dst[3],
                         const int ds[3], int dw, int dh, int flags)
{
    /* byte positions of r, g, b: rgb24, bgr24, and argb 25 / rgba 26 / abgr 27 / bgra 28 (rgb16_32ToY / ToUV / ToUV_half templates,
     * input.c:230-330: the same sums on the 8-bit channels, both sides of the shift scaled by 2^8) */
    static const int pos32[4][3] = { { 1, 2, 3 }, { 0, 1, 2 }, { 3, 2, 1 }, { 2, 1, 0 } };
    const int src32 = src_fmt >= 25 && src_fmt <= 28;
    const int rgb_src = src_fmt == 2 || src_fmt == 3 || src32;
    const int ro = src32 ? pos32[src_fmt - 25][0] : src_fmt == 3 ? 2 : 0, go = src32 ? pos32[src_fmt - 25][1] : 1, bo = src32 ? pos32[src_fmt - 25][2] : 2 - ro;
    const int rgb_bpp = src32 ? 4 : 3;
    const int real_rgb_dst = dst_fmt == 2 || dst_fmt == 3 || (dst_fmt >= 25 && dst_fmt <= 28);
    const int rgb_dst = real_rgb_dst || dst_fmt == 1 || dst_fmt == 15;       /* packed destinations share the output stage */
    const int dst32 = dst_fmt >= 25 && dst_fmt <= 28;
    if (sw == dw && sh == dh && rgb_src && real_rgb_dst && (src32 || dst32) && !g_nospecial) {
        /* rgbToRgbWrapper (swscale_unscaled.c:590-705) with the byte converters of rgb2rgb.c:139-175,335-352 / rgb2rgb_template.c:31-78,338-350:
         * every pair it serves is a channel remap (alpha copied 32 -> 32, 255 for 24 -> 32, dropped 32 -> 24).  A 24-bit source to argb / abgr
         * runs the 4-byte writer one byte into the row (ALT32_CORR, :691-692): the first alpha byte is never written and the last write lands
         * past the row -- no defined result, refused.  Only the sw x sh pixels are written here (the reference's one-call variant also converts the
         * row padding when the pitches are proportional, :694-697). */
        static const int chan[6][4] = { { 0, 1, 2, -1 }, { 2, 1, 0, -1 }, { 1, 2, 3, 0 }, { 0, 1, 2, 3 }, { 3, 2, 1, 0 }, { 2, 1, 0, 3 } };  /* byte of r, g, b, a: rgb24 bgr24 argb rgba abgr bgra */
        const int *sc = chan[src32 ? src_fmt - 23 : src_fmt - 2], *dc = chan[dst32 ? dst_fmt - 23 : dst_fmt - 2];
        const int dbpp = dst32 ? 4 : 3;
        if (!src32 && (dst_fmt == 25 || dst_fmt == 27)) return -1;
        for (int y = 0; y < sh; y++)
            for (int x = 0; x < sw; x++) {
                const uint8_t *q = src + (size_t)y * stride + (size_t)rgb_bpp * x;
                uint8_t *d = dst[0] + (size_t)y * ds[0] + (size_t)dbpp * x;
                for (int k = 0; k < 3; k++) d[dc[k]] = q[sc[k]];
                if (dst32) d[dc[3]] = src32 ? q[sc[3]] : 255;
            }
        return sh;
    }
    /* 32 -> 32 bit at another size scales the alpha plane as well: not restated */
    if (src32 && dst32) return -1;
    if (sw == dw && sh == dh && !g_nospecial && (!g_range || real_rgb_dst)) {
        if (rgb_src && !src32 && real_rgb_dst) {
            for (int y = 0; y < sh; y++)
                for (int x = 0; x < sw; x++)
                    for (int k = 0; k < 3; k++)
                        dst[0][(size_t)y * ds[0] + 3 * x + k] = src[(size_t)y * stride + 3 * x + (src_fmt == dst_fmt ? k : 2 - k)];
            return sh;
        }
        if (src_fmt == 3 && dst_fmt == 0 && !(flags & F_ACCURATE_RND)) {
            if (sh & 1) return -1;                           /* the reference's loop converts rows in pairs */
            for (int y = 0; y < sh; y++)
                for (int i = 0; i < (sw >> 1); i++)
                    for (int k = 0; k < 2; k++) {
                        const uint8_t *p = src + (size_t)y * stride + 6 * i + 3 * k;
                        const int b = p[0], g = p[1], r = p[2];
                        dst[0][(size_t)y * ds[0] + 2 * i + k] = (uint8_t)(((66 * r + 129 * g + 25 * b) >> 8) + 16);
                        if (!(y & 1) && !k) {
                            dst[1][(size_t)(y >> 1) * ds[1] + i] = (uint8_t)(((-37 * r - 73 * g + 112 * b) >> 8) + 128);
                            dst[2][(size_t)(y >> 1) * ds[2] + i] = (uint8_t)(((112 * r - 93 * g - 17 * b) >> 8) + 128);
                        }
                    }
            return sh;
        }
        if (!rgb_src && dst_fmt == 4) {                     /* yuyvtoyuv422_c / uyvytoyuv422_c (rgb2rgb_template.c:873-888,912-927) */
            const int yo = src_fmt == 15, co = 1 - yo, cw = (sw + 1) >> 1;
            for (int y = 0; y < sh; y++) {
                const uint8_t *s = src + (size_t)y * stride;
                for (int x = 0; x < sw; x++) dst[0][(size_t)y * ds[0] + x] = s[2 * x + yo];
                for (int i = 0; i < cw; i++) {
                    const int ok = 2 * i + 1 < sw || 4 * i + 4 <= stride || y < sh - 1;
                    dst[1][(size_t)y * ds[1] + i] = s[4 * i + co];
                    dst[2][(size_t)y * ds[2] + i] = ok ? s[4 * i + 2 + co] : s[4 * i + co];
                }
            }
            return sh;
        }
        if (!rgb_src && dst_fmt == 0) {
            const int yo = src_fmt == 15, co = 1 - yo, cw = (sw + 1) >> 1;
            for (int y = 0; y < sh; y++) {
                const uint8_t *s = src + (size_t)y * stride, *p = s - stride;
                for (int x = 0; x < sw; x++) dst[0][(size_t)y * ds[0] + x] = s[2 * x + yo];
                if (y & 1)
                    for (int i = 0; i < cw; i++) {
                        const int ok = 2 * i + 1 < sw || 4 * i + 4 <= stride || y < sh - 1;
                        dst[1][(size_t)(y >> 1) * ds[1] + i] = (uint8_t)((p[4 * i + co] + s[4 * i + co]) >> 1);
                        dst[2][(size_t)(y >> 1) * ds[2] + i] = (uint8_t)((p[4 * i + 2 + co] + (ok ? s[4 * i + 2 + co] : s[4 * i + co])) >> 1);
                    }
            }
            return sh;
        }
    }
    int hs = 1;
    if (rgb_src) {
        const int chr_dst_hsub = !rgb_dst ? g_dhs : ((flags & F_FULL_CHR_H_INT) && real_rgb_dst) ? 0 : 1;
        hs = (!(flags & 0x4000) && ((dw >> chr_dst_hsub) <= (sw >> 1) || (flags & 1))) ? 1 : 0;
    }
    const int cw = -((-sw) >> hs), yp = sw + 16, cp = cw + 16;
    uint8_t *Y = calloc((size_t)yp * sh + 2 * (size_t)cp * sh, 1), *U = Y + (size_t)yp * sh, *V = U + (size_t)cp * sh;
    if (!Y) return -1;
    for (int y = 0; y < sh; y++) {
        const uint8_t *row = src + (size_t)y * stride;
        for (int i = 0; 2 * i < sw; i++) {
            const int second = 2 * i + 1 < sw;
            const int bpp = rgb_src ? rgb_bpp : 2;
            const int ok = second || y < sh - 1 || (2 * i + 2) * bpp <= stride;   /* the sample past an odd width is inside the frame */
            const uint8_t *s = row + 2 * i * bpp;
            if (rgb_src) {
                const int r0 = s[ro], g0 = s[go], b0 = s[bo];
                const int r1 = ok ? s[rgb_bpp + ro] : r0, g1 = ok ? s[rgb_bpp + go] : g0, b1 = ok ? s[rgb_bpp + bo] : b0;
                Y[(size_t)y * yp + 2 * i] = (uint8_t)((C_RY * r0 + C_GY * g0 + C_BY * b0 + (33 << (RSH - 1))) >> RSH);
                if (second) Y[(size_t)y * yp + 2 * i + 1] = (uint8_t)((C_RY * r1 + C_GY * g1 + C_BY * b1 + (33 << (RSH - 1))) >> RSH);
                if (hs) {
                    const int r = r0 + r1, g = g0 + g1, b = b0 + b1;
                    U[(size_t)y * cp + i] = (uint8_t)((C_RU * r + C_GU * g + C_BU * b + (257 << RSH)) >> (RSH + 1));
                    V[(size_t)y * cp + i] = (uint8_t)((C_RV * r + C_GV * g + C_BV * b + (257 << RSH)) >> (RSH + 1));
                } else {
                    U[(size_t)y * cp + 2 * i] = (uint8_t)((C_RU * r0 + C_GU * g0 + C_BU * b0 + (257 << (RSH - 1))) >> RSH);
                    V[(size_t)y * cp + 2 * i] = (uint8_t)((C_RV * r0 + C_GV * g0 + C_BV * b0 + (257 << (RSH - 1))) >> RSH);
                    if (second) {
                        U[(size_t)y * cp + 2 * i + 1] = (uint8_t)((C_RU * r1 + C_GU * g1 + C_BU * b1 + (257 << (RSH - 1))) >> RSH);
                        V[(size_t)y * cp + 2 * i + 1] = (uint8_t)((C_RV * r1 + C_GV * g1 + C_BV * b1 + (257 << (RSH - 1))) >> RSH);
                    }
                }
            } else {
                const int yo = src_fmt == 15, co = 1 - yo;
                Y[(size_t)y * yp + 2 * i] = s[yo];
                if (second) Y[(size_t)y * yp + 2 * i + 1] = s[2 + yo];
                U[(size_t)y * cp + i] = s[co];
                V[(size_t)y * cp + i] = ok ? s[2 + co] : s[co];
            }
        }
    }
    const uint8_t *pl[3] = { Y, U, V };
    const int ss[3] = { yp, cp, cp };
    g_hs = hs; g_vs = 0; g_nocopy = 1;
    /* (the unscaled table converter and planarCopyWrapper are only installed for planar yuv sources: keep the port off those branches) */
    int r = rgb_dst ? to_rgb_or_bgr(pl, ss, sw, sh, dst_fmt, dst[0], ds[0], dw, dh, flags | F_ACCURATE_RND)
                    : orc_sws_yuv420p_to_yuv420p(pl, ss, sw, sh, dst, ds, dw, dh, flags);
    g_hs = 1; g_vs = 1; g_nocopy = 0;
    free(Y);
    return r;
}

/* pal8 (AV_PIX_FMT_PAL8 = 11) sources: src[0] = one index per pixel, src[1] = 256 native-endian 0xAARRGGBB entries.  sws_scale() turns the palette into
 * limited-range y / u / v per entry for every call (swscale_unscaled.c:1236-1268, the rgb -> yuv constants of swscale_internal.h with RGB2YUV_SHIFT 15)
 * and the readers palToY_c / palToUV_c look the samples up (input.c:321-343): chroma at full resolution (pixdesc: no sub-sampling).  At the same size a
 * 24 / 32-bit rgb destination is palToRgbWrapper, a lookup of r, g, b with alpha 255 (:342-384,1270-1295); the palette's alpha byte is never used. */
static int pal8_source(const uint8_t *src, int stride, const uint8_t *pal, int sw, int sh, int dst_fmt, uint8_t *const dst[3], const int ds[3], int dw, int dh, int flags)
{
    int R[256], G[256], B[256];
    for (int i = 0; i < 256; i++) {
        uint32_t p; memcpy(&p, pal + 4 * i, 4);
        R[i] = (p >> 16) & 255; G[i] = (p >> 8) & 255; B[i] = p & 255;
    }
    const int real_rgb_dst = dst_fmt == 2 || dst_fmt == 3 || (dst_fmt >= 25 && dst_fmt <= 28);
    const int rgb_dst = real_rgb_dst || dst_fmt == 1 || dst_fmt == 15 || IS_RGB16(dst_fmt) || IS_RGB48(dst_fmt);
    if (IS_RGB16(dst_fmt) || IS_RGB48(dst_fmt)) return -1;            /* (like the packed rgb sources: not restated) */
    if (sw == dw && sh == dh && real_rgb_dst && !uses_filter()) {
        static const int chan[6][4] = { { 0, 1, 2, -1 }, { 2, 1, 0, -1 }, { 1, 2, 3, 0 }, { 0, 1, 2, 3 }, { 3, 2, 1, 0 }, { 2, 1, 0, 3 } };  /* byte of r, g, b, a */
        const int *dc = chan[dst_fmt >= 25 ? dst_fmt - 23 : dst_fmt - 2], bpp = dst_fmt >= 25 ? 4 : 3;
        for (int y = 0; y < sh; y++)
            for (int x = 0; x < sw; x++) {
                const int i = src[(size_t)y * stride + x];
                uint8_t *d = dst[0] + (size_t)y * ds[0] + (size_t)bpp * x;
                d[dc[0]] = (uint8_t)R[i]; d[dc[1]] = (uint8_t)G[i]; d[dc[2]] = (uint8_t)B[i];
                if (bpp == 4) d[dc[3]] = 255;
            }
        return sh;
    }
    if (uses_filter()) return -1;
    const int yp = sw + 16;
    uint8_t *Y = calloc((size_t)yp * sh * 3, 1), *U = Y + (size_t)yp * sh, *V = U + (size_t)yp * sh;
    if (!Y) return -1;
    for (int y = 0; y < sh; y++)
        for (int x = 0; x < sw; x++) {
            const int i = src[(size_t)y * stride + x], r = R[i], g = G[i], b = B[i];
            Y[(size_t)y * yp + x] = u8clip((C_RY * r + C_GY * g + C_BY * b + (33 << (RSH - 1))) >> RSH);
            U[(size_t)y * yp + x] = u8clip((C_RU * r + C_GU * g + C_BU * b + (257 << (RSH - 1))) >> RSH);
            V[(size_t)y * yp + x] = u8clip((C_RV * r + C_GV * g + C_BV * b + (257 << (RSH - 1))) >> RSH);
        }
    const uint8_t *pl[3] = { Y, U, V };
    const int ss[3] = { yp, yp, yp };
    g_hs = 0; g_vs = 0; g_nocopy = 1;
    /* (planarCopyWrapper is only installed for planar yuv / gray sources: keep the port off that branch) */
    int r = rgb_dst ? to_rgb_or_bgr(pl, ss, sw, sh, dst_fmt, dst[0], ds[0], dw, dh, flags)
                    : orc_sws_yuv420p_to_yuv420p(pl, ss, sw, sh, dst, ds, dw, dh, flags);
    g_hs = 1; g_vs = 1; g_nocopy = 0;
    free(Y);
    return r;
}

/* planar destinations: AV_PIX_FMT_YUV420P 0, YUV422P 4, YUV444P 5, YUV410P 6, YUV411P 7, YUV440P 31, and the little-endian
 * YUV420P9 62, YUV420P10 64, YUV422P10 66, YUV444P9 68, YUV444P10 70, YUV422P9 72 */
static int planar_dst(int fmt, int *hs, int *vs, int *bits)
{
    *bits = 8; g_dbe = 0;
    switch (fmt) {               /* big-endian twins: 9 / 10-bit LE - 1, 16-bit LE + 1 */
    case 61: case 63: case 65: case 67: case 69: case 71: g_dbe = 1; fmt += 1; break;
    case 48: case 50: case 52: g_dbe = 1; fmt -= 1; break;
    }
    switch (fmt) {
    case 0: *hs = 1; *vs = 1; return 1;   case 4: *hs = 1; *vs = 0; return 1;   case 5: *hs = 0; *vs = 0; return 1;
    case 6: *hs = 2; *vs = 2; return 1;   case 7: *hs = 2; *vs = 0; return 1;   case 31: *hs = 0; *vs = 1; return 1;
    case 62: case 64: *hs = 1; *vs = 1; *bits = fmt == 62 ? 9 : 10; return 1;
    case 72: case 66: *hs = 1; *vs = 0; *bits = fmt == 72 ? 9 : 10; return 1;
    case 68: case 70: *hs = 0; *vs = 0; *bits = fmt == 68 ? 9 : 10; return 1;
    case 47: *hs = 1; *vs = 1; *bits = 16; return 1;      /* YUV420P16LE, YUV422P16LE 49, YUV444P16LE 51 */
    case 49: *hs = 1; *vs = 0; *bits = 16; return 1;
    case 51: *hs = 0; *vs = 0; *bits = 16; return 1;
    }
    return 0;
}

/* Any source format the product takes over, to rgb24 (dst_fmt 2), bgr24 (3) or a planar yuv format (planar_dst above).  Planar 8-bit YUV sources of any chroma
 * sub-sampling (getSubSampleFactors, utils.c:983) run the same pipeline with chrSrcW / chrSrcH derived from the format.
 * src_fmt: AV_PIX_FMT_YUV420P 0, YUV422P 4, YUV444P 5, YUV410P 6, YUV411P 7, YUV440P 31; packed: YUYV422 1, RGB24 2, BGR24 3,
 * UYVY422 15 (src[0] / ss[0] only). */
static int sws_any(int src_fmt, const uint8_t *const src[3], const int ss[3], int sw, int sh, int dst_fmt,
                   uint8_t *const dst[3], const int dstride[3], int dw, int dh, int flags);

/* nv12 (23) / nv21 (24) destinations: yuv2nv12cX_c (output.c:267-303) is yuv2planeX_8_c on both chroma planes, written
 * interleaved (the dither is the constant 64 for 8-bit sources, swscale.c:395-399); yuv420p of the same size goes through
 * planarToNv12Wrapper (swscale_unscaled.c:138-156, srcW / 2 x srcH / 2 chroma samples). */
int orc_sws_planar(int src_fmt, const uint8_t *const src[3], const int ss[3], int sw, int sh, int dst_fmt,
                   uint8_t *const dst[3], const int dstride[3], int dw, int dh, int flags)
{
    if (dst_fmt != 23 && dst_fmt != 24) return sws_any(src_fmt, src, ss, sw, sh, dst_fmt, dst, dstride, dw, dh, flags);
    if (uses_filter()) return -1;
    const int swap = dst_fmt == 24;
    if (sw == dw && sh == dh) {
        if (src_fmt == 23 || src_fmt == 24) return -1;         /* the reference's plane copy skips the chroma plane there */
        if (src_fmt == 0) {            /* (a yuvj420p source is format 12 here: it takes the scaler below, with the range conversion) */
            for (int y = 0; y < sh; y++) memcpy(dst[0] + (size_t)y * dstride[0], src[0] + (size_t)y * ss[0], sw);
            for (int y = 0; y < sh / 2; y++)
                for (int x = 0; x < sw / 2; x++) {
                    dst[1][(size_t)y * dstride[1] + 2 * x + swap] = src[1][(size_t)y * ss[1] + x];
                    dst[1][(size_t)y * dstride[1] + 2 * x + 1 - swap] = src[2][(size_t)y * ss[2] + x];
                }
            return sh;
        }
    }
    const int cw = (dw + 1) >> 1, ch = (dh + 1) >> 1;
    uint8_t *u = malloc((size_t)cw * ch * 2), *v = u + (size_t)cw * ch;
    if (!u) return -1;
    uint8_t *const d3[3] = { dst[0], u, v };
    const int s3[3] = { dstride[0], cw, cw };
    /* (the inner call's destination "yuv420p" must not take the unscaled special converters the reference only installs for a real
     * yuv420p destination: yuyvtoyuv420, rgb24toyv12 ...) */
    g_nospecial = 1;
    int r = sws_any(src_fmt, src, ss, sw, sh, 0, d3, s3, dw, dh, flags);
    g_nospecial = 0;
    for (int y = 0; r == dh && y < ch; y++)
        for (int x = 0; x < cw; x++) {
            dst[1][(size_t)y * dstride[1] + 2 * x + swap] = u[(size_t)y * cw + x];
            dst[1][(size_t)y * dstride[1] + 2 * x + 1 - swap] = v[(size_t)y * cw + x];
        }
    free(u);
    return r;
}

static int sws_any(int src_fmt, const uint8_t *const src[3], const int ss[3], int sw, int sh, int dst_fmt,
                   uint8_t *const dst[3], const int dstride[3], int dw, int dh, int flags)
{
    int hs, vs, r;
    if (src_fmt == 8 && !g_src_gray) {
        /* gray8 source: one plane, chroma geometry of a format without sub-sampling (pixdesc: log2_chroma_w = log2_chroma_h = 0); the chroma lines are
         * never converted (see g_src_gray).  Not restated: range conversion to a full-range yuvj destination (only the luma function would ever run),
         * semi-planar destinations (planarCopyWrapper fills half of the interleaved chroma row), SwsFilter vectors. */
        const int dj = (dst_fmt >= 12 && dst_fmt <= 14) || dst_fmt == 32;
        if (dj || dst_fmt == 23 || dst_fmt == 24 || uses_filter() || g_nospecial) return -1;
        { int h2, v2, b2 = 8; if (planar_dst(dst_fmt, &h2, &v2, &b2) && b2 == 16) { g_dbe = 0; return -1; } g_dbe = 0; }      /* (19-bit lines from a gray source: not restated for planar destinations) */
        if (sw == dw && sh == dh && (dst_fmt == 2 || dst_fmt == 3 || (dst_fmt >= 25 && dst_fmt <= 28))) {
            /* gray8 is a pseudo-palette format: at the same size a 24 / 32-bit rgb destination gets palToRgbWrapper (swscale_unscaled.c:342-384,
             * installed at :1114-1121) with the palette sws_scale() builds for it, r = g = b = the sample (:1257-1259), alpha 255 (:1270-1295) */
            const int bpp = dst_fmt <= 3 ? 3 : 4, a = (dst_fmt == 25 || dst_fmt == 27) ? 0 : 3;
            for (int y = 0; y < sh; y++)
                for (int x = 0; x < sw; x++) {
                    uint8_t *d = dst[0] + (size_t)y * dstride[0] + (size_t)bpp * x;
                    const uint8_t v = src[0][(size_t)y * ss[0] + x];
                    if (bpp == 3) { d[0] = d[1] = d[2] = v; }
                    else { d[0] = d[1] = d[2] = d[3] = v; d[a] = 255; }
                }
            return sh;
        }
        const uint8_t *const s3[3] = { src[0], src[0], src[0] };
        const int ss3[3] = { ss[0], ss[0], ss[0] };
        g_src_gray = 1;
        r = sws_any(5, s3, ss3, sw, sh, dst_fmt, dst, dstride, dw, dh, flags);
        g_src_gray = 0;
        return r;
    }
    if (src_fmt == 33) {
        /* yuva420p: the alpha plane is only read when the destination has alpha too (c->alpPixBuf, utils.c:1244; needAlpha, yuv2rgb.c:870); for
         * every other destination the format is treated like yuv420p wherever the reference tests formats (swscale_unscaled.c:1041-1153) */
        if (dst_fmt >= 25 && dst_fmt <= 28) return -1;
        src_fmt = 0;
    }
    if (dst_fmt == 8) {
        /* gray8: luma only.  swscale() skips the chroma of a gray destination (swscale.c:618-630) and the same-size case is the plane copy for
         * every planar yuv source (isPlanarYUV(src) && isGray(dst), swscale_unscaled.c:1155): the luma plane of the conversion to a planar
         * picture of the source's own sub-sampling (4:2:0 for the other sources), chroma into scratch */
        int df = (src_fmt == 0 || src_fmt == 4 || src_fmt == 5 || src_fmt == 6 || src_fmt == 7 || src_fmt == 31) ? src_fmt : 0;
        {   /* 9 / 10 / 16-bit planar sources count as planar yuv too: at the same size the reference runs planarCopyWrapper's depth conversion
             * (not restated: the twin of the source's sub-sampling makes the refusal below apply) */
            int f = src_fmt;
            if (f == 61 || f == 63 || f == 65 || f == 67 || f == 69 || f == 71) f += 1; else if (f == 48 || f == 50 || f == 52) f -= 1;
            if (f == 72 || f == 66 || f == 49) df = 4; else if (f == 68 || f == 70 || f == 51) df = 5;
        }
        const int pitch = dw + 64;
        uint8_t *tmp = calloc((size_t)pitch * (dh + 2), 2);
        if (!tmp) return -1;
        uint8_t *const d3[3] = { dst[0], tmp, tmp + (size_t)pitch * (dh + 2) };
        const int ds3[3] = { dstride[0], pitch, pitch };
        {   /* (the unscaled converters of packed sources -- rgb24toyv12, yuyvtoyuv420 ... -- are installed for real yuv420p / yuv422p destinations, swscale_unscaled.c:
             * 1063-1067,1140-1147: a gray8 destination goes through swscale()) */
            const int packed = src_fmt == 1 || src_fmt == 2 || src_fmt == 3 || src_fmt == 15 || (src_fmt >= 25 && src_fmt <= 28);
            const int keep = g_nospecial;
            if (packed) g_nospecial = 1;
            r = sws_any(src_fmt, src, ss, sw, sh, df, d3, ds3, dw, dh, flags);
            g_nospecial = keep;
        }
        free(tmp);
        return r;
    }
    const int pk = dst_fmt == 1 || dst_fmt == 15;
    {   /* 9 / 10 / 16-bit planar sources: yuv420p 62 64 47, yuv422p 72 66 49, yuv444p 68 70 51 (LE), big-endian twins -1 / +1 */
        int f = src_fmt, be = 0, bits = 0, base = -1;
        if (f == 61 || f == 63 || f == 65 || f == 67 || f == 69 || f == 71) { be = 1; f += 1; }
        else if (f == 48 || f == 50 || f == 52) { be = 1; f -= 1; }
        switch (f) {
        case 62: base = 0; bits = 9; break;   case 64: base = 0; bits = 10; break;  case 47: base = 0; bits = 16; break;
        case 72: base = 4; bits = 9; break;   case 66: base = 4; bits = 10; break;  case 49: base = 4; bits = 16; break;
        case 68: base = 5; bits = 9; break;   case 70: base = 5; bits = 10; break;  case 51: base = 5; bits = 16; break;
        }
        if (base >= 0 && g_sbits == 8) {
            int h2, v2, b2 = 8;
            if (IS_RGB48(dst_fmt)) return -1;          /* (hScale16To19_c lines: not restated) */
            const int rgbd = dst_fmt == 2 || dst_fmt == 3 || (dst_fmt >= 25 && dst_fmt <= 28) || IS_RGB16(dst_fmt), pkd = dst_fmt == 1 || dst_fmt == 15, nvd = g_nospecial;
            const int shs = base == 5 ? 0 : 1, svs = base == 0 ? 1 : 0;
            if (uses_filter()) return -1;
            if (!rgbd && !pkd) {
                if (!planar_dst(dst_fmt, &h2, &v2, &b2) || b2 == 16) { g_dbe = 0; return -1; }        /* 19-bit lines: not restated for these sources */
                /* same size and sub-sampling: planarCopyWrapper with its own depth conversions (swscale_unscaled.c:793-1020), not restated */
                if (sw == dw && sh == dh && !nvd && h2 == shs && v2 == svs) { g_dbe = 0; return -1; }
            }
            g_sbits = bits; g_sbe = be;
            r = sws_any(base, src, ss, sw, sh, dst_fmt, dst, dstride, dw, dh, flags);
            g_sbits = 8; g_sbe = 0;
            return r;
        }
    }
    /* handle_jpeg() (utils.c:855-873) on both sides: yuvj420p 12 / 422p 13 / 444p 14 / 440p 32 are their limited-range twins with
     * srcRange / dstRange = 1.  An rgb destination folds the source range into its colour tables; a yuv destination of the other range
     * gets the range conversion (restated for planar 8-bit sources to planar 8 / 9 / 10-bit destinations), of the same range nothing. */
    const int src_j = (src_fmt >= 12 && src_fmt <= 14) || src_fmt == 32, dst_j = (dst_fmt >= 12 && dst_fmt <= 14) || dst_fmt == 32;
    if (src_j || dst_j) {
        const int sf = src_fmt == 12 ? 0 : src_fmt == 13 ? 4 : src_fmt == 14 ? 5 : src_fmt == 32 ? 31 : src_fmt;
        const int df = dst_fmt == 12 ? 0 : dst_fmt == 13 ? 4 : dst_fmt == 14 ? 5 : dst_fmt == 32 ? 31 : dst_fmt;
        const int dst_rgb = df == 2 || df == 3 || (df >= 25 && df <= 28) || IS_RGB16(df) || IS_RGB48(df);
        if (dst_rgb) {
            g_cs_jpeg = 1;
            r = sws_any(sf, src, ss, sw, sh, df, dst, dstride, dw, dh, flags);
            g_cs_jpeg = 0;
            return r;
        }
        if (src_j == dst_j) return sws_any(sf, src, ss, sw, sh, df, dst, dstride, dw, dh, flags);
        int h2, v2, b2;
        /* any source the port restates; destinations: planar 8 / 9 / 10 bit (also as the inner planes of an nv12 / nv21 destination) and packed 4:2:2 */
        if (df != 1 && df != 15 && (!planar_dst(df, &h2, &v2, &b2) || b2 == 16)) { g_dbe = 0; return -1; }
        g_range = src_j ? 1 : 2;
        r = sws_any(sf, src, ss, sw, sh, df, dst, dstride, dw, dh, flags);
        g_range = 0;
        return r;
    }
    const int rgb = dst_fmt == 2 || dst_fmt == 3 || (dst_fmt >= 25 && dst_fmt <= 28) || pk || IS_RGB16(dst_fmt) || IS_RGB48(dst_fmt);
    for (int wv = 1; wv < 4; wv += 2)                      /* asymmetric vertical vectors (lumV, chrV): the reference's last rows depend on its ring buffer state */
        for (int i = 0; i < g_fv[wv][0].length / 2; i++)
            if (g_fv[wv][0].coeff[i] != g_fv[wv][0].coeff[g_fv[wv][0].length - 1 - i]) return -1;
    if (uses_filter()) {        /* restated for planar 8-bit yuv sources to packed rgb / planar 8-bit yuv destinations only */
        int h2, v2, b2 = 8;
        const int planar_src = src_fmt == 0 || src_fmt == 4 || src_fmt == 5 || src_fmt == 6 || src_fmt == 7 || src_fmt == 31;
        if (!planar_src || pk || g_nospecial || (!rgb && (!planar_dst(dst_fmt, &h2, &v2, &b2) || b2 != 8))) { g_dbe = 0; return -1; }
    }
    if (pk && sw == dw && sh == dh && !g_range && g_sbits == 8) {
        /* the reference's unscaled converters to packed 4:2:2 (swscale_unscaled.c:1123-1139,1152-1176): from yuv422p always, from yuv420p
         * with the fast-bilinear / point flags (yuvPlanartoyuy2_c, rgb2rgb_template.c:322-420: width >> 1 pairs), same format = copy */
        const int vshift = src_fmt == 0 ? 1 : 0, uyvy = dst_fmt == 15;
        if (src_fmt == 4 || (src_fmt == 0 && (flags & (F_FAST_BILINEAR | F_POINT)))) {
            /* the 64-bit build of the loop (HAVE_FAST_64BIT, :374-386) converts two pairs per step: an odd width >> 1 is rounded up, the
             * extra pair reads / writes past the nominal width (here: when the rows have room for it) */
            const int pairs = sw >> 1, pairs_r = (pairs + 1) & ~1;
            const int extra = pairs_r > pairs && dstride[0] >= 4 * pairs_r && ss[0] >= 2 * pairs_r && ss[1] >= pairs_r && ss[2] >= pairs_r;
            for (int y = 0; y < sh; y++)
                for (int i = 0; i < (extra ? pairs_r : pairs); i++) {
                    uint8_t *d = dst[0] + (size_t)y * dstride[0] + 4 * i;
                    const uint8_t *py = src[0] + (size_t)y * ss[0] + 2 * i;
                    const uint8_t u = src[1][(size_t)(y >> vshift) * ss[1] + i], v = src[2][(size_t)(y >> vshift) * ss[2] + i];
                    if (uyvy) { d[0] = u; d[1] = py[0]; d[2] = v; d[3] = py[1]; } else { d[0] = py[0]; d[1] = u; d[2] = py[1]; d[3] = v; }
                }
            return sh;
        }
        if (src_fmt == dst_fmt) {
            for (int y = 0; y < sh; y++) memcpy(dst[0] + (size_t)y * dstride[0], src[0] + (size_t)y * ss[0], (size_t)sw * 2);
            return sh;
        }
    }
    if (pk || IS_RGB16(dst_fmt) || IS_RGB48(dst_fmt)) flags &= ~F_FULL_CHR_H_INT;
    if ((IS_RGB16(dst_fmt) || IS_RGB48(dst_fmt)) && uses_filter()) return -1;
    if (IS_RGB48(dst_fmt) && (src_fmt == 23 || src_fmt == 24)) return -1;      /* (48-bit destinations: planar 8-bit yuv sources only) */
    if (dst_fmt == 27 && (flags & F_FULL_CHR_H_INT) &&
        !(sw == dw && sh == dh && (src_fmt == 2 || src_fmt == 3 || src_fmt == 11 || (src_fmt >= 25 && src_fmt <= 28)) && !g_nospecial))      /* (rgbToRgbWrapper / palToRgbWrapper never reach that function) */
        return -1;    /* yuv2rgb_full_X_c advances twice per abgr pixel (output.c:1231-1237): no defined result */
    if (!rgb && !planar_dst(dst_fmt, &g_dhs, &g_dvs, &g_dbits)) return -1;
    switch (src_fmt) {
    case 0: hs = 1; vs = 1; break;  case 4: hs = 1; vs = 0; break;  case 5: hs = 0; vs = 0; break;
    case 6: hs = 2; vs = 2; break;  case 7: hs = 2; vs = 0; break;  case 31: hs = 0; vs = 1; break;
    case 23: case 24:                                   /* src[0] luma, src[1] interleaved chroma */
        r = orc_sws_nv12(src_fmt == 24, src[0], ss[0], src[1], ss[1], sw, sh, dst_fmt, dst, dstride, dw, dh, flags);
        g_dhs = g_dvs = 1; g_dbits = 8; g_dbe = 0;
        return r;
    case 11:                                            /* pal8: src[0] indices, src[1] the palette */
        if (!src[1] || g_sbits != 8) { g_dhs = g_dvs = 1; g_dbits = 8; g_dbe = 0; return -1; }
        r = pal8_source(src[0], ss[0], src[1], sw, sh, dst_fmt, dst, dstride, dw, dh, flags);
        g_dhs = g_dvs = 1; g_dbits = 8; g_dbe = 0;
        return r;
    case 1: case 2: case 3: case 15: case 25: case 26: case 27: case 28:
        if (IS_RGB16(dst_fmt) || IS_RGB48(dst_fmt)) return -1;   /* (rgb2rgb converter families and readers in front of the 16 / 48-bpp output stage: not restated) */
        r = packed_source(src_fmt, src[0], ss[0], sw, sh, dst_fmt, dst, dstride, dw, dh, flags);
        g_dhs = g_dvs = 1; g_dbits = 8; g_dbe = 0;
        return r;
    default: g_dhs = g_dvs = 1; g_dbits = 8; g_dbe = 0; return -1;
    }
    /* yuv410p -> yuv420p of the same size without SWS_BITEXACT is the reference's yvu9ToYv12Wrapper (swscale_unscaled.c:1057-1061,
     * rgb2rgb.c planar2x): not restated */
    if (src_fmt == 6 && dst_fmt == 0 && sw == dw && sh == dh && !(flags & F_BITEXACT) && !g_nospecial && !g_range) return -1;
    g_hs = hs; g_vs = vs;
    r = rgb ? to_rgb_or_bgr(src, ss, sw, sh, dst_fmt, dst[0], dstride[0], dw, dh, flags)
            : orc_sws_yuv420p_to_yuv420p(src, ss, sw, sh, dst, dstride, dw, dh, flags);
    g_hs = 1; g_vs = 1; g_dhs = g_dvs = 1; g_dbits = 8; g_dbe = 0;
    return r;
}

/* ---- SwsContext per-line slots -------------------------------------------------------------------------------------------------
 * The functions sws_init_swscale() / ff_sws_init_output_funcs() install (swscale.c:723-769, output.c:1357-1590), restated per line. */
int orc_sws_line_range(int kind, int16_t *dst1, int16_t *dst2, int width)
{
    for (int i = 0; i < width; i++) {
        dst1[i] = range_sample(dst1[i], kind < 2 ? 1 : 2, kind & 1);
        if (kind & 1) dst2[i] = range_sample(dst2[i], kind < 2 ? 1 : 2, 1);
    }
    return 0;
}

int orc_sws_line_hscale(int dst_fmt, int flags, void *dst, int dstW, const uint8_t *src, const int16_t *filter, const int32_t *filterPos, int filterSize)
{
    int hs, vs, bits;
    (void)flags;
    const int wide = planar_dst(dst_fmt, &hs, &vs, &bits) && bits == 16;      /* hScale8To19_c (swscale.c:149-164) feeds the 16-bit output functions */
    for (int i = 0; i < dstW; i++) {
        int val = 0;
        for (int j = 0; j < filterSize; j++) val += src[filterPos[i] + j] * filter[filterSize * i + j];
        if (wide) ((int32_t *)dst)[i] = (val >> 3) < (1 << 19) - 1 ? val >> 3 : (1 << 19) - 1;
        else      ((int16_t *)dst)[i] = (int16_t)((val >> 7) < (1 << 15) - 1 ? val >> 7 : (1 << 15) - 1);      /* swscale.c:133-147 */
    }
    return 0;
}

int orc_sws_line_hfast(int chroma, int16_t *dst1, int16_t *dst2, int dstW, const uint8_t *src1, const uint8_t *src2, int srcW, int xInc)
{
    (void)srcW;                            /* swscale.c:238-250, 286-299: the loops never look at srcW */
    unsigned xpos = 0;
    for (int i = 0; i < dstW; i++, xpos += (unsigned)xInc) {
        const unsigned xx = xpos >> 16, xa = (xpos & 0xFFFF) >> 9;
        if (!chroma) dst1[i] = (int16_t)((src1[xx] << 7) + (src1[xx + 1] - src1[xx]) * (int)xa);
        else {
            dst1[i] = (int16_t)(src1[xx] * (int)(xa ^ 127) + src1[xx + 1] * (int)xa);
            dst2[i] = (int16_t)(src2[xx] * (int)(xa ^ 127) + src2[xx + 1] * (int)xa);
        }
    }
    return 0;
}

int orc_sws_line_plane(int dst_fmt, const int16_t *filter, int filterSize, const void *const *src, uint8_t *dest, int dstW, const uint8_t *dither, int offset)
{
    int hs, vs, bits = 8;
    if (!planar_dst(dst_fmt, &hs, &vs, &bits)) { bits = 8; g_dbe = 0; }      /* packed and nv12 / nv21 destinations get the 8-bit functions (output.c:1385-1390) */
    const int be = g_dbe;
    for (int i = 0; i < dstW; i++) {
        int v;
        if (bits == 8) {                   /* output.c:242-265 */
            if (!filterSize) v = (((const int16_t *)src[0])[i] + dither[(i + offset) & 7]) >> 7;
            else {
                v = dither[(i + offset) & 7] << 12;
                for (int j = 0; j < filterSize; j++) v += ((const int16_t *)src[j])[i] * filter[j];
                v >>= 19;
            }
            dest[i] = u8clip(v);
            continue;
        }
        if (bits == 16) {                  /* output.c:136-172 on 19-bit int32 lines */
            if (!filterSize) { v = (((const int32_t *)src[0])[i] + 4) >> 3; v = v < 0 ? 0 : v > 65535 ? 65535 : v; }
            else {
                unsigned acc = (1u << 14) - 0x40000000u;
                for (int j = 0; j < filterSize; j++) acc += (unsigned)(((const int32_t *)src[j])[i] * filter[j]);
                v = (int)acc >> 15; v = (v < -32768 ? -32768 : v > 32767 ? 32767 : v) + 0x8000;
            }
        } else {                           /* output.c:183-213 */
            if (!filterSize) v = (((const int16_t *)src[0])[i] + (1 << (14 - bits))) >> (15 - bits);
            else {
                v = 1 << (26 - bits);
                for (int j = 0; j < filterSize; j++) v += ((const int16_t *)src[j])[i] * filter[j];
                v >>= 27 - bits;
            }
            v = v < 0 ? 0 : v > (1 << bits) - 1 ? (1 << bits) - 1 : v;
        }
        if (be) { dest[2 * i] = (uint8_t)(v >> 8); dest[2 * i + 1] = (uint8_t)v; } else { dest[2 * i] = (uint8_t)v; dest[2 * i + 1] = (uint8_t)(v >> 8); }
    }
    return 0;
}

int orc_sws_line_nv12(int dst_fmt, const int16_t *chrFilter, int chrFilterSize, const int16_t *const *chrU, const int16_t *const *chrV, uint8_t *dest, int chrDstW)
{
    if (dst_fmt != 23 && dst_fmt != 24) return -1;          /* output.c:1388-1389 */
    for (int i = 0; i < chrDstW; i++) {                     /* output.c:267-303, chrDither8 = 64 */
        int u = 64 << 12, v = 64 << 12;
        for (int j = 0; j < chrFilterSize; j++) { u += chrU[j][i] * chrFilter[j]; v += chrV[j][i] * chrFilter[j]; }
        dest[2 * i + (dst_fmt == 24)] = u8clip(u >> 19);
        dest[2 * i + (dst_fmt != 24)] = u8clip(v >> 19);
    }
    return 0;
}

int orc_sws_line_packed(int dst_fmt, int flags, int kind, const int16_t *lumFilter, const int16_t *const *lumSrc, int lumFilterSize,
                        const int16_t *chrFilter, const int16_t *const *chrU, const int16_t *const *chrV, int chrFilterSize,
                        uint8_t *dest, int dstW, int yalpha, int uvalpha, int y)
{
    (void)y;
    const int is422 = dst_fmt == 1 || dst_fmt == 15, is32 = dst_fmt >= 25 && dst_fmt <= 28;
    if (!is422 && !is32 && dst_fmt != 2 && dst_fmt != 3) return -1;
    /* byte positions of r, g, b (and alpha = 255) in a pixel: rgb24, bgr24, argb, rgba, abgr, bgra */
    const int bpp = is32 ? 4 : 3;
    const int ro = dst_fmt == 2 ? 0 : dst_fmt == 3 ? 2 : dst_fmt == 25 ? 1 : dst_fmt == 26 ? 0 : dst_fmt == 27 ? 3 : 2;
    const int go = is32 ? (dst_fmt <= 26 ? ro + 1 : ro - 1) : 1;
    const int bo = dst_fmt == 2 ? 2 : dst_fmt == 3 ? 0 : dst_fmt == 25 ? 3 : dst_fmt == 26 ? 2 : dst_fmt == 27 ? 1 : 0;
    const int ao = dst_fmt == 26 || dst_fmt == 28 ? 3 : 0;
    if ((flags & 0x2000) && !is422) {                      /* SWS_FULL_CHR_H_INT: yuv2rgb_full_X_c_template (output.c:1165-1250), only yuv2packedX exists */
        if (kind) return -1;
        int64_t kcy, koy, kcrv, kcbu, kcgu, kcgv; int kyoffs;
        cs_coeffs(&kcy, &koy, &kcrv, &kcbu, &kcgu, &kcgv, &kyoffs);
#define R16(f) ((int16_t)({ int r_ = (int)(((int64_t)(f) + (1 << 15)) >> 16); r_ < -0x7FFF ? -0x8000 : r_ > 0x7FFF ? 0x7FFF : r_; }))
        const int y_coeff = R16(kcy << 13), y_offset = R16(koy << 9), v2r = R16(kcrv << 13), v2g = R16(kcgv << 13), u2g = R16(kcgu << 13), u2b = R16(kcbu << 13);
#undef R16
        for (int i = 0; i < dstW; i++) {
            int Y = 0, U = -128 * (1 << 19), V = -128 * (1 << 19);
            for (int j = 0; j < lumFilterSize; j++) Y += lumSrc[j][i] * lumFilter[j];
            for (int j = 0; j < chrFilterSize; j++) { U += chrU[j][i] * chrFilter[j]; V += chrV[j][i] * chrFilter[j]; }
            Y >>= 10; U >>= 10; V >>= 10;
            Y = (Y - y_offset) * y_coeff + (1 << 21);
            int R = Y + V * v2r, G = Y + V * v2g + U * u2g, B = Y + U * u2b;
            if ((R | G | B) & 0xC0000000) {
                R = R < 0 ? 0 : R > 0x3FFFFFFF ? 0x3FFFFFFF : R; G = G < 0 ? 0 : G > 0x3FFFFFFF ? 0x3FFFFFFF : G; B = B < 0 ? 0 : B > 0x3FFFFFFF ? 0x3FFFFFFF : B;
            }
            uint8_t *d = dest + (size_t)bpp * i;
            d[ro] = (uint8_t)(R >> 22); d[go] = (uint8_t)(G >> 22); d[bo] = (uint8_t)(B >> 22);
            if (is32) d[ao] = 255;
        }
        return 0;
    }
    uint8_t ytab[1024]; int32_t rv[256], gu[256], gv[256], bu[256];
    orc_sws_rgb24_tables(ytab, rv, gu, gv, bu);
    for (int i = 0; i < (dstW + 1) >> 1; i++) {
        int Y1, Y2, U, V;
        if (kind == 1) {                                    /* output.c:1042-1110 / 531-576 */
            Y1 = lumSrc[0][2 * i] >> 7; Y2 = lumSrc[0][2 * i + 1] >> 7;
            if (uvalpha < 2048) { U = chrU[0][i] >> 7; V = chrV[0][i] >> 7; }
            else { U = (chrU[0][i] + chrU[1][i]) >> 8; V = (chrV[0][i] + chrV[1][i]) >> 8; }
            Y1 = u8clip(Y1); Y2 = u8clip(Y2); U = u8clip(U); V = u8clip(V);
        } else if (kind == 2) {                             /* output.c:997-1040 / 498-529 */
            const int ya1 = 4096 - yalpha, ua1 = 4096 - uvalpha;
            Y1 = (lumSrc[0][2 * i] * ya1 + lumSrc[1][2 * i] * yalpha) >> 19;
            Y2 = (lumSrc[0][2 * i + 1] * ya1 + lumSrc[1][2 * i + 1] * yalpha) >> 19;
            U = (chrU[0][i] * ua1 + chrU[1][i] * uvalpha) >> 19;
            V = (chrV[0][i] * ua1 + chrV[1][i] * uvalpha) >> 19;
            Y1 = u8clip(Y1); Y2 = u8clip(Y2); U = u8clip(U); V = u8clip(V);
        } else {                                            /* output.c:936-995 / 456-496 */
            Y1 = Y2 = U = V = 1 << 18;
            for (int j = 0; j < lumFilterSize; j++) { Y1 += lumSrc[j][2 * i] * lumFilter[j]; Y2 += lumSrc[j][2 * i + 1] * lumFilter[j]; }
            for (int j = 0; j < chrFilterSize; j++) { U += chrU[j][i] * chrFilter[j]; V += chrV[j][i] * chrFilter[j]; }
            Y1 >>= 19; Y2 >>= 19; U >>= 19; V >>= 19;
            if ((Y1 | Y2 | U | V) & 0x100) { Y1 = u8clip(Y1); Y2 = u8clip(Y2); U = u8clip(U); V = u8clip(V); }
        }
        if (is422) {
            uint8_t *d = dest + 4 * (size_t)i;
            if (dst_fmt == 1) { d[0] = (uint8_t)Y1; d[1] = (uint8_t)U; d[2] = (uint8_t)Y2; d[3] = (uint8_t)V; }
            else              { d[0] = (uint8_t)U; d[1] = (uint8_t)Y1; d[2] = (uint8_t)V; d[3] = (uint8_t)Y2; }
            continue;
        }
        const uint8_t *r = ytab + rv[V], *g = ytab + gu[U] + gv[V], *b = ytab + bu[U];
        uint8_t *d = dest + (size_t)bpp * 2 * i;
        d[ro] = r[Y1]; d[go] = g[Y1]; d[bo] = b[Y1]; d[bpp + ro] = r[Y2]; d[bpp + go] = g[Y2]; d[bpp + bo] = b[Y2];
        if (is32) d[ao] = d[bpp + ao] = 255;
    }
    return 0;
}
