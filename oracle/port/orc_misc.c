/*
 * oracle/port/orc_misc.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the remaining 8-bit DSP tables:
 *   FDCTDSPContext   jpeg_fdct_islow_8 / fdct248_islow_8     libavcodec/jfdctint_template.c:182-398
 *                    fdct_ifast / fdct_ifast248              libavcodec/jfdctfst.c:141-332
 *   MECmpContext     pix_abs / sad / sse / hadamard8 / vsad / vsse / nsse / sum_abs_dctelem
 *                                                            libavcodec/me_cmp.c:29-357, :434-536, :784-885
 *   full search      motion_est_template.c:620-655 with get_limits (motion_est.c:517-548), lambda 0
 *   HpelDSPContext   put / avg / no_rnd x {full, x2, y2, xy2}   libavcodec/hpeldsp.c:38-328
 * Each metric is written as a per-sample formula over a generic w x h window instead of the reference's
 * unrolled per-width functions.  Pinned against oracle/_ref in tests/test_oracle_misc_cpu.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../oracle_api.h"

/* ------------------------------------------------------------------ forward DCT ---------------- */
enum { K0298 = 2446, K0390 = 3196, K0541 = 4433, K0765 = 6270, K0899 = 7373, K1175 = 9633, K1501 = 12299,
       K1847 = 15137, K1961 = 16069, K2053 = 16819, K2562 = 20995, K3072 = 25172 };
static inline int rshift_round(int v, int n) { return (v + (1 << (n - 1))) >> n; }

/* one accurate 8-point pass; even outputs 0,4 are scaled by `up` (<0: rounded right shift), the rest descaled by `dn` */
static void islow_1d(const int in[8], int out[8], int up, int dn)
{
    int s0 = in[0] + in[7], d0 = in[0] - in[7], s1 = in[1] + in[6], d1 = in[1] - in[6];
    int s2 = in[2] + in[5], d2 = in[2] - in[5], s3 = in[3] + in[4], d3 = in[3] - in[4];
    int e0 = s0 + s3, e3 = s0 - s3, e1 = s1 + s2, e2 = s1 - s2;
    out[0] = up >= 0 ? (e0 + e1) * (1 << up) : rshift_round(e0 + e1, -up);
    out[4] = up >= 0 ? (e0 - e1) * (1 << up) : rshift_round(e0 - e1, -up);
    int z = (e2 + e3) * K0541;
    out[2] = rshift_round(z + e3 * K0765, dn);
    out[6] = rshift_round(z - e2 * K1847, dn);
    int z1 = d3 + d0, z2 = d2 + d1, z3 = d3 + d1, z4 = d2 + d0, z5 = (z3 + z4) * K1175;
    int t4 = d3 * K0298, t5 = d2 * K2053, t6 = d1 * K3072, t7 = d0 * K1501;
    z1 *= -K0899; z2 *= -K2562; z3 = z3 * -K1961 + z5; z4 = z4 * -K0390 + z5;
    out[7] = rshift_round(t4 + z1 + z3, dn);
    out[5] = rshift_round(t5 + z2 + z4, dn);
    out[3] = rshift_round(t6 + z2 + z3, dn);
    out[1] = rshift_round(t7 + z1 + z4, dn);
}
/* 2-4-8 column pass (two interleaved 4-point DCTs), jfdctint_template.c:342-398 */
static void islow_248_col(const int in[8], int out[8], int sh, int dn)       /* sh = OUT_SHIFT, dn = CONST_BITS + OUT_SHIFT */
{
    int a0 = in[0] + in[1], a1 = in[2] + in[3], a2 = in[4] + in[5], a3 = in[6] + in[7];
    int b0 = in[0] - in[1], b1 = in[2] - in[3], b2 = in[4] - in[5], b3 = in[6] - in[7];
    int e0 = a0 + a3, e1 = a1 + a2, e2 = a1 - a2, e3 = a0 - a3, z;
    out[0] = rshift_round(e0 + e1, sh); out[4] = rshift_round(e0 - e1, sh);
    z = (e2 + e3) * K0541;
    out[2] = rshift_round(z + e3 * K0765, dn); out[6] = rshift_round(z - e2 * K1847, dn);
    e0 = b0 + b3; e1 = b1 + b2; e2 = b1 - b2; e3 = b0 - b3;
    out[1] = rshift_round(e0 + e1, sh); out[5] = rshift_round(e0 - e1, sh);
    z = (e2 + e3) * K0541;
    out[3] = rshift_round(z + e3 * K0765, dn); out[7] = rshift_round(z - e2 * K1847, dn);
}
/* AAN "ifast": products are shifted down by 8 without rounding and truncated to int16 (jfdctfst.c MULTIPLY) */
static inline int fmul(int v, int k) { return (int16_t)((v * k) >> 8); }
static void ifast_1d(const int in[8], int out[8])
{
    int s0 = in[0] + in[7], d0 = in[0] - in[7], s1 = in[1] + in[6], d1 = in[1] - in[6];
    int s2 = in[2] + in[5], d2 = in[2] - in[5], s3 = in[3] + in[4], d3 = in[3] - in[4];
    int e0 = s0 + s3, e3 = s0 - s3, e1 = s1 + s2, e2 = s1 - s2;
    out[0] = e0 + e1; out[4] = e0 - e1;
    int z1 = fmul(e2 + e3, 181);
    out[2] = e3 + z1; out[6] = e3 - z1;
    int p0 = d3 + d2, p1 = d2 + d1, p2 = d1 + d0;
    int z5 = fmul(p0 - p2, 98), z2 = fmul(p0, 139) + z5, z4 = fmul(p2, 334) + z5, z3 = fmul(p1, 181);
    int z11 = d0 + z3, z13 = d0 - z3;
    out[5] = z13 + z2; out[3] = z13 - z2; out[1] = z11 + z4; out[7] = z11 - z4;
}
static void ifast_248_col(const int in[8], int out[8])
{
    int a0 = in[0] + in[1], a1 = in[2] + in[3], a2 = in[4] + in[5], a3 = in[6] + in[7];
    int b0 = in[0] - in[1], b1 = in[2] - in[3], b2 = in[4] - in[5], b3 = in[6] - in[7];
    int e0 = a0 + a3, e1 = a1 + a2, e2 = a1 - a2, e3 = a0 - a3, z;
    out[0] = e0 + e1; out[4] = e0 - e1; z = fmul(e2 + e3, 181); out[2] = e3 + z; out[6] = e3 - z;
    e0 = b0 + b3; e1 = b1 + b2; e2 = b1 - b2; e3 = b0 - b3;
    out[1] = e0 + e1; out[5] = e0 - e1; z = fmul(e2 + e3, 181); out[3] = e3 + z; out[7] = e3 - z;
}

void orc_fdct(int which, int16_t *b)
{
    /* which 4 / 5: jpeg_fdct_islow_10 / fdct248_islow_10 -- the same code with PASS1_BITS 1, OUT_SHIFT 2 (jfdctint_template.c:126-130) */
    const int ten = which >= 4, p1 = ten ? 1 : 4, osh = ten ? 2 : 4;
    int in[8], out[8];
    for (int r = 0; r < 8; r++) {                 /* rows: results go back as int16 */
        for (int k = 0; k < 8; k++) in[k] = b[8 * r + k];
        if (which < 2 || ten) islow_1d(in, out, p1, 13 - p1); else ifast_1d(in, out);
        for (int k = 0; k < 8; k++) b[8 * r + k] = (int16_t)out[k];
    }
    for (int c = 0; c < 8; c++) {
        for (int k = 0; k < 8; k++) in[k] = b[8 * k + c];
        switch (which) {
        case 0: case 4: islow_1d(in, out, -osh, 13 + osh); break;
        case 1: case 5: islow_248_col(in, out, osh, 13 + osh); break;
        case 2: ifast_1d(in, out); break;
        default: ifast_248_col(in, out); break;
        }
        for (int k = 0; k < 8; k++) b[8 * k + c] = (int16_t)out[k];
    }
}

/* ------------------------------------------------------------------ compare functions ---------- */
static inline int iabs(int v) { return v < 0 ? -v : v; }

static int sad_generic(const uint8_t *a, const uint8_t *b, ptrdiff_t st, int w, int h, int dxy)
{
    int s = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = b + y * st + x;
            int r = dxy == 0 ? p[0] : dxy == 1 ? (p[0] + p[1] + 1) >> 1 : dxy == 2 ? (p[0] + p[st] + 1) >> 1
                                                                        : (p[0] + p[1] + p[st] + p[st + 1] + 2) >> 2;
            s += iabs(a[y * st + x] - r);
        }
    return s;
}
static int sse_generic(const uint8_t *a, const uint8_t *b, ptrdiff_t st, int w, int h)
{
    int s = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) { int d = a[y * st + x] - b[y * st + x]; s += d * d; }
    return s;
}
static void wht8(int *v, int step)
{
    for (int len = 1; len < 8; len <<= 1)
        for (int i = 0; i < 8; i += 2 * len)
            for (int j = i; j < i + len; j++) { int p = v[j * step], q = v[(j + len) * step]; v[j * step] = p + q; v[(j + len) * step] = p - q; }
}
static int hadamard8(const uint8_t *a, const uint8_t *b, ptrdiff_t st, int intra)
{
    int t[64], s = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) t[8 * y + x] = intra ? a[y * st + x] : b[y * st + x] - a[y * st + x];
    for (int y = 0; y < 8; y++) wht8(t + 8 * y, 1);
    for (int x = 0; x < 8; x++) wht8(t + x, 8);
    for (int i = 0; i < 64; i++) s += iabs(t[i]);
    return intra ? s - iabs(t[0]) : s;
}
static int hadamard_wrap(const uint8_t *a, const uint8_t *b, ptrdiff_t st, int w16, int h, int intra)
{
    if (!w16) return hadamard8(a, b, st, intra);
    int s = hadamard8(a, b, st, intra) + hadamard8(a + 8, b + 8, st, intra);
    if (h == 16) s += hadamard8(a + 8 * st, b + 8 * st, st, intra) + hadamard8(a + 8 * st + 8, b + 8 * st + 8, st, intra);
    return s;
}

int orc_me_cmp(int kind, int sidx, int dxy, const uint8_t *a, const uint8_t *b, ptrdiff_t st, int h)
{
    int w = sidx == 0 ? 16 : sidx == 1 ? 8 : 4, s = 0;
    switch (kind) {
    case 0: return sidx > 1 ? -1 : sad_generic(a, b, st, w, h, dxy);
    case 1: return sidx > 1 ? -1 : sad_generic(a, b, st, w, h, 0);
    case 2: return sidx > 2 ? -1 : sse_generic(a, b, st, w, h);
    case 3: return sidx > 1 ? -1 : hadamard_wrap(a, b, st, sidx == 0, h, 0);
    case 7: return sidx > 1 ? -1 : hadamard_wrap(a, b, st, sidx == 0, h, 1);
    case 4: case 5:                                     /* vsad16 / vsse16 only (me_cmp.c:930-934) */
        if (sidx != 0) return -1;
        for (int y = 1; y < h; y++)
            for (int x = 0; x < 16; x++) {
                int d = a[(y - 1) * st + x] - b[(y - 1) * st + x] - a[y * st + x] + b[y * st + x];
                s += kind == 4 ? iabs(d) : d * d;
            }
        return s;
    case 8: case 9:                                     /* vsad_intra / vsse_intra, 16 and 8 wide */
        if (sidx > 1) return -1;
        for (int y = 1; y < h; y++)
            for (int x = 0; x < w; x++) { int d = a[(y - 1) * st + x] - a[y * st + x]; s += kind == 8 ? iabs(d) : d * d; }
        return s;
    case 6: {                                           /* nsse, weight 8 (NULL context) */
        if (sidx > 1) return -1;
        int s2 = 0;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) { int d = a[y * st + x] - b[y * st + x]; s += d * d; }
            if (y + 1 < h)
                for (int x = 0; x < w - 1; x++) {
                    const uint8_t *p = a + y * st + x, *q = b + y * st + x;
                    s2 += iabs(p[0] - p[st] - p[1] + p[st + 1]) - iabs(q[0] - q[st] - q[1] + q[st + 1]);
                }
        }
        return s + iabs(s2) * 8;
    }
    case 10: { const int16_t *c = (const int16_t *)a; for (int i = 0; i < 64; i++) s += iabs(c[i]); return s; }
    case 11: case 12: case 13: {
        /* encoder-state metrics that only need the DSP tables: dct_sad8x8_c / dct_max8x8_c (me_cmp.c:538-548, :606-621)
         * = diff_pixels -> fdsp.fdct (dxy: 0 islow, 2 ifast) -> sum / max of |coefficient|; dct264_sad8x8_c (:551-603) =
         * diff -> the H.264 8x8 forward transform, rows then columns, sum of |result|.  16-wide = sum over the 8x8
         * quadrants the wrapper visits (WRAPPER8_16_SQ, :859-874: two for h == 8, four for h == 16). */
        if (sidx > 1) return -1;
        const int nblk = sidx == 0 ? (h == 16 ? 4 : 2) : 1;
        for (int q = 0; q < nblk; q++) {
            const uint8_t *pa = a + (q & 1) * 8 + (q >> 1) * 8 * st, *pb = b + (q & 1) * 8 + (q >> 1) * 8 * st;
            int16_t t[64];
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) t[8 * y + x] = (int16_t)(pa[y * st + x] - pb[y * st + x]);
            if (kind != 13) {
                orc_fdct(dxy == 2 ? 2 : 0, t);
                int m = 0;
                for (int i = 0; i < 64; i++) { if (kind == 11) s += iabs(t[i]); else if (iabs(t[i]) > m) m = iabs(t[i]); }
                s += m;
            } else {
                for (int pass = 0; pass < 2; pass++)
                    for (int i = 0; i < 8; i++) {
                        int v[8], o[8];
                        for (int k = 0; k < 8; k++) v[k] = pass ? t[8 * k + i] : t[8 * i + k];
                        const int s07 = v[0] + v[7], s16 = v[1] + v[6], s25 = v[2] + v[5], s34 = v[3] + v[4];
                        const int a0 = s07 + s34, a1 = s16 + s25, a2 = s07 - s34, a3 = s16 - s25;
                        const int d07 = v[0] - v[7], d16 = v[1] - v[6], d25 = v[2] - v[5], d34 = v[3] - v[4];
                        const int a4 = d16 + d25 + (d07 + (d07 >> 1)), a5 = d07 - d34 - (d25 + (d25 >> 1));
                        const int a6 = d07 + d34 - (d16 + (d16 >> 1)), a7 = d16 - d25 + (d34 + (d34 >> 1));
                        o[0] = a0 + a1; o[1] = a4 + (a7 >> 2); o[2] = a2 + (a3 >> 1); o[3] = a5 + (a6 >> 2);
                        o[4] = a0 - a1; o[5] = a6 - (a5 >> 2); o[6] = (a2 >> 1) - a3; o[7] = (a4 >> 2) - a7;
                        for (int k = 0; k < 8; k++) { if (pass) s += iabs(o[k]); else t[8 * i + k] = (int16_t)o[k]; }
                    }
            }
        }
        return s;
    }
    }
    return -1;
}

struct fs_span { const uint8_t *cur, *ref; int stride, w, h, range, y0, y1; int32_t *out; };
static void *fs_run(void *arg)
{
    struct fs_span *j = arg;
    int mbw = j->w / 16;
    for (int mby = j->y0; mby < j->y1; mby++)
        for (int mbx = 0; mbx < mbw; mbx++) {
            int px = 16 * mbx, py = 16 * mby;
            int x0 = -px > -j->range ? -px : -j->range, x1 = j->w - 16 - px < j->range ? j->w - 16 - px : j->range;
            int y0 = -py > -j->range ? -py : -j->range, y1 = j->h - 16 - py < j->range ? j->h - 16 - py : j->range;
            const uint8_t *c = j->cur + py * j->stride + px;
            int best = 1 << 30, bx = 0, by = 0;
            for (int y = y0; y <= y1; y++)
                for (int x = x0; x <= x1; x++) {
                    int d = sad_generic(c, j->ref + (py + y) * j->stride + px + x, j->stride, 16, 16, 0);
                    if (d < best) { best = d; bx = x; by = y; }        /* strict <: first in raster order wins */
                }
            int32_t *o = j->out + 3 * (mby * mbw + mbx);
            o[0] = bx; o[1] = by; o[2] = best;
        }
    return NULL;
}
void orc_full_search(const uint8_t *cur, const uint8_t *ref, int stride, int w, int h, int range, int y0, int y1,
                     int32_t *out, int nthreads)
{
    enum { MAXT = 256 };
    pthread_t th[MAXT]; struct fs_span sp[MAXT];
    if (nthreads < 1) nthreads = 1;
    if (nthreads > MAXT) nthreads = MAXT;
    int rows = y1 - y0;
    for (int t = 0; t < nthreads; t++) {
        sp[t] = (struct fs_span){ cur, ref, stride, w, h, range, y0 + rows * t / nthreads, y0 + rows * (t + 1) / nthreads, out };
        if (nthreads > 1) pthread_create(&th[t], NULL, fs_run, &sp[t]); else fs_run(&sp[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ half-pel MC ---------------- */
int orc_hpel(int tab, int sidx, int dxy, uint8_t *block, const uint8_t *pixels, ptrdiff_t st, int h)
{
    int w = 16 >> sidx, no_rnd = tab >= 2, avg = tab & 1;
    if (avg && sidx == 3 && dxy == 3) avg = 0;      /* avg_pixels2_xy2_8_c stores without averaging ("FIXME non put", hpeldsp.c:151) */
    if ((tab == 2 && sidx > 1) || (tab == 3 && sidx != 0)) return -1;       /* slots the reference leaves NULL */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = pixels + y * st + x;
            int v;
            switch (dxy) {
            case 0: v = p[0]; break;
            case 1: v = (p[0] + p[1] + 1 - no_rnd) >> 1; break;
            case 2: v = (p[0] + p[st] + 1 - no_rnd) >> 1; break;
            default: v = (p[0] + p[1] + p[st] + p[st + 1] + 2 - no_rnd) >> 2; break;
            }
            /* avg_* always folds into the destination with the ROUNDING average (op_avg = rnd_avg32, hpeldsp.c:330);
             * only the interpolation itself honours no_rnd */
            block[y * st + x] = (uint8_t)(avg ? (block[y * st + x] + v + 1) >> 1 : v);
        }
    return 0;
}


/* PixblockDSPContext: get_pixels / diff_pixels, libavcodec/pixblockdsp_template.c:24-66 (8-bit) */
void orc_pixblock(int kind, int16_t *block, const uint8_t *s1, const uint8_t *s2, ptrdiff_t stride)
{
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            block[8 * y + x] = kind ? (int16_t)(s1[y * stride + x] - s2[y * stride + x]) : s1[y * stride + x];
}


/* QpelDSPContext: MPEG-4 quarter-pel MC (libavcodec/qpeldsp.c:39-700), one evaluation for all 16 phases:
 *   L(s, i)  = 8-tap (-1 3 -6 20 20 -6 3 -1) over s[0..N] with the indices mirrored at both ends (:46-126)
 *   Hx(r, c) = phase x of row r: F, avg(F, H), H, avg(F shifted by one, H), H = clip((L + rnd) >> 5)
 *   out(c, r)= phase y over the column of Hx: Hx, avg(Hx, V), V, avg(Hx one row down, V), V = clip((L over Hx + rnd) >> 5)
 * rnd = 16 (15 for the no_rnd table), the inner averages round up (down for no_rnd), the avg table finally averages
 * with dst rounding up. */
static int q_tap(const int *s, int n, int i)
{
#define QM(j) s[(j) < 0 ? -1 - (j) : (j) > n ? 2 * n + 1 - (j) : (j)]
    return (QM(i) + QM(i + 1)) * 20 - (QM(i - 1) + QM(i + 2)) * 6 + (QM(i - 2) + QM(i + 3)) * 3 - (QM(i - 3) + QM(i + 4));
#undef QM
}
static inline int q_clip(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

void orc_mpeg4_qpel(int kind, int sidx, int mc, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    const int n = sidx ? 8 : 16, x = mc & 3, y = mc >> 2, rnd = kind == 1 ? 15 : 16, up = kind == 1 ? 0 : 1;
    int hx[17][16], line[17];
    const int rows = y ? n + 1 : n;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < n; c++) {
            if (!x) { hx[r][c] = src[r * stride + c]; continue; }
            for (int k = 0; k <= n; k++) line[k] = src[r * stride + k];
            const int h = q_clip((q_tap(line, n, c) + rnd) >> 5);
            hx[r][c] = x == 2 ? h : (src[r * stride + c + (x == 3)] + h + up) >> 1;
        }
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) {
            int v = hx[r][c];
            if (y) {
                for (int k = 0; k <= n; k++) line[k] = hx[k][c];
                const int vv = q_clip((q_tap(line, n, r) + rnd) >> 5);
                v = y == 2 ? vv : (hx[r + (y == 3)][c] + vv + up) >> 1;
            }
            uint8_t *d = dst + r * stride + c;
            *d = (uint8_t)(kind == 2 ? (*d + v + 1) >> 1 : v);
        }
}
