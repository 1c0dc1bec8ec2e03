/*
 * oracle/port/orc_h264_hbd.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the 9 / 10-bit instances of the H.264 DSP tables (BIT_DEPTH > 8 in libavcodec/bit_depth_template.c:49-67:
 * pixel = uint16_t, dctcoef = int32_t, clipping to BIT_DEPTH bits):
 *   H264DSPContext    libavcodec/h264idct_template.c:33-324, h264dsp_template.c:30-328, h264addpx_template.c:30-77
 *   H264QpelContext   libavcodec/h264qpel_template.c:77-537     H264ChromaContext  libavcodec/h264chroma_template.c:27-173
 * What differs from the 8-bit instance (orc_h264.c): no int16 truncation between the transform passes (the coefficients are int32),
 * alpha / beta / tc0 and the weighting offset scaled by 2^(bits - 8), every clip to (1 << bits) - 1.  The quarter-pel centre position
 * is stated on the unrounded sums (the reference's `pad` only keeps its int16 temporaries in range, h264qpel_template.c:38-42).
 * `stride` is in BYTES everywhere, like the reference's.  Pinned against oracle/_ref in tests/test_oracle_h264_hbd_cpu.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../oracle_api.h"

typedef uint16_t px;
static inline int clipb(int v, int bits) { int m = (1 << bits) - 1; return v < 0 ? 0 : v > m ? m : v; }
static inline int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int iabs(int v) { return v < 0 ? -v : v; }
static int scan8_of(int i)
{
    int plane = i >> 4, k = i & 15;
    return 4 + (k & 1) + 2 * ((k >> 2) & 1) + 8 * (1 + ((k >> 1) & 1) + 2 * (k >> 3) + 5 * plane);
}

/* ---- residual ---- */
static void idct4_add(int bits, px *dst, int32_t *b, int st)
{
    b[0] += 32;
    for (int i = 0; i < 4; i++) {
        int z0 = b[i] + b[i + 8], z1 = b[i] - b[i + 8], z2 = (b[i + 4] >> 1) - b[i + 12], z3 = b[i + 4] + (b[i + 12] >> 1);
        b[i] = z0 + z3; b[i + 4] = z1 + z2; b[i + 8] = z1 - z2; b[i + 12] = z0 - z3;
    }
    for (int i = 0; i < 4; i++) {
        int z0 = b[4 * i] + b[4 * i + 2], z1 = b[4 * i] - b[4 * i + 2], z2 = (b[4 * i + 1] >> 1) - b[4 * i + 3], z3 = b[4 * i + 1] + (b[4 * i + 3] >> 1);
        dst[i + 0 * st] = (px)clipb(dst[i + 0 * st] + ((z0 + z3) >> 6), bits);
        dst[i + 1 * st] = (px)clipb(dst[i + 1 * st] + ((z1 + z2) >> 6), bits);
        dst[i + 2 * st] = (px)clipb(dst[i + 2 * st] + ((z1 - z2) >> 6), bits);
        dst[i + 3 * st] = (px)clipb(dst[i + 3 * st] + ((z0 - z3) >> 6), bits);
    }
    memset(b, 0, 16 * sizeof(*b));
}
static void idct8_1d(const int *v, int *o)
{
    int a0 = v[0] + v[4], a2 = v[0] - v[4], a4 = (v[2] >> 1) - v[6], a6 = (v[6] >> 1) + v[2];
    int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    int a1 = -v[3] + v[5] - v[7] - (v[7] >> 1), a3 = v[1] + v[7] - v[3] - (v[3] >> 1);
    int a5 = -v[1] + v[7] + v[5] + (v[5] >> 1), a7 = v[3] + v[5] + v[1] + (v[1] >> 1);
    int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[7] = b0 - b7; o[1] = b2 + b5; o[6] = b2 - b5; o[2] = b4 + b3; o[5] = b4 - b3; o[3] = b6 + b1; o[4] = b6 - b1;
}
static void idct8_add(int bits, px *dst, int32_t *b, int st)
{
    b[0] += 32;
    for (int i = 0; i < 8; i++) {
        int v[8], o[8];
        for (int k = 0; k < 8; k++) v[k] = b[i + 8 * k];
        idct8_1d(v, o);
        for (int k = 0; k < 8; k++) b[i + 8 * k] = o[k];
    }
    for (int i = 0; i < 8; i++) {
        int v[8], o[8];
        for (int k = 0; k < 8; k++) v[k] = b[8 * i + k];
        idct8_1d(v, o);
        for (int k = 0; k < 8; k++) dst[i + k * st] = (px)clipb(dst[i + k * st] + (o[k] >> 6), bits);
    }
    memset(b, 0, 64 * sizeof(*b));
}
static void dc_add(int bits, px *dst, int32_t *b, int st, int n)
{
    int dc = (b[0] + 32) >> 6;
    b[0] = 0;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) dst[y * st + x] = (px)clipb(dst[y * st + x] + dc, bits);
}

void orc_h264_hbd_idct(int bits, int which, uint8_t *dst, int32_t *block, int stride)
{
    px *d = (px *)dst; int st = stride / 2;
    if (which == 0) idct4_add(bits, d, block, st);
    else if (which == 1) idct8_add(bits, d, block, st);
    else dc_add(bits, d, block, st, which == 2 ? 4 : 8);
}

void orc_h264_hbd_idct_mb(int bits, int which, uint8_t *dst, uint8_t **dst2, const int *bo, int32_t *block, int stride, const uint8_t *nnzc)
{
    const int st = stride / 2;
    if (which <= 1) {
        for (int i = 0; i < 16; i++) {
            int nnz = nnzc[scan8_of(i)];
            int32_t *b = block + 16 * i;
            px *d = (px *)(dst + bo[i]);
            if (which == 0) { if (nnz) { if (nnz == 1 && b[0]) dc_add(bits, d, b, st, 4); else idct4_add(bits, d, b, st); } }
            else { if (nnz) idct4_add(bits, d, b, st); else if (b[0]) dc_add(bits, d, b, st, 4); }
        }
    } else if (which == 2) {
        for (int i = 0; i < 16; i += 4) {
            int nnz = nnzc[scan8_of(i)];
            int32_t *b = block + 16 * i;
            px *d = (px *)(dst + bo[i]);
            if (nnz) { if (nnz == 1 && b[0]) dc_add(bits, d, b, st, 8); else idct8_add(bits, d, b, st); }
        }
    } else {
        for (int half = 0; half < (which == 4 ? 2 : 1); half++)
            for (int j = 1; j < 3; j++)
                for (int k = 0; k < 4; k++) {
                    int i = 16 * j + 4 * half + k, e = i + 4 * half;
                    int32_t *b = block + 16 * i;
                    px *d = (px *)(dst2[j - 1] + bo[e]);
                    if (nnzc[scan8_of(e)]) idct4_add(bits, d, b, st); else if (b[0]) dc_add(bits, d, b, st, 4);
                }
    }
}

/* kind 0 luma (16 values in, scattered 16 coefficients apart), 1 chroma 4:2:0 (in place), 2 chroma 4:2:2 (in place) */
void orc_h264_hbd_dc_dequant(int bits, int kind, int32_t *out, int32_t *in, int qmul)
{
    (void)bits;
    if (kind == 0) {
        static const int xoff[4] = { 0, 32, 128, 160 };
        int t[16];
        for (int i = 0; i < 4; i++) {
            int z0 = in[4 * i] + in[4 * i + 1], z1 = in[4 * i] - in[4 * i + 1], z2 = in[4 * i + 2] - in[4 * i + 3], z3 = in[4 * i + 2] + in[4 * i + 3];
            t[4 * i] = z0 + z3; t[4 * i + 1] = z0 - z3; t[4 * i + 2] = z1 - z2; t[4 * i + 3] = z1 + z2;
        }
        for (int i = 0; i < 4; i++) {
            int z0 = t[i] + t[8 + i], z1 = t[i] - t[8 + i], z2 = t[4 + i] - t[12 + i], z3 = t[4 + i] + t[12 + i];
            out[xoff[i] + 0] = ((z0 + z3) * qmul + 128) >> 8; out[xoff[i] + 16] = ((z1 + z2) * qmul + 128) >> 8;
            out[xoff[i] + 64] = ((z1 - z2) * qmul + 128) >> 8; out[xoff[i] + 80] = ((z0 - z3) * qmul + 128) >> 8;
        }
    } else if (kind == 1) {
        int32_t *b = out;
        int a = b[0], bb = b[16], c = b[32], d = b[48];
        int e = a - bb; a += bb; bb = c - d; c += d;
        b[0] = ((a + c) * qmul) >> 7; b[16] = ((e + bb) * qmul) >> 7; b[32] = ((a - c) * qmul) >> 7; b[48] = ((e - bb) * qmul) >> 7;
    } else {
        int32_t *b = out;
        int t[8];
        for (int i = 0; i < 4; i++) { t[2 * i] = b[32 * i] + b[32 * i + 16]; t[2 * i + 1] = b[32 * i] - b[32 * i + 16]; }
        for (int i = 0; i < 2; i++) {
            int z0 = t[i] + t[4 + i], z1 = t[i] - t[4 + i], z2 = t[2 + i] - t[6 + i], z3 = t[2 + i] + t[6 + i];
            b[16 * i] = ((z0 + z3) * qmul + 128) >> 8; b[16 * i + 32] = ((z1 + z2) * qmul + 128) >> 8;
            b[16 * i + 64] = ((z1 - z2) * qmul + 128) >> 8; b[16 * i + 96] = ((z0 - z3) * qmul + 128) >> 8;
        }
    }
}

void orc_h264_hbd_add_pixels_clear(int bits, int w8, uint8_t *dst, int32_t *block, int stride)
{
    (void)bits;
    px *d = (px *)dst; int st = stride / 2, n = w8 ? 8 : 4;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) d[y * st + x] = (px)(d[y * st + x] + block[n * y + x]);      /* wraps in 16 bits, no clip */
    memset(block, 0, (size_t)n * n * sizeof(*block));
}

/* ---- weighted prediction ---- */
void orc_h264_hbd_weight(int bits, int widx, uint8_t *p, int stride, int height, int ld, int w, int off)
{
    px *b = (px *)p; int st = stride / 2, W = 16 >> widx;
    off <<= ld + (bits - 8);
    if (ld) off += 1 << (ld - 1);
    for (int y = 0; y < height; y++) for (int x = 0; x < W; x++) b[y * st + x] = (px)clipb((b[y * st + x] * w + off) >> ld, bits);
}
void orc_h264_hbd_biweight(int bits, int widx, uint8_t *dp, uint8_t *sp, int stride, int height, int ld, int wd, int ws, int off)
{
    px *d = (px *)dp, *s = (px *)sp; int st = stride / 2, W = 16 >> widx;
    off <<= bits - 8;
    off = ((off + 1) | 1) << ld;
    for (int y = 0; y < height; y++) for (int x = 0; x < W; x++) d[y * st + x] = (px)clipb((s[y * st + x] * ws + d[y * st + x] * wd + off) >> (ld + 1), bits);
}

/* ---- deblocking: one line across an edge; ps = distance (in samples) between samples across the edge ---- */
static void luma_line(int bits, px *q, int ps, int alpha, int beta, int tc0)
{
    int p0 = q[-ps], p1 = q[-2 * ps], p2 = q[-3 * ps], q0 = q[0], q1 = q[ps], q2 = q[2 * ps];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    int tc = tc0;
    if (iabs(p2 - p0) < beta) { if (tc0) q[-2 * ps] = (px)(p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0)); tc++; }
    if (iabs(q2 - q0) < beta) { if (tc0) q[ps] = (px)(q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0)); tc++; }
    int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    q[-ps] = (px)clipb(p0 + d, bits); q[0] = (px)clipb(q0 - d, bits);
}
static void luma_intra_line(px *q, int ps, int alpha, int beta)
{
    int p2 = q[-3 * ps], p1 = q[-2 * ps], p0 = q[-ps], q0 = q[0], q1 = q[ps], q2 = q[2 * ps];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (iabs(p0 - q0) < ((alpha >> 2) + 2)) {
        if (iabs(p2 - p0) < beta) {
            int p3 = q[-4 * ps];
            q[-ps] = (px)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3); q[-2 * ps] = (px)((p2 + p1 + p0 + q0 + 2) >> 2); q[-3 * ps] = (px)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else q[-ps] = (px)((2 * p1 + p0 + q1 + 2) >> 2);
        if (iabs(q2 - q0) < beta) {
            int q3 = q[3 * ps];
            q[0] = (px)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3); q[ps] = (px)((p0 + q0 + q1 + q2 + 2) >> 2); q[2 * ps] = (px)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else q[0] = (px)((2 * q1 + q0 + p1 + 2) >> 2);
    } else { q[-ps] = (px)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (px)((2 * q1 + q0 + p1 + 2) >> 2); }
}
static void chroma_line(int bits, px *q, int ps, int alpha, int beta, int tc, int intra)
{
    int p0 = q[-ps], p1 = q[-2 * ps], q0 = q[0], q1 = q[ps];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (intra) { q[-ps] = (px)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (px)((2 * q1 + q0 + p1 + 2) >> 2); }
    else { int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc); q[-ps] = (px)clipb(p0 + d, bits); q[0] = (px)clipb(q0 - d, bits); }
}
/* which: the numbering of orc_h264_loop_filter (0..15) */
void orc_h264_hbd_loop_filter(int bits, int which, uint8_t *pix, int stride, int alpha, int beta, const int8_t *tc0)
{
    static const uint8_t ext_lines[8] = { 8, 8, 4, 4, 16, 16, 8, 8 };
    const int ext = which >= 8, st = stride / 2, sh = bits - 8;
    const int horiz_edge = !ext && !(which & 1), across = horiz_edge ? st : 1, along = horiz_edge ? 1 : st;
    const int chroma = ext ? which >= 10 : which >= 4, intra = ext ? (which & 1) : (which & 2) != 0;
    const int lines = ext ? ext_lines[which - 8] : chroma ? 8 : 16, per_group = lines / 4;
    alpha <<= sh; beta <<= sh;
    for (int l = 0; l < lines; l++) {
        px *q = (px *)pix + l * along;
        const int g = l / per_group;
        if (!chroma) {
            if (intra) luma_intra_line(q, across, alpha, beta);
            else if (tc0[g] >= 0) luma_line(bits, q, across, alpha, beta, tc0[g] << sh);       /* tc_orig = tc0[i] << (BIT_DEPTH - 8), h264dsp_template.c:113 */
        } else {
            const int tc = intra ? 0 : ((tc0[g] - 1) << sh) + 1;                                 /* :240 */
            if (intra) chroma_line(bits, q, across, alpha, beta, 0, 1);
            else if (tc > 0) chroma_line(bits, q, across, alpha, beta, tc, 0);
        }
    }
}

/* ---- luma quarter-pel, chroma eighth-pel ---- */
static int tap6(const px *s, int step) { return (s[0] + s[step]) * 20 - (s[-step] + s[2 * step]) * 5 + (s[-2 * step] + s[3 * step]); }
static int plane_h(int bits, const px *s) { return clipb((tap6(s, 1) + 16) >> 5, bits); }
static int plane_v(int bits, const px *s, int st) { return clipb((tap6(s, st) + 16) >> 5, bits); }
static int plane_hv(int bits, const px *s, int st)
{
    int t[6];
    for (int k = 0; k < 6; k++) t[k] = tap6(s + (k - 2) * st, 1);
    return clipb(((t[2] + t[3]) * 20 - (t[1] + t[4]) * 5 + (t[0] + t[5]) + 512) >> 10, bits);
}
void orc_h264_hbd_qpel(int bits, int avg, int sidx, int mc, uint8_t *dstp, const uint8_t *srcp, ptrdiff_t stride)
{
    const int n = 16 >> sidx, fx = mc & 3, fy = mc >> 2, st = (int)(stride / 2);
    px *dst = (px *)dstp; const px *src = (const px *)srcp;
    px out[16 * 16];
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            const px *s = src + y * st + x;
            int a, b = -1;
            if (!fx && !fy) a = s[0];
            else if (!fy) { a = plane_h(bits, s); if (fx != 2) b = s[fx == 3]; }
            else if (!fx) { a = plane_v(bits, s, st); if (fy != 2) b = s[(fy == 3) * st]; }
            else if (fx == 2 && fy == 2) a = plane_hv(bits, s, st);
            else if (fx == 2) { a = plane_hv(bits, s, st); b = plane_h(bits, s + (fy == 3) * st); }
            else if (fy == 2) { a = plane_hv(bits, s, st); b = plane_v(bits, s + (fx == 3), st); }
            else { a = plane_h(bits, s + (fy == 3) * st); b = plane_v(bits, s + (fx == 3), st); }
            out[y * n + x] = (px)(b < 0 ? a : (a + b + 1) >> 1);
        }
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) dst[y * st + x] = avg ? (px)((dst[y * st + x] + out[y * n + x] + 1) >> 1) : out[y * n + x];
}
void orc_h264_hbd_chroma(int bits, int avg, int widx, uint8_t *dstp, uint8_t *srcp, ptrdiff_t stride, int h, int x, int y)
{
    (void)bits;
    const int W = 8 >> widx, A = (8 - x) * (8 - y), B = x * (8 - y), Cc = (8 - x) * y, D = x * y, st = (int)(stride / 2);
    px *dst = (px *)dstp; const px *src = (const px *)srcp;
    px out[8 * 16];
    for (int j = 0; j < h; j++)
        for (int i = 0; i < W; i++) {
            const px *s = src + j * st + i;
            int v = A * s[0] + (B ? B * s[1] : 0) + (Cc ? Cc * s[st] : 0) + (D ? D * s[st + 1] : 0);
            out[j * W + i] = (px)((v + 32) >> 6);
        }
    for (int j = 0; j < h; j++) for (int i = 0; i < W; i++) dst[j * st + i] = avg ? (px)((dst[j * st + i] + out[j * W + i] + 1) >> 1) : out[j * W + i];
}
