/*
 * oracle/port/orc_h264.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (8-bit, 4:2:0) of the H.264 DSP tables:
 *   H264DSPContext   idct / dc / per-MB dispatch / DC dequant   libavcodec/h264idct_template.c:33-324
 *                    weight / biweight                          libavcodec/h264dsp_template.c:30-98
 *                    loop filters                               libavcodec/h264dsp_template.c:104-328
 *                    add_pixels{4,8}_clear                      libavcodec/h264addpx_template.c
 *   H264QpelContext  16 quarter-pel positions x {put,avg}       libavcodec/h264qpel_template.c:77-537
 *   H264ChromaContext bilinear 1/8-pel                          libavcodec/h264chroma_template.c:27-173
 * Written plane-wise (every quarter-pel position is an average of at most two of the planes
 * F, H, V, HV evaluated per pixel) instead of the reference's macro-generated per-size functions.
 * Pinned byte-for-byte against oracle/_ref in tests/test_oracle_h264_cpu.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../oracle_api.h"

static inline int clip_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* scan8[] (libavcodec/h264dec.h:631-645): cache position of 4x4 block i */
static int scan8_of(int i)
{
    int plane = i >> 4, k = i & 15;
    int col = 4 + (k & 1) + 2 * ((k >> 2) & 1);
    int row = 1 + ((k >> 1) & 1) + 2 * (k >> 3) + 5 * plane;
    return col + 8 * row;
}

/* ---- residual ---- */
static void idct4_add(uint8_t *dst, int16_t *b, int stride)
{
    b[0] = (int16_t)(b[0] + 32);
    for (int i = 0; i < 4; i++) {
        int z0 = b[i] + b[i + 8], z1 = b[i] - b[i + 8];
        int z2 = (b[i + 4] >> 1) - b[i + 12], z3 = b[i + 4] + (b[i + 12] >> 1);
        b[i] = (int16_t)(z0 + z3); b[i + 4] = (int16_t)(z1 + z2); b[i + 8] = (int16_t)(z1 - z2); b[i + 12] = (int16_t)(z0 - z3);
    }
    for (int i = 0; i < 4; i++) {
        const int16_t *r = b + 4 * i;
        int z0 = r[0] + r[2], z1 = r[0] - r[2], z2 = (r[1] >> 1) - r[3], z3 = r[1] + (r[3] >> 1);
        dst[i + 0 * stride] = clip_u8(dst[i + 0 * stride] + ((z0 + z3) >> 6));
        dst[i + 1 * stride] = clip_u8(dst[i + 1 * stride] + ((z1 + z2) >> 6));
        dst[i + 2 * stride] = clip_u8(dst[i + 2 * stride] + ((z1 - z2) >> 6));
        dst[i + 3 * stride] = clip_u8(dst[i + 3 * stride] + ((z0 - z3) >> 6));
    }
    memset(b, 0, 16 * sizeof(*b));
}

static void idct8_1d(const int v[8], int o[8])
{
    int a0 = v[0] + v[4], a2 = v[0] - v[4], a4 = (v[2] >> 1) - v[6], a6 = (v[6] >> 1) + v[2];
    int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    int a1 = -v[3] + v[5] - v[7] - (v[7] >> 1), a3 = v[1] + v[7] - v[3] - (v[3] >> 1);
    int a5 = -v[1] + v[7] + v[5] + (v[5] >> 1), a7 = v[3] + v[5] + v[1] + (v[1] >> 1);
    int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[7] = b0 - b7; o[1] = b2 + b5; o[6] = b2 - b5;
    o[2] = b4 + b3; o[5] = b4 - b3; o[3] = b6 + b1; o[4] = b6 - b1;
}

static void idct8_add(uint8_t *dst, int16_t *b, int stride)
{
    int v[8], o[8];
    b[0] = (int16_t)(b[0] + 32);
    for (int i = 0; i < 8; i++) {
        for (int k = 0; k < 8; k++) v[k] = b[i + 8 * k];
        idct8_1d(v, o);
        for (int k = 0; k < 8; k++) b[i + 8 * k] = (int16_t)o[k];
    }
    for (int i = 0; i < 8; i++) {
        for (int k = 0; k < 8; k++) v[k] = b[8 * i + k];
        idct8_1d(v, o);
        for (int k = 0; k < 8; k++) dst[i + k * stride] = clip_u8(dst[i + k * stride] + (o[k] >> 6));
    }
    memset(b, 0, 64 * sizeof(*b));
}

static void dc_add(uint8_t *dst, int16_t *b, int stride, int n)
{
    int dc = (b[0] + 32) >> 6;
    b[0] = 0;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) dst[y * stride + x] = clip_u8(dst[y * stride + x] + dc);
}

void orc_h264_idct(int which, uint8_t *dst, int16_t *block, int stride)
{
    if (which == 0) idct4_add(dst, block, stride);
    else if (which == 1) idct8_add(dst, block, stride);
    else dc_add(dst, block, stride, which == 2 ? 4 : 8);
}

void orc_h264_idct_mb(int which, uint8_t *dst, uint8_t **dst2, const int *bo, int16_t *block, int stride, const uint8_t *nnzc)
{
    if (which == 0 || which == 1) {
        for (int i = 0; i < 16; i++) {
            int nnz = nnzc[scan8_of(i)];
            int16_t *b = block + 16 * i;
            if (which == 0) {
                if (!nnz) continue;
                if (nnz == 1 && b[0]) dc_add(dst + bo[i], b, stride, 4); else idct4_add(dst + bo[i], b, stride);
            } else {
                if (nnz) idct4_add(dst + bo[i], b, stride); else if (b[0]) dc_add(dst + bo[i], b, stride, 4);
            }
        }
    } else if (which == 2) {
        for (int i = 0; i < 16; i += 4) {
            int nnz = nnzc[scan8_of(i)];
            int16_t *b = block + 16 * i;
            if (!nnz) continue;
            if (nnz == 1 && b[0]) dc_add(dst + bo[i], b, stride, 8); else idct8_add(dst + bo[i], b, stride);
        }
    } else {
        /* which 3: idct_add8 (4:2:0, four blocks per plane); which 4: idct_add8_422 -- the lower four blocks of a plane keep
         * their coefficients at block i but their nnz / offset entries live four places further (h264idct_template.c:216-236) */
        for (int half = 0; half < (which == 4 ? 2 : 1); half++)
            for (int j = 1; j < 3; j++)
                for (int k = 0; k < 4; k++) {
                    int i = 16 * j + 4 * half + k, e = i + 4 * half;
                    int16_t *b = block + 16 * i;
                    if (nnzc[scan8_of(e)]) idct4_add(dst2[j - 1] + bo[e], b, stride);
                    else if (b[0]) dc_add(dst2[j - 1] + bo[e], b, stride, 4);
                }
    }
}

void orc_h264_luma_dc_dequant_idct(int16_t *out, int16_t *in, int qmul)
{
    static const int xoff[4] = { 0, 32, 128, 160 };
    int t[16];
    for (int i = 0; i < 4; i++) {
        int z0 = in[4 * i] + in[4 * i + 1], z1 = in[4 * i] - in[4 * i + 1];
        int z2 = in[4 * i + 2] - in[4 * i + 3], z3 = in[4 * i + 2] + in[4 * i + 3];
        t[4 * i] = z0 + z3; t[4 * i + 1] = z0 - z3; t[4 * i + 2] = z1 - z2; t[4 * i + 3] = z1 + z2;
    }
    for (int i = 0; i < 4; i++) {
        int z0 = t[i] + t[8 + i], z1 = t[i] - t[8 + i], z2 = t[4 + i] - t[12 + i], z3 = t[4 + i] + t[12 + i];
        out[xoff[i] + 0]  = (int16_t)(((z0 + z3) * qmul + 128) >> 8);
        out[xoff[i] + 16] = (int16_t)(((z1 + z2) * qmul + 128) >> 8);
        out[xoff[i] + 64] = (int16_t)(((z1 - z2) * qmul + 128) >> 8);
        out[xoff[i] + 80] = (int16_t)(((z0 - z3) * qmul + 128) >> 8);
    }
}

void orc_h264_chroma_dc_dequant_idct(int16_t *b, int qmul)
{
    int a = b[0], bb = b[16], c = b[32], d = b[48];
    int e = a - bb; a += bb; bb = c - d; c += d;
    b[0] = (int16_t)(((a + c) * qmul) >> 7);  b[16] = (int16_t)(((e + bb) * qmul) >> 7);
    b[32] = (int16_t)(((a - c) * qmul) >> 7); b[48] = (int16_t)(((e - bb) * qmul) >> 7);
}

/* 4:2:2 chroma DC: 2 (across) x 4 (down) Hadamard, the eight DC values sit 16 coefficients apart (h264idct_template.c:277-302) */
void orc_h264_chroma422_dc_dequant_idct(int16_t *b, int qmul)
{
    int t[8];
    for (int i = 0; i < 4; i++) { t[2 * i] = b[32 * i] + b[32 * i + 16]; t[2 * i + 1] = b[32 * i] - b[32 * i + 16]; }
    for (int i = 0; i < 2; i++) {
        int z0 = t[i] + t[4 + i], z1 = t[i] - t[4 + i], z2 = t[2 + i] - t[6 + i], z3 = t[2 + i] + t[6 + i];
        b[16 * i + 0]  = (int16_t)(((z0 + z3) * qmul + 128) >> 8);
        b[16 * i + 32] = (int16_t)(((z1 + z2) * qmul + 128) >> 8);
        b[16 * i + 64] = (int16_t)(((z1 - z2) * qmul + 128) >> 8);
        b[16 * i + 96] = (int16_t)(((z0 - z3) * qmul + 128) >> 8);
    }
}

void orc_h264_add_pixels_clear(int w8, uint8_t *dst, int16_t *block, int stride)
{
    int n = w8 ? 8 : 4;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) dst[y * stride + x] = (uint8_t)(dst[y * stride + x] + block[n * y + x]);   /* wraps, no clip */
    memset(block, 0, n * n * sizeof(*block));
}

/* ---- weighted prediction ---- */
void orc_h264_weight(int widx, uint8_t *p, int stride, int height, int ld, int w, int off)
{
    int W = 16 >> widx;
    off <<= ld;
    if (ld) off += 1 << (ld - 1);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < W; x++) p[y * stride + x] = clip_u8((p[y * stride + x] * w + off) >> ld);
}

void orc_h264_biweight(int widx, uint8_t *d, uint8_t *s, int stride, int height, int ld, int wd, int ws, int off)
{
    int W = 16 >> widx;
    off = ((off + 1) | 1) << ld;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < W; x++) d[y * stride + x] = clip_u8((s[y * stride + x] * ws + d[y * stride + x] * wd + off) >> (ld + 1));
}

/* ---- deblocking: one line across an edge; px = step between samples across the edge ---- */
static void luma_line(uint8_t *q, int px, int alpha, int beta, int tc0)
{
    int p0 = q[-px], p1 = q[-2 * px], p2 = q[-3 * px], q0 = q[0], q1 = q[px], q2 = q[2 * px];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    int tc = tc0;
    if (iabs(p2 - p0) < beta) { if (tc0) q[-2 * px] = (uint8_t)(p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0)); tc++; }
    if (iabs(q2 - q0) < beta) { if (tc0) q[px] = (uint8_t)(q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0)); tc++; }
    int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    q[-px] = clip_u8(p0 + d);
    q[0] = clip_u8(q0 - d);
}

static void luma_intra_line(uint8_t *q, int px, int alpha, int beta)
{
    int p2 = q[-3 * px], p1 = q[-2 * px], p0 = q[-px], q0 = q[0], q1 = q[px], q2 = q[2 * px];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (iabs(p0 - q0) < ((alpha >> 2) + 2)) {
        if (iabs(p2 - p0) < beta) {
            int p3 = q[-4 * px];
            q[-px] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            q[-2 * px] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            q[-3 * px] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else q[-px] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (iabs(q2 - q0) < beta) {
            int q3 = q[3 * px];
            q[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            q[px] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            q[2 * px] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
        q[-px] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

static void chroma_line(uint8_t *q, int px, int alpha, int beta, int tc, int intra)
{
    int p0 = q[-px], p1 = q[-2 * px], q0 = q[0], q1 = q[px];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (intra) { q[-px] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2); }
    else {
        int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        q[-px] = clip_u8(p0 + d); q[0] = clip_u8(q0 - d);
    }
}

/* which 0..7: the 4:2:0 frame filters; 8..15: the h_ (vertical-edge) variants that only differ in the number of lines per tc0 entry
 * (h264dsp_template.c:158-163, 222-229, 272-283, 314-328): 8 luma_mbaff 9 luma_mbaff_intra 10 chroma_mbaff 11 chroma_mbaff_intra
 * 12 chroma422 13 chroma422_intra 14 chroma422_mbaff 15 chroma422_mbaff_intra */
void orc_h264_loop_filter(int which, uint8_t *pix, int stride, int alpha, int beta, const int8_t *tc0)
{
    static const uint8_t ext_lines[8] = { 8, 8, 4, 4, 16, 16, 8, 8 };
    int ext = which >= 8;
    int horiz_edge = !ext && !(which & 1);    /* v_* filters: samples across the edge are `stride` apart */
    int across = horiz_edge ? stride : 1, along = horiz_edge ? 1 : stride;
    int chroma = ext ? which >= 10 : which >= 4, intra = ext ? (which & 1) : (which & 2) != 0;
    int lines = ext ? ext_lines[which - 8] : chroma ? 8 : 16, per_group = lines / 4;
    for (int l = 0; l < lines; l++) {
        uint8_t *q = pix + l * along;
        int g = l / per_group;
        if (!chroma) {
            if (intra) luma_intra_line(q, across, alpha, beta);
            else if (tc0[g] >= 0) luma_line(q, across, alpha, beta, tc0[g]);
        } else {
            if (intra) chroma_line(q, across, alpha, beta, 0, 1);
            else if (tc0[g] > 0) chroma_line(q, across, alpha, beta, tc0[g], 0);
        }
    }
}

/* ---- luma quarter-pel: planes evaluated per pixel ---- */
static int tap6(const uint8_t *s, int step) { return (s[0] + s[step]) * 20 - (s[-step] + s[2 * step]) * 5 + (s[-2 * step] + s[3 * step]); }
static int plane_h(const uint8_t *s, ptrdiff_t st) { (void)st; return clip_u8((tap6(s, 1) + 16) >> 5); }
static int plane_v(const uint8_t *s, ptrdiff_t st) { return clip_u8((tap6(s, (int)st) + 16) >> 5); }
static int plane_hv(const uint8_t *s, ptrdiff_t st)
{
    int t[6];
    for (int k = 0; k < 6; k++) t[k] = tap6(s + (k - 2) * st, 1);
    return clip_u8(((t[2] + t[3]) * 20 - (t[1] + t[4]) * 5 + (t[0] + t[5]) + 512) >> 10);
}

void orc_h264_qpel(int avg, int sidx, int mc, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    int n = 16 >> sidx, fx = mc & 3, fy = mc >> 2;
    uint8_t out[16 * 16];
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            const uint8_t *s = src + y * stride + x;
            int a, b = -1;
            if (!fx && !fy) a = s[0];
            else if (!fy) { a = plane_h(s, stride); if (fx != 2) b = s[fx == 3]; }
            else if (!fx) { a = plane_v(s, stride); if (fy != 2) b = s[(fy == 3) * stride]; }
            else if (fx == 2 && fy == 2) a = plane_hv(s, stride);
            else if (fx == 2) { a = plane_hv(s, stride); b = plane_h(s + (fy == 3) * stride, stride); }
            else if (fy == 2) { a = plane_hv(s, stride); b = plane_v(s + (fx == 3), stride); }
            else { a = plane_h(s + (fy == 3) * stride, stride); b = plane_v(s + (fx == 3), stride); }
            out[y * n + x] = (uint8_t)(b < 0 ? a : (a + b + 1) >> 1);
        }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            dst[y * stride + x] = avg ? (uint8_t)((dst[y * stride + x] + out[y * n + x] + 1) >> 1) : out[y * n + x];
}

void orc_h264_chroma(int avg, int widx, uint8_t *dst, uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    int W = 8 >> widx, A = (8 - x) * (8 - y), B = x * (8 - y), Cc = (8 - x) * y, D = x * y;
    uint8_t out[8 * 16];
    for (int j = 0; j < h; j++)
        for (int i = 0; i < W; i++) {
            const uint8_t *s = src + j * stride + i;
            /* the reference only touches the extra column / row when its weight is non-zero (:41-63) */
            int v = A * s[0] + (B ? B * s[1] : 0) + (Cc ? Cc * s[stride] : 0) + (D ? D * s[stride + 1] : 0);
            out[j * W + i] = (uint8_t)((v + 32) >> 6);
        }
    for (int j = 0; j < h; j++)
        for (int i = 0; i < W; i++)
            dst[j * stride + i] = avg ? (uint8_t)((dst[j * stride + i] + out[j * W + i] + 1) >> 1) : out[j * W + i];
}
