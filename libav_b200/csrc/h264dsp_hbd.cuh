// libav_b200/csrc/h264dsp_hbd.cuh -- device-side arithmetic of the 9 / 10-bit instances of the H.264 DSP tables (BIT_DEPTH > 8 in
// libavcodec/bit_depth_template.c:49-67: uint16 samples, int32 coefficients, clipping to `bits` bits), used by the per-call slot
// kernel in slots_hbd.cu.  Same formulas as h264dsp.cuh with what the wider types change: no int16 write-back between the transform
// passes, alpha / beta / tc0 / weighting offsets scaled by 2^(bits - 8) by the CALLER of these functions (slots_hbd.cu).
//   libavcodec/h264idct_template.c:33-324, h264dsp_template.c:30-328, h264addpx_template.c:30-77,
//   h264qpel_template.c:77-537, h264chroma_template.c:27-173
// `st` is the distance between rows in SAMPLES.
#pragma once
#include "common.cuh"

namespace avb {
namespace hbd {

typedef uint16_t px;
__device__ __forceinline__ int clipb(int v, int bits) { return min(max(v, 0), (1 << bits) - 1); }
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int clip3(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ inline void idct4_add(int bits, px *dst, int32_t *b, int st)
{
    int c[16];
    for (int i = 0; i < 16; i++) c[i] = b[i];
    c[0] += 32;
    for (int i = 0; i < 4; i++) {
        int z0 = c[i] + c[i + 8], z1 = c[i] - c[i + 8], z2 = (c[i + 4] >> 1) - c[i + 12], z3 = c[i + 4] + (c[i + 12] >> 1);
        c[i] = z0 + z3; c[i + 4] = z1 + z2; c[i + 8] = z1 - z2; c[i + 12] = z0 - z3;
    }
    for (int i = 0; i < 4; i++) {
        int z0 = c[4 * i] + c[4 * i + 2], z1 = c[4 * i] - c[4 * i + 2], z2 = (c[4 * i + 1] >> 1) - c[4 * i + 3], z3 = c[4 * i + 1] + (c[4 * i + 3] >> 1);
        dst[i + 0 * st] = (px)clipb(dst[i + 0 * st] + ((z0 + z3) >> 6), bits);
        dst[i + 1 * st] = (px)clipb(dst[i + 1 * st] + ((z1 + z2) >> 6), bits);
        dst[i + 2 * st] = (px)clipb(dst[i + 2 * st] + ((z1 - z2) >> 6), bits);
        dst[i + 3 * st] = (px)clipb(dst[i + 3 * st] + ((z0 - z3) >> 6), bits);
    }
    for (int i = 0; i < 16; i++) b[i] = 0;
}
__device__ __forceinline__ void idct8_1d(const int (&v)[8], int (&o)[8])
{
    int a0 = v[0] + v[4], a2 = v[0] - v[4], a4 = (v[2] >> 1) - v[6], a6 = (v[6] >> 1) + v[2];
    int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    int a1 = -v[3] + v[5] - v[7] - (v[7] >> 1), a3 = v[1] + v[7] - v[3] - (v[3] >> 1);
    int a5 = -v[1] + v[7] + v[5] + (v[5] >> 1), a7 = v[3] + v[5] + v[1] + (v[1] >> 1);
    int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[7] = b0 - b7; o[1] = b2 + b5; o[6] = b2 - b5; o[2] = b4 + b3; o[5] = b4 - b3; o[3] = b6 + b1; o[4] = b6 - b1;
}
__device__ inline void idct8_add(int bits, px *dst, int32_t *b, int st)
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
    for (int i = 0; i < 64; i++) b[i] = 0;
}
__device__ inline void dc_add(int bits, px *dst, int32_t *b, int st, int n)
{
    const int dc = (b[0] + 32) >> 6;
    b[0] = 0;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) dst[y * st + x] = (px)clipb(dst[y * st + x] + dc, bits);
}
__device__ inline void luma_dc_dequant(int32_t *out, const int32_t *in, int qmul)
{
    const int xoff[4] = { 0, 32, 128, 160 };
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
}
__device__ inline void chroma_dc_dequant(int32_t *b, int qmul)
{
    int a = b[0], bb = b[16], c = b[32], d = b[48];
    int e = a - bb; a += bb; bb = c - d; c += d;
    b[0] = ((a + c) * qmul) >> 7; b[16] = ((e + bb) * qmul) >> 7; b[32] = ((a - c) * qmul) >> 7; b[48] = ((e - bb) * qmul) >> 7;
}
__device__ inline void chroma422_dc_dequant(int32_t *b, int qmul)
{
    int t[8];
    for (int i = 0; i < 4; i++) { t[2 * i] = b[32 * i] + b[32 * i + 16]; t[2 * i + 1] = b[32 * i] - b[32 * i + 16]; }
    for (int i = 0; i < 2; i++) {
        int z0 = t[i] + t[4 + i], z1 = t[i] - t[4 + i], z2 = t[2 + i] - t[6 + i], z3 = t[2 + i] + t[6 + i];
        b[16 * i] = ((z0 + z3) * qmul + 128) >> 8; b[16 * i + 32] = ((z1 + z2) * qmul + 128) >> 8;
        b[16 * i + 64] = ((z1 - z2) * qmul + 128) >> 8; b[16 * i + 96] = ((z0 - z3) * qmul + 128) >> 8;
    }
}

// deblocking: one line across an edge; ps = distance (samples) between samples across the edge; alpha / beta / tc already scaled
__device__ inline void luma_line(int bits, px *q, int ps, int alpha, int beta, int tc0)
{
    int p0 = q[-ps], p1 = q[-2 * ps], p2 = q[-3 * ps], q0 = q[0], q1 = q[ps], q2 = q[2 * ps];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    int tc = tc0;
    if (iabs(p2 - p0) < beta) { if (tc0) q[-2 * ps] = (px)(p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0)); tc++; }
    if (iabs(q2 - q0) < beta) { if (tc0) q[ps] = (px)(q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0)); tc++; }
    const int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    q[-ps] = (px)clipb(p0 + d, bits); q[0] = (px)clipb(q0 - d, bits);
}
__device__ inline void luma_intra_line(px *q, int ps, int alpha, int beta)
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
__device__ inline void chroma_line(int bits, px *q, int ps, int alpha, int beta, int tc, int intra)
{
    int p0 = q[-ps], p1 = q[-2 * ps], q0 = q[0], q1 = q[ps];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (intra) { q[-ps] = (px)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (px)((2 * q1 + q0 + p1 + 2) >> 2); }
    else { const int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc); q[-ps] = (px)clipb(p0 + d, bits); q[0] = (px)clipb(q0 - d, bits); }
}

// motion compensation on a staged source window (st samples between rows)
__device__ __forceinline__ int tap6(const px *s, int step) { return (s[0] + s[step]) * 20 - (s[-step] + s[2 * step]) * 5 + (s[-2 * step] + s[3 * step]); }
__device__ __forceinline__ int plane_h(int bits, const px *s) { return clipb((tap6(s, 1) + 16) >> 5, bits); }
__device__ __forceinline__ int plane_v(int bits, const px *s, int st) { return clipb((tap6(s, st) + 16) >> 5, bits); }
__device__ inline int plane_hv(int bits, const px *s, int st)
{
    int t[6];
    for (int k = 0; k < 6; k++) t[k] = tap6(s + (k - 2) * st, 1);
    return clipb(((t[2] + t[3]) * 20 - (t[1] + t[4]) * 5 + (t[0] + t[5]) + 512) >> 10, bits);
}
__device__ inline int qpel_sample(int bits, const px *s, int st, int fx, int fy)       // h264qpel_template.c:380-531
{
    int a, b = -1;
    if (!fx && !fy) a = s[0];
    else if (!fy) { a = plane_h(bits, s); if (fx != 2) b = s[fx == 3]; }
    else if (!fx) { a = plane_v(bits, s, st); if (fy != 2) b = s[(fy == 3) * st]; }
    else if (fx == 2 && fy == 2) a = plane_hv(bits, s, st);
    else if (fx == 2) { a = plane_hv(bits, s, st); b = plane_h(bits, s + (fy == 3) * st); }
    else if (fy == 2) { a = plane_hv(bits, s, st); b = plane_v(bits, s + (fx == 3), st); }
    else { a = plane_h(bits, s + (fy == 3) * st); b = plane_v(bits, s + (fx == 3), st); }
    return b < 0 ? a : (a + b + 1) >> 1;
}
__device__ inline int chroma_sample(const px *s, int st, int fx, int fy)               // zero-weight taps are not fetched
{
    const int A = (8 - fx) * (8 - fy), B = fx * (8 - fy), Cc = (8 - fx) * fy, D = fx * fy;
    int v = A * s[0];
    if (B) v += B * s[1];
    if (Cc) v += Cc * s[st];
    if (D) v += D * s[st + 1];
    return (v + 32) >> 6;
}

}  // namespace hbd
}  // namespace avb
