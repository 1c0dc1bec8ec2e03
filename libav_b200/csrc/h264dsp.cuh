// libav_b200/csrc/h264dsp.cuh -- device-side arithmetic of the H.264 DSP tables (8 bit, 4:2:0), shared by the
// batched kernels (h264dsp.cu) and the per-call slot kernel (slots.cu).  Integer semantics follow
//   libavcodec/h264idct_template.c:33-324, libavcodec/h264dsp_template.c:30-328,
//   libavcodec/h264qpel_template.c:77-537, libavcodec/h264chroma_template.c:27-173, libavcodec/h264addpx_template.c
// exactly (int16 write-back between the transform passes, read-all-then-write per deblock line, two-stage rounding
// of the quarter-pel planes).
#pragma once
#include "common.cuh"

namespace avb {

__device__ __forceinline__ int iabs_(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int clip3_(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int s16(int v) { return (int)(int16_t)v; }

// scan8[] (libavcodec/h264dec.h:631-645): position of 4x4 block i in the 8-wide non_zero_count_cache
__device__ __forceinline__ int scan8_of(int i)
{
    int plane = i >> 4, k = i & 15;
    return 4 + (k & 1) + 2 * ((k >> 2) & 1) + 8 * (1 + ((k >> 1) & 1) + 2 * (k >> 3) + 5 * plane);
}
// default frame-MB block offsets (libavcodec/h264_slice.c:486-493): 4x4 block k of a 16x16 (luma) / 8x8 (chroma) MB
__device__ __forceinline__ int blk_x(int k) { return 4 * ((k & 1) + 2 * ((k >> 2) & 1)); }
__device__ __forceinline__ int blk_y(int k) { return 4 * (((k >> 1) & 1) + 2 * (k >> 3)); }

// ---- residual: one thread transforms one block --------------------------------------------------------------
__device__ inline void h264_idct4_add(uint8_t *dst, int16_t *b, int stride)
{
    int c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = b[i];
    c[0] = s16(c[0] + 32);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int z0 = c[i] + c[i + 8], z1 = c[i] - c[i + 8], z2 = (c[i + 4] >> 1) - c[i + 12], z3 = c[i + 4] + (c[i + 12] >> 1);
        c[i] = s16(z0 + z3); c[i + 4] = s16(z1 + z2); c[i + 8] = s16(z1 - z2); c[i + 12] = s16(z0 - z3);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int z0 = c[4 * i] + c[4 * i + 2], z1 = c[4 * i] - c[4 * i + 2];
        int z2 = (c[4 * i + 1] >> 1) - c[4 * i + 3], z3 = c[4 * i + 1] + (c[4 * i + 3] >> 1);
        dst[i + 0 * stride] = (uint8_t)clip_u8(dst[i + 0 * stride] + ((z0 + z3) >> 6));
        dst[i + 1 * stride] = (uint8_t)clip_u8(dst[i + 1 * stride] + ((z1 + z2) >> 6));
        dst[i + 2 * stride] = (uint8_t)clip_u8(dst[i + 2 * stride] + ((z1 - z2) >> 6));
        dst[i + 3 * stride] = (uint8_t)clip_u8(dst[i + 3 * stride] + ((z0 - z3) >> 6));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) b[i] = 0;
}

__device__ __forceinline__ void h264_idct8_1d(const int (&v)[8], int (&o)[8])
{
    int a0 = v[0] + v[4], a2 = v[0] - v[4], a4 = (v[2] >> 1) - v[6], a6 = (v[6] >> 1) + v[2];
    int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    int a1 = -v[3] + v[5] - v[7] - (v[7] >> 1), a3 = v[1] + v[7] - v[3] - (v[3] >> 1);
    int a5 = -v[1] + v[7] + v[5] + (v[5] >> 1), a7 = v[3] + v[5] + v[1] + (v[1] >> 1);
    int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[7] = b0 - b7; o[1] = b2 + b5; o[6] = b2 - b5;
    o[2] = b4 + b3; o[5] = b4 - b3; o[3] = b6 + b1; o[4] = b6 - b1;
}

__device__ inline void h264_idct8_add(uint8_t *dst, int16_t *b, int stride)
{
    b[0] = (int16_t)(b[0] + 32);
    for (int i = 0; i < 8; i++) {
        int v[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = b[i + 8 * k];
        h264_idct8_1d(v, o);
#pragma unroll
        for (int k = 0; k < 8; k++) b[i + 8 * k] = (int16_t)o[k];
    }
    for (int i = 0; i < 8; i++) {
        int v[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = b[8 * i + k];
        h264_idct8_1d(v, o);
#pragma unroll
        for (int k = 0; k < 8; k++) dst[i + k * stride] = (uint8_t)clip_u8(dst[i + k * stride] + (o[k] >> 6));
    }
    for (int i = 0; i < 64; i++) b[i] = 0;
}

__device__ inline void h264_dc_add(uint8_t *dst, int16_t *b, int stride, int n)
{
    int dc = (b[0] + 32) >> 6;
    b[0] = 0;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) dst[y * stride + x] = (uint8_t)clip_u8(dst[y * stride + x] + dc);
}

__device__ inline void h264_luma_dc_dequant(int16_t *out, const int16_t *in, int qmul)
{
    const int xoff[4] = { 0, 32, 128, 160 };
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int z0 = in[4 * i] + in[4 * i + 1], z1 = in[4 * i] - in[4 * i + 1], z2 = in[4 * i + 2] - in[4 * i + 3], z3 = in[4 * i + 2] + in[4 * i + 3];
        t[4 * i] = z0 + z3; t[4 * i + 1] = z0 - z3; t[4 * i + 2] = z1 - z2; t[4 * i + 3] = z1 + z2;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int z0 = t[i] + t[8 + i], z1 = t[i] - t[8 + i], z2 = t[4 + i] - t[12 + i], z3 = t[4 + i] + t[12 + i];
        out[xoff[i] + 0]  = (int16_t)(((z0 + z3) * qmul + 128) >> 8);
        out[xoff[i] + 16] = (int16_t)(((z1 + z2) * qmul + 128) >> 8);
        out[xoff[i] + 64] = (int16_t)(((z1 - z2) * qmul + 128) >> 8);
        out[xoff[i] + 80] = (int16_t)(((z0 - z3) * qmul + 128) >> 8);
    }
}

__device__ inline void h264_chroma_dc_dequant(int16_t *b, int qmul)
{
    int a = b[0], bb = b[16], c = b[32], d = b[48];
    int e = a - bb; a += bb; bb = c - d; c += d;
    b[0] = (int16_t)(((a + c) * qmul) >> 7);  b[16] = (int16_t)(((e + bb) * qmul) >> 7);
    b[32] = (int16_t)(((a - c) * qmul) >> 7); b[48] = (int16_t)(((e - bb) * qmul) >> 7);
}

// 4:2:2 chroma DC (h264idct_template.c:277-302): 2 (across) x 4 (down) Hadamard over the eight DC values, which sit 16 coefficients
// apart in the macroblock's coefficient array; (v * qmul + 128) >> 8 like the luma DC
__device__ inline void h264_chroma422_dc_dequant(int16_t *b, int qmul)
{
    int t[8];
#pragma unroll
    for (int i = 0; i < 4; i++) { t[2 * i] = b[32 * i] + b[32 * i + 16]; t[2 * i + 1] = b[32 * i] - b[32 * i + 16]; }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int z0 = t[i] + t[4 + i], z1 = t[i] - t[4 + i], z2 = t[2 + i] - t[6 + i], z3 = t[2 + i] + t[6 + i];
        b[16 * i + 0]  = (int16_t)(((z0 + z3) * qmul + 128) >> 8);
        b[16 * i + 32] = (int16_t)(((z1 + z2) * qmul + 128) >> 8);
        b[16 * i + 64] = (int16_t)(((z1 - z2) * qmul + 128) >> 8);
        b[16 * i + 96] = (int16_t)(((z0 - z3) * qmul + 128) >> 8);
    }
}

// ---- deblocking: one line across an edge; `px` = distance between samples across the edge -------------------
__device__ inline void h264_luma_line(uint8_t *q, int px, int alpha, int beta, int tc0)
{
    int p0 = q[-px], p1 = q[-2 * px], p2 = q[-3 * px], q0 = q[0], q1 = q[px], q2 = q[2 * px];
    if (iabs_(p0 - q0) >= alpha || iabs_(p1 - p0) >= beta || iabs_(q1 - q0) >= beta) return;
    int tc = tc0;
    if (iabs_(p2 - p0) < beta) { if (tc0) q[-2 * px] = (uint8_t)(p1 + clip3_(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0)); tc++; }
    if (iabs_(q2 - q0) < beta) { if (tc0) q[px] = (uint8_t)(q1 + clip3_(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0)); tc++; }
    int d = clip3_((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    q[-px] = (uint8_t)clip_u8(p0 + d);
    q[0] = (uint8_t)clip_u8(q0 - d);
}

__device__ inline void h264_luma_intra_line(uint8_t *q, int px, int alpha, int beta)
{
    int p2 = q[-3 * px], p1 = q[-2 * px], p0 = q[-px], q0 = q[0], q1 = q[px], q2 = q[2 * px];
    if (iabs_(p0 - q0) >= alpha || iabs_(p1 - p0) >= beta || iabs_(q1 - q0) >= beta) return;
    if (iabs_(p0 - q0) < ((alpha >> 2) + 2)) {
        if (iabs_(p2 - p0) < beta) {
            int p3 = q[-4 * px];
            q[-px] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            q[-2 * px] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            q[-3 * px] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else q[-px] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (iabs_(q2 - q0) < beta) {
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

__device__ inline void h264_chroma_line(uint8_t *q, int px, int alpha, int beta, int tc, int intra)
{
    int p0 = q[-px], p1 = q[-2 * px], q0 = q[0], q1 = q[px];
    if (iabs_(p0 - q0) >= alpha || iabs_(p1 - p0) >= beta || iabs_(q1 - q0) >= beta) return;
    if (intra) { q[-px] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2); }
    else {
        int d = clip3_((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        q[-px] = (uint8_t)clip_u8(p0 + d); q[0] = (uint8_t)clip_u8(q0 - d);
    }
}

// ---- motion compensation: every quarter-pel position is one of / the rounded mean of two of F, H, V, HV -------
// A sample fetcher `S(x, y)` returns the (edge-clamped) reference sample; all planes are evaluated per pixel.
template <class Fetch> __device__ __forceinline__ int qpel_h(const Fetch &S, int x, int y)
{ return clip_u8(((S(x, y) + S(x + 1, y)) * 20 - (S(x - 1, y) + S(x + 2, y)) * 5 + (S(x - 2, y) + S(x + 3, y)) + 16) >> 5); }
template <class Fetch> __device__ __forceinline__ int qpel_v(const Fetch &S, int x, int y)
{ return clip_u8(((S(x, y) + S(x, y + 1)) * 20 - (S(x, y - 1) + S(x, y + 2)) * 5 + (S(x, y - 2) + S(x, y + 3)) + 16) >> 5); }
template <class Fetch> __device__ __forceinline__ int qpel_hraw(const Fetch &S, int x, int y)
{ return (S(x, y) + S(x + 1, y)) * 20 - (S(x - 1, y) + S(x + 2, y)) * 5 + (S(x - 2, y) + S(x + 3, y)); }
template <class Fetch> __device__ __forceinline__ int qpel_hv(const Fetch &S, int x, int y)
{
    int t0 = qpel_hraw(S, x, y - 2), t1 = qpel_hraw(S, x, y - 1), t2 = qpel_hraw(S, x, y);
    int t3 = qpel_hraw(S, x, y + 1), t4 = qpel_hraw(S, x, y + 2), t5 = qpel_hraw(S, x, y + 3);
    return clip_u8(((t2 + t3) * 20 - (t1 + t4) * 5 + (t0 + t5) + 512) >> 10);
}
// fx, fy in 0..3 (mc = fx + 4 * fy), h264qpel_template.c:380-531
template <class Fetch> __device__ __forceinline__ int qpel_sample(const Fetch &S, int x, int y, int fx, int fy)
{
    int a, b = -1;
    if (!fx && !fy) a = S(x, y);
    else if (!fy) { a = qpel_h(S, x, y); if (fx != 2) b = S(x + (fx == 3), y); }
    else if (!fx) { a = qpel_v(S, x, y); if (fy != 2) b = S(x, y + (fy == 3)); }
    else if (fx == 2 && fy == 2) a = qpel_hv(S, x, y);
    else if (fx == 2) { a = qpel_hv(S, x, y); b = qpel_h(S, x, y + (fy == 3)); }
    else if (fy == 2) { a = qpel_hv(S, x, y); b = qpel_v(S, x + (fx == 3), y); }
    else { a = qpel_h(S, x, y + (fy == 3)); b = qpel_v(S, x + (fx == 3), y); }
    return b < 0 ? a : (a + b + 1) >> 1;
}
// 1/8-pel bilinear chroma, h264chroma_template.c:27-170.  Zero-weight taps are not fetched (the reference never
// touches the extra row / column then).
template <class Fetch> __device__ __forceinline__ int chroma_sample(const Fetch &S, int x, int y, int fx, int fy)
{
    int A = (8 - fx) * (8 - fy), B = fx * (8 - fy), Cc = (8 - fx) * fy, D = fx * fy;
    int v = A * S(x, y);
    if (B) v += B * S(x + 1, y);
    if (Cc) v += Cc * S(x, y + 1);
    if (D) v += D * S(x + 1, y + 1);
    return (v + 32) >> 6;
}

}  // namespace avb
