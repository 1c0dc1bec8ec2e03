// libav_b200/csrc/h264_mc_block.cuh -- H.264 motion compensation of one 4x4 luma block and its two 2x2 chroma blocks by ONE thread.
//
// Replaces, per block, what mc_dir_part() runs (libavcodec/h264_mb.c:204-320): H264QpelContext.put / avg_h264_qpel_pixels_tab
// (h264qpel_template.c:77-537: the sixteen quarter-sample positions are F, H, V, J or the rounded mean of two of them, :380-531),
// H264ChromaContext.put / avg_h264_chroma_pixels_tab (h264chroma_template.c:27-173, eighth-sample bilinear) and, for blocks that
// reach outside the picture, VideoDSPContext.emulated_edge_mc (clamped coordinates).
//
// Everything is packed-byte arithmetic on registers:
//   * the 9 x 9 source patch arrives as three aligned words per row (one funnel shift each puts column -2 in byte 0);
//   * the unrounded horizontal 6-tap of a row's four samples is nine IDP.4A against constant tap words (no byte extraction, no shifts);
//   * the vertical 6-tap runs on 4 x 4 byte transposes of the column-aligned rows, again as IDP.4A against shifted tap words;
//   * the centre position J filters the int16 horizontal results vertically with IDP.2A on (row, row + 1) pairs;
//   * rounding, clipping and packing four samples is two shifts-and-adds per sample and two saturating packs per word; the mean of two
//     components and the `avg` destination update are per-byte word operations.
// The function is thread-independent (no shared memory, no warp collectives), so tests/hostsim/ runs it on the CPU against the
// compiled reference; h264_mc.cu supplies the work distribution.
#pragma once
#include "common.cuh"

namespace avb {

#ifdef AVB_HOSTSIM
inline uint32_t mc_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
inline int mc_dp4a_us(uint32_t a, uint32_t b, int c) { for (int k = 0; k < 4; k++) c += (int)((a >> (8 * k)) & 255) * (int)(int8_t)((b >> (8 * k)) & 255); return c; }
inline int mc_dp4a_uu(uint32_t a, uint32_t b, int c) { for (int k = 0; k < 4; k++) c += (int)((a >> (8 * k)) & 255) * (int)((b >> (8 * k)) & 255); return c; }
inline int mc_dp2a_lo_ss(uint32_t a, uint32_t b, int c) { return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 255) + (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 8) & 255); }
#else
__device__ __forceinline__ uint32_t mc_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
__device__ __forceinline__ int mc_dp4a_us(uint32_t a, uint32_t b, int c)
{ int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int mc_dp4a_uu(uint32_t a, uint32_t b, int c)
{ int d; asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int mc_dp2a_lo_ss(uint32_t a, uint32_t b, int c)
{ int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
#endif

// four signed tap bytes as the second IDP.4A operand
#define AVB_TAPS(a, b, c, d) ((uint32_t)((a) & 255) | (uint32_t)((b) & 255) << 8 | (uint32_t)((c) & 255) << 16 | (uint32_t)((d) & 255) << 24)

// per-byte rounded mean (a + b + 1) >> 1 of two words of four samples
__device__ __forceinline__ uint32_t mc_avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7fu); }

// 6-tap over eight consecutive samples held as (w0 = samples 0..3, w1 = samples 4..7) + sample 8 in byte 0 of w2:
// out[c] = s[c] - 5 s[c+1] + 20 s[c+2] + 20 s[c+3] - 5 s[c+4] + s[c+5], c = 0..3
// (+ RND: the rounding constant rides in the accumulator of the first dot product)
template <int RND>
__device__ __forceinline__ void mc_tap6(uint32_t w0, uint32_t w1, uint32_t w2, int (&out)[4])
{
    out[0] = mc_dp4a_us(w1, AVB_TAPS(-5, 1, 0, 0), mc_dp4a_us(w0, AVB_TAPS(1, -5, 20, 20), RND));
    out[1] = mc_dp4a_us(w1, AVB_TAPS(20, -5, 1, 0), mc_dp4a_us(w0, AVB_TAPS(0, 1, -5, 20), RND));
    out[2] = mc_dp4a_us(w1, AVB_TAPS(20, 20, -5, 1), mc_dp4a_us(w0, AVB_TAPS(0, 0, 1, -5), RND));
    out[3] = mc_dp4a_us(w2, AVB_TAPS(1, 0, 0, 0), mc_dp4a_us(w1, AVB_TAPS(-5, 20, 20, -5), mc_dp4a_us(w0, AVB_TAPS(0, 0, 0, 1), RND)));
}
// rows r0..r3 hold 4 columns each: out[c] = (r0.c, r1.c, r2.c, r3.c)
__device__ __forceinline__ void mc_transpose4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t (&out)[4])
{
    const uint32_t t0 = __byte_perm(r0, r1, 0x5140), t1 = __byte_perm(r2, r3, 0x5140), t2 = __byte_perm(r0, r1, 0x7362), t3 = __byte_perm(r2, r3, 0x7362);
    out[0] = __byte_perm(t0, t1, 0x5410); out[1] = __byte_perm(t0, t1, 0x7632); out[2] = __byte_perm(t2, t3, 0x5410); out[3] = __byte_perm(t2, t3, 0x7632);
}
template <int RND, int SH> __device__ __forceinline__ uint32_t mc_round_pack(const int (&v)[4])
{ return pack4_sat_u8((v[0] + RND) >> SH, (v[1] + RND) >> SH, (v[2] + RND) >> SH, (v[3] + RND) >> SH); }

struct McPlanes { const uint8_t *y, *cb, *cr; };

// One 4x4 luma block at (x, y) of the (stacked) destination, quarter-sample source position (mx, my) = 4 * position + vector, and its
// 2x2 chroma blocks.  [ly0, ly0 + ph) are the luma rows of the picture the block belongs to (clamping = emulated_edge_mc).
// need_h / need_v / need_j: which components any thread of the calling group needs (uniform, only saves work).
// true when both the 9 x 9 luma patch and the 3 x 3 chroma patches of the block can be fetched as aligned words inside the picture
__device__ __forceinline__ bool mc_block_inside(int pw, int ph, int ly0, int mx, int my, bool aligned)
{
    const int sx = (mx >> 2) - 2, sy = (my >> 2) - 2, cx = mx >> 3, cy = my >> 3, cy0 = ly0 >> 1;
    return aligned && sx >= 0 && (sx & ~3) + 12 <= pw && sy >= ly0 && sy + 9 <= ly0 + ph &&
           cx >= 0 && (cx & ~3) + 8 <= (pw >> 1) && cy >= cy0 && cy + 3 <= cy0 + (ph >> 1);
}

// EDGE = false: mc_block_inside() holds (word fetches); EDGE = true: every sample is fetched on its own with clamped coordinates.
// Callers keep the two kinds in separate groups of threads: a warp that mixed them would execute both fetch paths.
template <bool EDGE>
__device__ __forceinline__ void mc_block(const McPlanes &ref, uint8_t *__restrict__ dy, uint8_t *__restrict__ dcb, uint8_t *__restrict__ dcr,
                                         int ls, int uvls, int pw, int ph, int ly0, int x, int y, int mx, int my, int avg,
                                         bool need_h, bool need_v, bool need_j)
{
    const int fx = mx & 3, fy = my & 3;
    // ---- luma patch: rows -2 .. 6, words of columns -2..1, 2..5, 6..9 ----
    uint32_t W0[9], W1[9], W2[9];
    {
        const int sx = (mx >> 2) - 2, sy = (my >> 2) - 2;
        const int ax = sx & ~3;
        if (!EDGE) {
            const uint32_t sh = 8u * (uint32_t)(sx & 3);
            const uint8_t *p = ref.y + (size_t)sy * ls + ax;
#pragma unroll
            for (int r = 0; r < 9; r++) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(p + (size_t)r * ls);
                const uint32_t a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2);
                W0[r] = mc_funnel_r(a0, a1, sh); W1[r] = mc_funnel_r(a1, a2, sh); W2[r] = mc_funnel_r(a2, 0u, sh);
            }
        } else {
            int cxs[9];                                            // clamped columns (emulated_edge_mc replicates the border samples)
#pragma unroll
            for (int k = 0; k < 9; k++) cxs[k] = min(max(sx + k, 0), pw - 1);
#pragma unroll
            for (int r = 0; r < 9; r++) {                          // (fully unrolled: the patch stays in registers)
                const uint8_t *row = ref.y + (size_t)min(max(sy + r, ly0), ly0 + ph - 1) * ls;
                uint32_t w[3] = { 0, 0, 0 };
#pragma unroll
                for (int k = 0; k < 9; k++) w[k >> 2] |= (uint32_t)__ldg(row + cxs[k]) << (8 * (k & 3));
                W0[r] = w[0]; W1[r] = w[1]; W2[r] = w[2];
            }
        }
    }
    // ---- components: four words (one per output row) of four samples each ----
    uint32_t Fw[5] = { 0, 0, 0, 0, 0 }, Hp[5] = { 0, 0, 0, 0, 0 }, Vw[4] = { 0, 0, 0, 0 }, Jw[4] = { 0, 0, 0, 0 };
    const uint32_t xsh = 8u * (2u + (uint32_t)(fx == 3));            // F and V sit one column to the right for fx == 3
    uint32_t C[9];
#pragma unroll
    for (int r = 0; r < 9; r++) C[r] = mc_funnel_r(W0[r], W1[r], xsh);  // columns xoff .. xoff + 3 of every row
#pragma unroll
    for (int k = 0; k < 5; k++) Fw[k] = C[2 + k];                     // rows 0 .. 4
    if (need_h || need_j) {
        int h[9][4];
#pragma unroll
        for (int r = 0; r < 9; r++)
            if (need_j || (r >= 2 && r <= 6)) mc_tap6<0>(W0[r], W1[r], W2[r], h[r]);
#pragma unroll
        for (int k = 0; k < 5; k++) Hp[k] = mc_round_pack<16, 5>(h[2 + k]);      // rows 0 .. 4, rounded
        if (need_j) {
            // (row, row + 1) int16 pairs of every column: E[k] = rows 2k, 2k + 1; O[k] = rows 2k + 1, 2k + 2 (row index 0 = row -2)
#pragma unroll
            for (int yy = 0; yy < 4; yy++) {
                int j[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t p0 = __byte_perm((uint32_t)h[yy][c], (uint32_t)h[yy + 1][c], 0x5410);
                    const uint32_t p1 = __byte_perm((uint32_t)h[yy + 2][c], (uint32_t)h[yy + 3][c], 0x5410);
                    const uint32_t p2 = __byte_perm((uint32_t)h[yy + 4][c], (uint32_t)h[yy + 5][c], 0x5410);
                    j[c] = mc_dp2a_lo_ss(p2, AVB_TAPS(-5, 1, 0, 0), mc_dp2a_lo_ss(p1, AVB_TAPS(20, 20, 0, 0), mc_dp2a_lo_ss(p0, AVB_TAPS(1, -5, 0, 0), 0)));
                }
                Jw[yy] = mc_round_pack<512, 10>(j);
            }
        }
    }
    if (need_v) {
        uint32_t T0[4], T1[4];
        mc_transpose4(C[0], C[1], C[2], C[3], T0);
        mc_transpose4(C[4], C[5], C[6], C[7], T1);
        int v[4][4];                                                  // [column][row]
#pragma unroll
        for (int c = 0; c < 4; c++) mc_tap6<16>(T0[c], T1[c], __byte_perm(C[8], 0u, 0x4440 + c), v[c]);
#pragma unroll
        for (int yy = 0; yy < 4; yy++) {
            const int row[4] = { v[0][yy], v[1][yy], v[2][yy], v[3][yy] };
            Vw[yy] = mc_round_pack<0, 5>(row);
        }
    }
    // ---- the position picks one component or the mean of two (h264qpel_template.c:380-531) ----
    // codes: 0 F, 1 H, 2 V, 3 J, 4 none; nibble (fx + 4 fy) of each table
    const unsigned pos = (unsigned)(fx + 4 * fy);
    const int ca = (int)((0x1312333213121110ull >> (4 * pos)) & 15), cb_ = (int)((0x2120242421200404ull >> (4 * pos)) & 15);
    const int dn = fy == 3;                                           // F and H sit one row lower for fy == 3
#pragma unroll
    for (int yy = 0; yy < 4; yy++) {
        const uint32_t Fy = dn ? Fw[yy + 1] : Fw[yy], Hy = dn ? Hp[yy + 1] : Hp[yy];
        const uint32_t A = ca == 0 ? Fy : ca == 1 ? Hy : ca == 2 ? Vw[yy] : Jw[yy];
        const uint32_t B = cb_ == 0 ? Fy : cb_ == 1 ? Hy : Vw[yy];
        uint32_t v = cb_ == 4 ? A : mc_avg4(A, B);
        uint32_t *d = reinterpret_cast<uint32_t *>(dy + (size_t)(y + yy) * ls + x);
        if (avg) v = mc_avg4(*d, v);
        *d = v;
    }
    // ---- chroma: 2x2 per plane, eighth-sample bilinear ----
    {
        const int cfx = mx & 7, cfy = my & 7;
        const uint32_t wt = (uint32_t)((8 - cfx) * (8 - cfy)) | (uint32_t)(cfx * (8 - cfy)) << 8 | (uint32_t)((8 - cfx) * cfy) << 16 | (uint32_t)(cfx * cfy) << 24;
        const int cx = mx >> 3, cy = my >> 3, cw = pw >> 1, chh = ph >> 1, cy0 = ly0 >> 1, ax = cx & ~3;
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t *src = pl ? ref.cr : ref.cb;
            uint32_t w[3];
            if (!EDGE) {
                const uint32_t sh = 8u * (uint32_t)(cx & 3);
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(src + (size_t)(cy + r) * uvls + ax);
                    w[r] = mc_funnel_r(__ldg(q), __ldg(q + 1), sh);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint8_t *row = src + (size_t)min(max(cy + r, cy0), cy0 + chh - 1) * uvls;
                    w[r] = 0;
#pragma unroll
                    for (int k = 0; k < 3; k++) w[r] |= (uint32_t)__ldg(row + min(max(cx + k, 0), cw - 1)) << (8 * k);
                }
            }
            uint8_t *dp = (pl ? dcr : dcb) + (size_t)(y >> 1) * uvls + (x >> 1);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int v0 = mc_dp4a_uu(__byte_perm(w[r], w[r + 1], 0x5410), wt, 32) >> 6, v1 = mc_dp4a_uu(__byte_perm(w[r], w[r + 1], 0x6521), wt, 32) >> 6;
                uint32_t v = (uint32_t)v0 | (uint32_t)v1 << 8;
                uint16_t *d = reinterpret_cast<uint16_t *>(dp + (size_t)r * uvls);
                if (avg) v = mc_avg4(*d, v) & 0xffffu;
                *d = (uint16_t)v;
            }
        }
    }
}

}  // namespace avb
