// libav_b200/csrc/h264_residual.cu -- batched H.264 residual add (config 3's second stage), register-resident.
//
// Replaces the dispatchers h264_idct_add16 / h264_idct_add16intra / h264_idct8_add4 / h264_idct_add8
// (libavcodec/h264idct_template.c:174-214) over the transforms ff_h264_idct_add / idct8_add / idct_dc_add / idct8_dc_add (:33-172)
// for every macroblock of a batch.  One warp per macroblock, one lane per block exactly as the C loops iterate: lanes 0-15 the luma
// 4x4 blocks (lanes 0-3 the 8x8 blocks of a transform-8x8 macroblock), lanes 16-23 the cb / cr 4x4 blocks.  A lane reads the
// non_zero_count_cache byte of its block first and touches memory only if the C code would: its 32 (128) bytes of coefficients as
// 16-byte loads straight into registers (consecutive lanes = consecutive blocks = one contiguous 512-byte request), its 4 x 4 pixels as
// four word loads, both passes of the transform in registers with the int16 truncation the reference stores between them, the
// consumed coefficients zeroed with 16-byte stores (a DC-only block clears its DC alone).  Blocks the C code skips cost one byte load:
// half of the synthetic pictures' blocks, most of a real stream's.  Lanes never communicate: tests/hostsim/ runs the kernel on the CPU.  (The first version staged every macroblock's 1536 + 384 bytes
// through shared memory and wrote all of it back: 0.56 ms per 60 pictures against a 0.31 ms traffic floor.)
#include "h264dsp.cuh"
#include "../../include/avdsp_b200.h"

namespace avb {

namespace {

__device__ __forceinline__ uint32_t add_clip4(uint32_t px, int a0, int a1, int a2, int a3)
{
    return pack4_sat_u8((int)(px & 255u) + a0, (int)((px >> 8) & 255u) + a1, (int)((px >> 16) & 255u) + a2, (int)(px >> 24) + a3);
}

// ff_h264_idct_add_8_c: cw = the block's 16 coefficients (8 words), px = its four pixel rows
__device__ __forceinline__ void idct4_add_regs(const uint4 &c0, const uint4 &c1, uint32_t (&px)[4])
{
    const uint32_t cw[8] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w };
    int c[16];
#pragma unroll
    for (int k = 0; k < 8; k++) { c[2 * k] = lo16s(cw[k]); c[2 * k + 1] = hi16s(cw[k]); }
    c[0] = s16(c[0] + 32);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int z0 = c[i] + c[i + 8], z1 = c[i] - c[i + 8], z2 = (c[i + 4] >> 1) - c[i + 12], z3 = c[i + 4] + (c[i + 12] >> 1);
        c[i] = s16(z0 + z3); c[i + 4] = s16(z1 + z2); c[i + 8] = s16(z1 - z2); c[i + 12] = s16(z0 - z3);
    }
    int add[4][4];                                                  // [row y][column x]
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int z0 = c[4 * i] + c[4 * i + 2], z1 = c[4 * i] - c[4 * i + 2];
        const int z2 = (c[4 * i + 1] >> 1) - c[4 * i + 3], z3 = c[4 * i + 1] + (c[4 * i + 3] >> 1);
        add[0][i] = (z0 + z3) >> 6; add[1][i] = (z1 + z2) >> 6; add[2][i] = (z1 - z2) >> 6; add[3][i] = (z0 - z3) >> 6;
    }
#pragma unroll
    for (int y = 0; y < 4; y++) px[y] = add_clip4(px[y], add[y][0], add[y][1], add[y][2], add[y][3]);
}

// a row of four pixels; records may place a macroblock at any byte offset (`al` = its offsets are multiples of 4, uniform per warp)
__device__ __forceinline__ uint32_t ld_px(const uint8_t *p, bool al)
{ return al ? *reinterpret_cast<const uint32_t *>(p) : (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
__device__ __forceinline__ void st_px(uint8_t *p, uint32_t v, bool al)
{
    if (al) *reinterpret_cast<uint32_t *>(p) = v;
    else { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
}

__global__ void __launch_bounds__(128)
h264_residual_kernel_v2(const FFH264ResidualMB *__restrict__ mbs, size_t n, int16_t *__restrict__ coeffs, size_t coeff_stride,
                        const uint8_t *__restrict__ nnzc, uint8_t *__restrict__ luma, uint8_t *__restrict__ cb,
                        uint8_t *__restrict__ cr, int ls, int uvls)
{
    const int lane = threadIdx.x & 31;
    const size_t mb = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (mb >= n) return;
    const uint2 rw = __ldg(reinterpret_cast<const uint2 *>(mbs + mb));       // luma_off, chroma_off
    const uint32_t rm = __ldg(reinterpret_cast<const uint32_t *>(mbs + mb) + 2);
    const int mode = rm & 255, chroma = (rm >> 8) & 255;
    const bool al = !((rw.x | rw.y) & 3u);
    int16_t *gc = coeffs + mb * coeff_stride;
    const uint8_t *nz = nnzc + mb * 120;

    if (lane < 16 && mode <= 1) {                                   // h264_idct_add16 (:174-183) / h264_idct_add16intra (:185-191)
        const int i = lane, nnz = __ldg(nz + scan8_of(i));
        int16_t *b = gc + 16 * i;
        uint8_t *d = luma + rw.x + blk_x(i) + (size_t)blk_y(i) * ls;
        int kind = 0;                                               // 1 full transform, 2 DC only
        uint32_t w0 = 0;
        if (nnz) { w0 = *reinterpret_cast<const uint32_t *>(b); kind = (mode == 0 && nnz == 1 && (w0 & 0xffffu)) ? 2 : 1; }
        else if (mode == 1) { w0 = *reinterpret_cast<const uint32_t *>(b); kind = (w0 & 0xffffu) ? 2 : 0; }
        if (kind) {
            uint32_t px[4];
#pragma unroll
            for (int y = 0; y < 4; y++) px[y] = ld_px(d + (size_t)y * ls, al);
            if (kind == 1) {
                const uint4 c0 = *reinterpret_cast<const uint4 *>(b), c1 = *reinterpret_cast<const uint4 *>(b + 8);
                idct4_add_regs(c0, c1, px);
                *reinterpret_cast<uint4 *>(b) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4 *>(b + 8) = make_uint4(0, 0, 0, 0);
            } else {
                const int dc = (lo16s(w0) + 32) >> 6;
#pragma unroll
                for (int y = 0; y < 4; y++) px[y] = add_clip4(px[y], dc, dc, dc, dc);
                b[0] = 0;
            }
#pragma unroll
            for (int y = 0; y < 4; y++) st_px(d + (size_t)y * ls, px[y], al);
        }
    } else if (lane < 4 && mode == 2) {                             // h264_idct8_add4 (:193-202): blocks 0, 4, 8, 12
        const int i = 4 * lane, nnz = __ldg(nz + scan8_of(i));
        if (nnz) {
            int16_t *b = gc + 16 * i;
            uint8_t *d = luma + rw.x + blk_x(i) + (size_t)blk_y(i) * ls;
            uint32_t px[8][2];
#pragma unroll
            for (int y = 0; y < 8; y++) { px[y][0] = ld_px(d + (size_t)y * ls, al); px[y][1] = ld_px(d + (size_t)y * ls + 4, al); }
            const int b0 = b[0];
            if (nnz == 1 && b0) {                                   // ff_h264_idct8_dc_add
                const int dc = (b0 + 32) >> 6;
#pragma unroll
                for (int y = 0; y < 8; y++) { px[y][0] = add_clip4(px[y][0], dc, dc, dc, dc); px[y][1] = add_clip4(px[y][1], dc, dc, dc, dc); }
                b[0] = 0;
            } else {                                                // ff_h264_idct8_add: columns, int16 write-back, rows
                int c[64];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(b + 8 * k);
                    c[8 * k] = lo16s(q.x); c[8 * k + 1] = hi16s(q.x); c[8 * k + 2] = lo16s(q.y); c[8 * k + 3] = hi16s(q.y);
                    c[8 * k + 4] = lo16s(q.z); c[8 * k + 5] = hi16s(q.z); c[8 * k + 6] = lo16s(q.w); c[8 * k + 7] = hi16s(q.w);
                }
                c[0] = s16(c[0] + 32);
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    int v[8], o[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = c[x + 8 * k];
                    h264_idct8_1d(v, o);
#pragma unroll
                    for (int k = 0; k < 8; k++) c[x + 8 * k] = s16(o[k]);
                }
#pragma unroll
                for (int x = 0; x < 8; x++) {                       // row x of the coefficient block -> pixel COLUMN x
                    int v[8], o[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = c[8 * x + k];
                    h264_idct8_1d(v, o);
#pragma unroll
                    for (int k = 0; k < 8; k++) c[8 * x + k] = o[k] >> 6;       // (reuse: c[8 x + k] = what pixel (column x, row k) gains)
                }
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    px[y][0] = add_clip4(px[y][0], c[y], c[8 + y], c[16 + y], c[24 + y]);
                    px[y][1] = add_clip4(px[y][1], c[32 + y], c[40 + y], c[48 + y], c[56 + y]);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) *reinterpret_cast<uint4 *>(b + 8 * k) = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int y = 0; y < 8; y++) { st_px(d + (size_t)y * ls, px[y][0], al); st_px(d + (size_t)y * ls + 4, px[y][1], al); }
        }
    } else if (lane >= 16 && lane < 24 && chroma) {                 // h264_idct_add8 (:204-214): blocks 16..19 (cb), 32..35 (cr)
        const int plane = (lane - 16) >> 2, k = (lane - 16) & 3, i = 16 + 16 * plane + k;
        int16_t *b = gc + 16 * i;
        uint8_t *d = (plane ? cr : cb) + rw.y + blk_x(k) + (size_t)blk_y(k) * uvls;
        const int nnz = __ldg(nz + scan8_of(i));
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(b);
        const int kind = nnz ? 1 : (w0 & 0xffffu) ? 2 : 0;
        if (kind) {
            uint32_t px[4];
#pragma unroll
            for (int y = 0; y < 4; y++) px[y] = ld_px(d + (size_t)y * uvls, al);
            if (kind == 1) {
                const uint4 c0 = *reinterpret_cast<const uint4 *>(b), c1 = *reinterpret_cast<const uint4 *>(b + 8);
                idct4_add_regs(c0, c1, px);
                *reinterpret_cast<uint4 *>(b) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4 *>(b + 8) = make_uint4(0, 0, 0, 0);
            } else {
                const int dc = (lo16s(w0) + 32) >> 6;
#pragma unroll
                for (int y = 0; y < 4; y++) px[y] = add_clip4(px[y], dc, dc, dc, dc);
                b[0] = 0;
            }
#pragma unroll
            for (int y = 0; y < 4; y++) st_px(d + (size_t)y * uvls, px[y], al);
        }
    }
}

}  // namespace

// 0 launched, 1 not applicable (coefficient arena / planes the vector accesses cannot address): the caller runs the staged kernel
int launch_h264_residual_v2(const FFH264ResidualMB *mbs, size_t n, int16_t *coeffs, size_t coeff_stride, const uint8_t *nnzc,
                            uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls, int uvls, cudaStream_t st)
{
    if ((coeff_stride & 7) || ((uintptr_t)coeffs & 15)) return 1;
    if (((uintptr_t)luma | (uintptr_t)cb | (uintptr_t)cr | (uintptr_t)ls | (uintptr_t)uvls | (uintptr_t)mbs) & 3) return 1;
    AVB_LAUNCH(h264_residual_kernel_v2, (unsigned)((n + 3) / 4), 128, 0, st)(mbs, n, coeffs, coeff_stride, nnzc, luma, cb, cr, ls, uvls);
    return check_launch("h264_idct_add_mb_batch") ? -1 : 0;
}

}  // namespace avb
