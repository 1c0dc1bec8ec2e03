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

// One warp per macroblock.  Two code paths, each executed once by all the lanes that have work (a warp that let every lane branch on
// its own block kind ran 5 of 32 threads per instruction):
//   4x4 path   lanes 0-15 = the luma blocks of an idct_add16 / add16intra macroblock, lanes 16-23 = the cb / cr blocks; the full
//              transform and the DC-only shortcut are computed side by side and selected per lane
//   8x8 path   a transform-8x8 macroblock: 8 lanes per block (32 busy lanes); a lane owns coefficient row t, then column t, then row t
//              again, then pixel row t -- the three transposes go through 4 x 144 bytes of shared memory
__global__ void __launch_bounds__(128)
h264_residual_kernel_v2(const FFH264ResidualMB *__restrict__ mbs, size_t n, int16_t *__restrict__ coeffs, size_t coeff_stride,
                        const uint8_t *__restrict__ nnzc, uint8_t *__restrict__ luma, uint8_t *__restrict__ cb,
                        uint8_t *__restrict__ cr, int ls, int uvls)
{
    const int lane = threadIdx.x & 31;
    const size_t mb = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (mb >= n) return;
    const uint32_t *rp = reinterpret_cast<const uint32_t *>(mbs + mb);      // (12-byte records: word loads)
    const uint2 rw = make_uint2(__ldg(rp), __ldg(rp + 1));                   // luma_off, chroma_off
    const uint32_t rm = __ldg(rp + 2);
    const int mode = rm & 255, chroma = (rm >> 8) & 255;
    const bool al = !((rw.x | rw.y) & 3u);
    int16_t *gc = coeffs + mb * coeff_stride;
    const uint8_t *nz = nnzc + mb * 120;

#ifndef AVB_HOSTSIM
    if (mode == 2) {                                                // h264_idct8_add4 (:193-202): blocks 0, 4, 8, 12
        __shared__ __align__(16) int16_t tr[4][4][72];              // [warp][block][8 rows x 8 + 8 pad]: 144 bytes per block keeps the groups on different banks
        const int g = lane >> 3, t = lane & 7, i = 4 * g;
        int16_t (&T)[72] = tr[threadIdx.x >> 5][g];
        const int nnz = __ldg(nz + scan8_of(i));
        int16_t *b = gc + 16 * i;
        uint8_t *d = luma + rw.x + blk_x(i) + (size_t)(blk_y(i) + t) * ls;         // this lane's pixel row
        const int b0 = nnz ? (int)b[0] : 0;
        const int kind = !nnz ? 0 : (nnz == 1 && b0) ? 2 : 1;       // uniform over the block's 8 lanes
        uint32_t p0 = 0, p1 = 0;
        if (kind) { p0 = ld_px(d, al); p1 = ld_px(d + 4, al); }
        int gain[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        uint4 q = make_uint4(0, 0, 0, 0);
        if (kind == 1) q = *reinterpret_cast<const uint4 *>(b + 8 * t);          // coefficient row t
        if (kind == 1 && t == 0) q.x = (q.x & 0xffff0000u) | (uint32_t)(uint16_t)(lo16s(q.x) + 32);   // block[0] += 32 (int16)
        *reinterpret_cast<uint4 *>(&T[8 * t]) = q;
        __syncwarp();
        int v[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = T[8 * k + t];            // column t
        h264_idct8_1d(v, o);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) T[8 * k + t] = (int16_t)o[k];   // written back as int16 like the C code
        __syncwarp();
        {
            const uint4 r = *reinterpret_cast<const uint4 *>(&T[8 * t]);         // row t of the half-transformed block
            v[0] = lo16s(r.x); v[1] = hi16s(r.x); v[2] = lo16s(r.y); v[3] = hi16s(r.y); v[4] = lo16s(r.z); v[5] = hi16s(r.z); v[6] = lo16s(r.w); v[7] = hi16s(r.w);
        }
        h264_idct8_1d(v, o);                                        // o[k] >> 6 is what pixel (column t, row k) gains
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) T[8 * k + t] = (int16_t)(o[k] >> 6);
        __syncwarp();
        {
            const uint4 r = *reinterpret_cast<const uint4 *>(&T[8 * t]);         // pixel row t
            gain[0] = lo16s(r.x); gain[1] = hi16s(r.x); gain[2] = lo16s(r.y); gain[3] = hi16s(r.y); gain[4] = lo16s(r.z); gain[5] = hi16s(r.z); gain[6] = lo16s(r.w); gain[7] = hi16s(r.w);
        }
        if (kind == 2) {                                            // ff_h264_idct8_dc_add
            const int dc = (b0 + 32) >> 6;
#pragma unroll
            for (int k = 0; k < 8; k++) gain[k] = dc;
        }
        if (kind) {
            st_px(d, add_clip4(p0, gain[0], gain[1], gain[2], gain[3]), al);
            st_px(d + 4, add_clip4(p1, gain[4], gain[5], gain[6], gain[7]), al);
            if (kind == 1) *reinterpret_cast<uint4 *>(b + 8 * t) = make_uint4(0, 0, 0, 0);
            else if (t == 0) b[0] = 0;
        }
    }
#endif
    // ---- 4x4 blocks: luma (h264_idct_add16 :174-183, h264_idct_add16intra :185-191) and chroma (h264_idct_add8 :204-214) ----
    {
        const bool is_luma = lane < 16;
        const int plane = (lane - 16) >> 2, k = is_luma ? lane : (lane - 16) & 3;
        const int i = is_luma ? lane : 16 + 16 * plane + k;
        const bool mine = is_luma ? mode <= 1 : (lane < 24 && chroma);
        if (mine) {
            int16_t *b = gc + 16 * i;
            const int pitch = is_luma ? ls : uvls;
            uint8_t *d = (is_luma ? luma + rw.x : (plane ? cr : cb) + rw.y) + blk_x(k) + (size_t)blk_y(k) * pitch;
            const int nnz = __ldg(nz + scan8_of(i));
            const bool dc_rule_inter = is_luma && mode == 0;        // idct_add16: DC-only shortcut only when nnz == 1; the others: when nnz == 0
            int kind = 0;                                           // 1 full transform, 2 DC only
            uint32_t w0 = 0;
            if (nnz || !dc_rule_inter) w0 = *reinterpret_cast<const uint32_t *>(b);
            if (nnz) kind = (dc_rule_inter && nnz == 1 && (w0 & 0xffffu)) ? 2 : 1;
            else if (!dc_rule_inter) kind = (w0 & 0xffffu) ? 2 : 0;
            if (kind) {
                uint32_t px[4], pd[4];
#pragma unroll
                for (int y = 0; y < 4; y++) pd[y] = px[y] = ld_px(d + (size_t)y * pitch, al);
                const uint4 c0 = *reinterpret_cast<const uint4 *>(b), c1 = *reinterpret_cast<const uint4 *>(b + 8);
                idct4_add_regs(c0, c1, px);                         // (computed for DC-only lanes too: the warp runs it anyway)
                const int dc = (lo16s(w0) + 32) >> 6;
#pragma unroll
                for (int y = 0; y < 4; y++) st_px(d + (size_t)y * pitch, kind == 1 ? px[y] : add_clip4(pd[y], dc, dc, dc, dc), al);
                if (kind == 1) { *reinterpret_cast<uint4 *>(b) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4 *>(b + 8) = make_uint4(0, 0, 0, 0); }
                else b[0] = 0;
            }
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
