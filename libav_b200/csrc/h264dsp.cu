// libav_b200/csrc/h264dsp.cu -- batched H.264 DSP for sm_100a: residual add, motion compensation, weighted
// prediction and in-loop deblocking, driven by per-macroblock records (include/avdsp_b200.h).
//
//   residual   one warp per macroblock, one lane per 4x4 (or 8x8) block: 32-byte coefficient rows in, 4-byte pixel
//              rows read-modify-written, consumed coefficients zeroed (parity also holds on the coefficient buffer).
//   MC         one warp per partition record, lanes stride over the partition's luma + 2 x chroma samples; reference
//              samples come through a clamped fetch (replaces emulated_edge_mc), read-only path, L1-resident window.
//   deblock    h264_deblock.cu (paired-row register-resident wavefront, rows handed out by an atomic ticket)
// All arithmetic is in h264dsp.cuh.
#include "h264dsp.cuh"
#include "../../include/avdsp_b200.h"

namespace avb {

// ---------------------------------------------------------------------------------------------------
// The macroblock's 1536 B of coefficients and its 384 pixels are staged in the warp's shared-memory slice with coalesced
// vector loads, transformed there (one lane per 4x4 / 8x8 block, as the C dispatchers iterate), and written back with
// vector stores -- the per-lane 32-byte-strided global accesses of the first version kept 4 of 32 lanes busy and the rest
// of the time waiting on memory.
struct ResSmem { __align__(16) int16_t co[768]; __align__(16) uint8_t y[16 * 16]; __align__(16) uint8_t c[2][8 * 8]; };

__global__ void __launch_bounds__(128)
h264_residual_kernel(const FFH264ResidualMB *__restrict__ mbs, size_t n, int16_t *__restrict__ coeffs, size_t coeff_stride,
                     const uint8_t *__restrict__ nnzc, uint8_t *__restrict__ luma, uint8_t *__restrict__ cb,
                     uint8_t *__restrict__ cr, int ls, int uvls)
{
    __shared__ ResSmem sm[4];
    const int lane = threadIdx.x & 31;
    ResSmem &S = sm[threadIdx.x >> 5];
    size_t mb = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (mb >= n) return;
    const FFH264ResidualMB r = mbs[mb];
    int16_t *gc = coeffs + mb * coeff_stride;
    const uint8_t *nz = nnzc + mb * 120;
    const bool do_luma = r.luma_mode <= 2, do_chroma = r.chroma != 0;
    if (!do_luma && !do_chroma) return;
    const bool vec = !(coeff_stride & 7) && !((uintptr_t)coeffs & 15);
    const bool pix4 = !((ls | uvls) & 3) && !((r.luma_off | r.chroma_off) & 3) && !(((uintptr_t)luma | (uintptr_t)cb | (uintptr_t)cr) & 3);
    // ---- stage ----
    if (vec) { for (int i = lane; i < 96; i += 32) reinterpret_cast<uint4 *>(S.co)[i] = reinterpret_cast<const uint4 *>(gc)[i]; }
    else     { for (int i = lane; i < 768; i += 32) S.co[i] = gc[i]; }
    if (pix4) {
        if (do_luma) for (int i = lane; i < 64; i += 32) reinterpret_cast<uint32_t *>(S.y)[i] = *reinterpret_cast<const uint32_t *>(luma + r.luma_off + (size_t)(i >> 2) * ls + 4 * (i & 3));
        if (do_chroma) { const int p = lane >> 4, k = lane & 15; reinterpret_cast<uint32_t *>(S.c[p])[k] = *reinterpret_cast<const uint32_t *>((p ? cr : cb) + r.chroma_off + (size_t)(k >> 1) * uvls + 4 * (k & 1)); }
    } else {
        if (do_luma) for (int i = lane; i < 256; i += 32) S.y[i] = luma[r.luma_off + (size_t)(i >> 4) * ls + (i & 15)];
        if (do_chroma) for (int i = lane; i < 128; i += 32) { const int p = i >> 6, k = i & 63; S.c[p][k] = (p ? cr : cb)[r.chroma_off + (size_t)(k >> 3) * uvls + (k & 7)]; }
    }
    __syncwarp();
    // ---- transform in shared memory ----
    if (lane < 16 && do_luma) {
        const int i = lane;
        uint8_t *d = S.y + blk_x(i) + blk_y(i) * 16;
        int16_t *b = S.co + 16 * i;
        const int nnz = nz[scan8_of(i)];
        if (r.luma_mode == 0) {                                   // h264_idct_add16, h264idct_template.c:174-183
            if (nnz) { if (nnz == 1 && b[0]) h264_dc_add(d, b, 16, 4); else h264_idct4_add(d, b, 16); }
        } else if (r.luma_mode == 1) {                            // h264_idct_add16intra, :185-191
            if (nnz) h264_idct4_add(d, b, 16); else if (b[0]) h264_dc_add(d, b, 16, 4);
        } else if ((i & 3) == 0) {                                // h264_idct8_add4, :193-202
            if (nnz) { if (nnz == 1 && b[0]) h264_dc_add(d, b, 16, 8); else h264_idct8_add(d, b, 16); }
        }
    } else if (lane >= 16 && lane < 24 && do_chroma) {            // h264_idct_add8, :204-214
        const int plane = (lane - 16) >> 2, k = (lane - 16) & 3, i = 16 + 16 * plane + k;
        uint8_t *d = S.c[plane] + blk_x(k) + blk_y(k) * 8;
        int16_t *b = S.co + 16 * i;
        if (nz[scan8_of(i)]) h264_idct4_add(d, b, 8); else if (b[0]) h264_dc_add(d, b, 8, 4);
    }
    __syncwarp();
    // ---- write back ----
    if (vec) { for (int i = lane; i < 96; i += 32) reinterpret_cast<uint4 *>(gc)[i] = reinterpret_cast<const uint4 *>(S.co)[i]; }
    else     { for (int i = lane; i < 768; i += 32) gc[i] = S.co[i]; }
    if (pix4) {
        if (do_luma) for (int i = lane; i < 64; i += 32) *reinterpret_cast<uint32_t *>(luma + r.luma_off + (size_t)(i >> 2) * ls + 4 * (i & 3)) = reinterpret_cast<const uint32_t *>(S.y)[i];
        if (do_chroma) { const int p = lane >> 4, k = lane & 15; *reinterpret_cast<uint32_t *>((p ? cr : cb) + r.chroma_off + (size_t)(k >> 1) * uvls + 4 * (k & 1)) = reinterpret_cast<const uint32_t *>(S.c[p])[k]; }
    } else {
        if (do_luma) for (int i = lane; i < 256; i += 32) luma[r.luma_off + (size_t)(i >> 4) * ls + (i & 15)] = S.y[i];
        if (do_chroma) for (int i = lane; i < 128; i += 32) { const int p = i >> 6, k = i & 63; (p ? cr : cb)[r.chroma_off + (size_t)(k >> 3) * uvls + (k & 7)] = S.c[p][k]; }
    }
}

// DC transforms + dequantisation ahead of the residual passes: what hl_decode_mb() does before the IDCTs
// (h264_mb.c:714-719 h264_luma_dc_dequant_idct for intra 16x16, h264_mb_template.c:182-189 chroma_dc_dequant_idct).
// Thread per macroblock; only the 16 + 8 DC positions of its coefficient arena are touched.
__global__ void __launch_bounds__(128)
h264_dc_dequant_kernel(int c422, const FFH264DCRecord *__restrict__ recs, size_t n, int16_t *__restrict__ coeffs, size_t coeff_stride,
                       const int16_t *__restrict__ luma_dc)
{
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const FFH264DCRecord r = recs[m];
    int16_t *mb = coeffs + m * coeff_stride;
    if (r.luma_qmul) {
        int16_t in[16];
#pragma unroll
        for (int i = 0; i < 16; i++) in[i] = luma_dc[16 * m + i];
        h264_luma_dc_dequant(mb, in, (int)r.luma_qmul);
    }
    for (int pl = 0; pl < 2; pl++) {
        if (!r.chroma_qmul[pl]) continue;
        if (c422) h264_chroma422_dc_dequant(mb + 256 * (pl + 1), (int)r.chroma_qmul[pl]);     // 4:2:2: the 2x4 transform (h264idct_template.c:271-306)
        else h264_chroma_dc_dequant(mb + 256 * (pl + 1), (int)r.chroma_qmul[pl]);
    }
}

// ---------------------------------------------------------------------------------------------------
// One warp per partition.  The (w+5) x (h+5) luma patch and the two (w/2+1) x (h/2+1) chroma patches are fetched once
// (clamped) into the warp's shared-memory slice; the unrounded horizontal 6-tap plane `tmp` (int16, rows -2..h+2) is
// built cooperatively when the position needs H or HV, then every lane finishes its pixels from shared memory:
// H = clip((tmp + 16) >> 5), HV = clip((6-tap over tmp + 512) >> 10), V = 6-tap over the patch, F = patch.
constexpr int MC_PW = 24;                 // luma patch pitch (21 used)
constexpr int MC_CW = 12;                 // chroma patch pitch (9 used)
struct MCSmem { uint8_t luma[21 * MC_PW]; int16_t tmp[21 * 16]; uint8_t chroma[2][9 * MC_CW]; };

// The partition width is a template parameter (the switch on r.w is warp uniform): every index split below is a shift or
// a division by a constant -- with run-time widths the integer divisions were most of the kernel.
template <int W>
__device__ __forceinline__ void mc_partition(MCSmem &S, const FFH264MCRecord &r, const FFH264RefPlanes &ref, uint8_t *__restrict__ dy,
                                             uint8_t *__restrict__ dcb, uint8_t *__restrict__ dcr, int ls, int uvls, int pw, int ph, int lane)
{
    constexpr int PW = W + 5, CW = W / 2 + 1, cw = W / 2;
    const int mx = r.mvx + r.x * 4, my = r.mvy + r.y * 4;          // quarter-pel position, h264_mb.c:216-217
    const int h = r.h, chh = h >> 1, fx = mx & 3, fy = my & 3;
    const int pic = r.y / ph, ly0 = pic * ph, cy0 = pic * (ph >> 1);
    // ---- fetch patches (edge replication by clamping = emulated_edge_mc) ----
    {
        const int bx = (mx >> 2) - 2, by = (my >> 2) - 2, PH = h + 5;
        for (int i = lane; i < PW * PH; i += 32) {
            const int yy = i / PW, xx = i - yy * PW;
            S.luma[yy * MC_PW + xx] = __ldg(ref.y + (size_t)min(max(by + yy, ly0), ly0 + ph - 1) * ls + min(max(bx + xx, 0), pw - 1));
        }
        const int cbx = mx >> 3, cby = my >> 3, CH = chh + 1, per = CW * CH;
        for (int i = lane; i < 2 * per; i += 32) {
            const int pl = i >= per, k = i - pl * per, yy = k / CW, xx = k - yy * CW;
            const uint8_t *src = pl ? ref.cr : ref.cb;
            S.chroma[pl][yy * MC_CW + xx] = __ldg(src + (size_t)min(max(cby + yy, cy0), cy0 + (ph >> 1) - 1) * uvls + min(max(cbx + xx, 0), (pw >> 1) - 1));
        }
    }
    __syncwarp();
    if (fx) {                                          // horizontal 6-tap, unrounded, rows -2 .. h+2 (tmp row = patch row)
        for (int i = lane; i < W * (h + 5); i += 32) {
            const int xx = i % W, yy = i / W;
            const uint8_t *p = &S.luma[yy * MC_PW + xx + 2];
            S.tmp[yy * 16 + xx] = (int16_t)((p[0] + p[1]) * 20 - (p[-1] + p[2]) * 5 + (p[-2] + p[3]));
        }
        __syncwarp();
    }
    // ---- luma ----
    for (int i = lane; i < W * h; i += 32) {
        const int px = i % W, py = i / W;
        const uint8_t *p = &S.luma[(py + 2) * MC_PW + px + 2];          // full-pel sample F(px, py)
        const int16_t *t = &S.tmp[(py + 2) * 16 + px];
        auto Hs = [&](int dyy) { return clip_u8((t[dyy * 16] + 16) >> 5); };
        auto Vs = [&](int dxx) { const uint8_t *q = p + dxx; return clip_u8(((q[0] + q[MC_PW]) * 20 - (q[-MC_PW] + q[2 * MC_PW]) * 5 + (q[-2 * MC_PW] + q[3 * MC_PW]) + 16) >> 5); };
        auto HVs = [&]() { return clip_u8(((t[0] + t[16]) * 20 - (t[-16] + t[32]) * 5 + (t[-32] + t[48]) + 512) >> 10); };
        int a, b = -1;                                                   // h264qpel_template.c:380-531
        if (!fx && !fy) a = p[0];
        else if (!fy) { a = Hs(0); if (fx != 2) b = p[fx == 3]; }
        else if (!fx) { a = Vs(0); if (fy != 2) b = p[(fy == 3) * MC_PW]; }
        else if (fx == 2 && fy == 2) a = HVs();
        else if (fx == 2) { a = HVs(); b = Hs(fy == 3); }
        else if (fy == 2) { a = HVs(); b = Vs(fx == 3); }
        else { a = Hs(fy == 3); b = Vs(fx == 3); }
        const int v = b < 0 ? a : (a + b + 1) >> 1;
        uint8_t *d = dy + (size_t)(r.y + py) * ls + r.x + px;
        *d = (uint8_t)(r.avg ? (*d + v + 1) >> 1 : v);
    }
    // ---- chroma (1/8-pel bilinear, h264chroma_template.c:27-170) ----
    {
        const int cfx = mx & 7, cfy = my & 7, A = (8 - cfx) * (8 - cfy), B = cfx * (8 - cfy), Cc = (8 - cfx) * cfy, D = cfx * cfy;
        const int per = cw * chh;
        for (int i = lane; i < 2 * per; i += 32) {
            const int pl = i >= per, k = i - pl * per, px = k % cw, py = k / cw;
            const uint8_t *p = &S.chroma[pl][py * MC_CW + px];
            const int v = (A * p[0] + B * p[1] + Cc * p[MC_CW] + D * p[MC_CW + 1] + 32) >> 6;
            uint8_t *d = (pl ? dcr : dcb) + (size_t)((r.y >> 1) + py) * uvls + (r.x >> 1) + px;
            *d = (uint8_t)(r.avg ? (*d + v + 1) >> 1 : v);
        }
    }
}

__global__ void __launch_bounds__(128)
h264_mc_kernel(const FFH264MCRecord *__restrict__ recs, size_t n, const FFH264RefPlanes *__restrict__ refs,
               uint8_t *__restrict__ dy, uint8_t *__restrict__ dcb, uint8_t *__restrict__ dcr, int ls, int uvls, int pw, int ph,
               int pass)   // ph = height of ONE picture; pictures of a batch are stacked vertically
{
    __shared__ MCSmem sm[4];
    const int lane = threadIdx.x & 31;
    MCSmem &S = sm[threadIdx.x >> 5];
    size_t ri = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ri >= n) return;
    const FFH264MCRecord r = recs[ri];
    if ((r.avg != 0) != (pass != 0)) return;          // pass 0: every `put`; pass 1: the `avg` second directions
    const FFH264RefPlanes ref = refs[r.ref];
    if (r.w == 16)     mc_partition<16>(S, r, ref, dy, dcb, dcr, ls, uvls, pw, ph, lane);
    else if (r.w == 8) mc_partition<8>(S, r, ref, dy, dcb, dcr, ls, uvls, pw, ph, lane);
    else               mc_partition<4>(S, r, ref, dy, dcb, dcr, ls, uvls, pw, ph, lane);
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
h264_weight_kernel(const FFH264WeightRecord *__restrict__ recs, size_t n, uint8_t *__restrict__ plane,
                   const uint8_t *__restrict__ src, int stride)
{
    const int lane = threadIdx.x & 31;
    size_t ri = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ri >= n) return;
    const FFH264WeightRecord r = recs[ri];
    const int ld = r.log2_denom;
    int off = r.offset;
    if (!src) { off <<= ld; if (ld) off += 1 << (ld - 1); }       // h264dsp_template.c:39-40
    else off = ((off + 1) | 1) << ld;                             // :70-71
    for (int it = lane; it < r.w * r.h; it += 32) {
        size_t o = r.off + (size_t)(it / r.w) * stride + it % r.w;
        int v = src ? (src[o] * r.weight_src + plane[o] * r.weight + off) >> (ld + 1) : (plane[o] * r.weight + off) >> ld;
        plane[o] = (uint8_t)clip_u8(v);
    }
}

// ---------------------------------------------------------------------------------------------------
// Deblocking wavefront: h264_deblock.cu (the one-warp-per-row kernel that lived here relied on CTAs being dispatched in block-index order and
// is gone).
int launch_h264_deblock_v3(const FFH264DeblockMB *mbs, int mb_w, int mb_h, int n_pictures, uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls,
                           int uvls, uint32_t *progress, cudaStream_t st);       // h264_deblock.cu

int launch_h264_mc_v2(const FFH264MCRecord *recs, size_t n, const FFH264RefPlanes *refs, uint8_t *dy, uint8_t *dcb, uint8_t *dcr,
                      int ls, int uvls, int pw, int ph, cudaStream_t st);                  // h264_mc.cu

int launch_h264_residual_v2(const FFH264ResidualMB *mbs, size_t n, int16_t *coeffs, size_t coeff_stride, const uint8_t *nnzc,
                            uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls, int uvls, cudaStream_t st);      // h264_residual.cu

static int warps_grid(size_t n, int warps_per_cta) { return (int)((n + warps_per_cta - 1) / warps_per_cta); }

int launch_h264_residual(const FFH264ResidualMB *mbs, size_t n, int16_t *coeffs, size_t coeff_stride, const uint8_t *nnzc,
                         uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls, int uvls, cudaStream_t st)
{
    if (!n) return 0;
    if (coeff_stride < 16 * 36) { set_error_msg("h264_idct_add_mb_batch", "coeff_stride must cover blocks 0..35 (>= 576)"); return -1; }
    // default: lane per block in registers, touching only what the C dispatchers touch (h264_residual.cu); residual_variant = 2 keeps
    // the shared-memory staged kernel, which also serves unaligned arenas / planes and records whose offsets are not multiples of 4
    if (tuning("residual_variant") != 2) {
        const int r = launch_h264_residual_v2(mbs, n, coeffs, coeff_stride, nnzc, luma, cb, cr, ls, uvls, st);
        if (r <= 0) return r;
    }
    h264_residual_kernel<<<warps_grid(n, 4), 128, 0, st>>>(mbs, n, coeffs, coeff_stride, nnzc, luma, cb, cr, ls, uvls);
    return check_launch("h264_idct_add_mb_batch");
}
int launch_h264_mc(const FFH264MCRecord *recs, size_t n, const FFH264RefPlanes *refs, uint8_t *dy, uint8_t *dcb, uint8_t *dcr,
                   int ls, int uvls, int pw, int ph, cudaStream_t st)
{
    if (!n) return 0;
    // default: one thread per 4x4 block on packed bytes (h264_mc.cu); mc_variant = 2 keeps the warp-per-partition kernel below, which
    // also serves destinations whose planes the word stores cannot address
    if (tuning("mc_variant") != 2) {
        const int r = launch_h264_mc_v2(recs, n, refs, dy, dcb, dcr, ls, uvls, pw, ph, st);
        if (r <= 0) return r;
    }
    // bi-prediction is put (list 0) then avg (list 1) on the same pixels (h264_mb.c:322-366): two ordered passes
    // over the record array keep that order without any ordering requirement on the records themselves
    h264_mc_kernel<<<warps_grid(n, 4), 128, 0, st>>>(recs, n, refs, dy, dcb, dcr, ls, uvls, pw, ph, 0);
    h264_mc_kernel<<<warps_grid(n, 4), 128, 0, st>>>(recs, n, refs, dy, dcb, dcr, ls, uvls, pw, ph, 1);
    return check_launch("h264_mc_batch");
}
int launch_h264_weight(const FFH264WeightRecord *recs, size_t n, uint8_t *plane, const uint8_t *src, int stride, cudaStream_t st)
{
    if (!n) return 0;
    h264_weight_kernel<<<warps_grid(n, 4), 128, 0, st>>>(recs, n, plane, src, stride);
    return check_launch("h264_weight_batch");
}
int launch_h264_deblock(const FFH264DeblockMB *mbs, int mb_w, int mb_h, int n_pictures, uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls,
                        int uvls, uint32_t *progress, cudaStream_t st)
{
    if (mb_w <= 0 || mb_h <= 0 || n_pictures <= 0) return 0;
    if ((ls & 3) || (uvls & 3) || ((uintptr_t)luma & 3) || ((uintptr_t)cb & 3) || ((uintptr_t)cr & 3)) {
        set_error_msg("h264_deblock_picture", "planes and line sizes must be 4-byte aligned"); return -1;
    }
    // the paired-row, register-resident wavefront of h264_deblock.cu (rows handed out by an atomic ticket)
    return launch_h264_deblock_v3(mbs, mb_w, mb_h, n_pictures, luma, cb, cr, ls, uvls, progress, st);
}

}  // namespace avb

using namespace avb;
extern "C" {
int ff_h264_idct_add_mb_batch_cuda(const FFH264ResidualMB *mbs, size_t n, int16_t *coeffs, size_t coeff_stride, const uint8_t *nnzc,
                                   uint8_t *luma, uint8_t *cb, uint8_t *cr, int linesize, int uvlinesize, void *stream)
{ avb::enter(); return launch_h264_residual(mbs, n, coeffs, coeff_stride, nnzc, luma, cb, cr, linesize, uvlinesize, (cudaStream_t)stream); }
int ff_h264_mc_batch_cuda(const FFH264MCRecord *recs, size_t n, const FFH264RefPlanes *refs, uint8_t *dst_y, uint8_t *dst_cb,
                          uint8_t *dst_cr, int linesize, int uvlinesize, int pic_w, int pic_h, void *stream)
{ avb::enter(); return launch_h264_mc(recs, n, refs, dst_y, dst_cb, dst_cr, linesize, uvlinesize, pic_w, pic_h, (cudaStream_t)stream); }
int ff_h264_weight_batch_cuda(const FFH264WeightRecord *recs, size_t n, uint8_t *plane, const uint8_t *src, int stride, void *stream)
{ avb::enter(); return launch_h264_weight(recs, n, plane, src, stride, (cudaStream_t)stream); }
int ff_h264_dc_dequant_batch_cuda(const FFH264DCRecord *recs, size_t n, int16_t *coeffs, size_t coeff_stride, const int16_t *luma_dc,
                                  void *stream)
{
    avb::enter();
    if (!n) return 0;
    h264_dc_dequant_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(0, recs, n, coeffs, coeff_stride, luma_dc);
    return check_launch("ff_h264_dc_dequant_batch_cuda");
}
int ff_h264_dc_dequant_batch_422_cuda(const FFH264DCRecord *recs, size_t n, int16_t *coeffs, size_t coeff_stride, const int16_t *luma_dc,
                                      void *stream)
{
    avb::enter();
    if (n && (!recs || !coeffs || !luma_dc)) { set_error_msg("ff_h264_dc_dequant_batch_422_cuda", "NULL argument"); return -1; }
    if (!n) return 0;
    h264_dc_dequant_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(1, recs, n, coeffs, coeff_stride, luma_dc);
    return check_launch("ff_h264_dc_dequant_batch_422_cuda");
}
int ff_h264_deblock_picture_cuda(const FFH264DeblockMB *mbs, int mb_w, int mb_h, uint8_t *luma, uint8_t *cb, uint8_t *cr,
                                 int linesize, int uvlinesize, uint32_t *progress, void *stream)
{ avb::enter(); return launch_h264_deblock(mbs, mb_w, mb_h, 1, luma, cb, cr, linesize, uvlinesize, progress, (cudaStream_t)stream); }
int ff_h264_deblock_batch_cuda(const FFH264DeblockMB *mbs, int mb_w, int mb_h, int n_pictures, uint8_t *luma, uint8_t *cb, uint8_t *cr,
                               int linesize, int uvlinesize, uint32_t *progress, void *stream)
{ avb::enter(); return launch_h264_deblock(mbs, mb_w, mb_h, n_pictures, luma, cb, cr, linesize, uvlinesize, progress, (cudaStream_t)stream); }
}
