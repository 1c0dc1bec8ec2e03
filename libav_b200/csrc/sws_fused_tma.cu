// libav_b200/csrc/sws_fused_tma.cu -- same-size yuv420p -> rgb24 / bgr24 (the 4K benchmark geometry) with TMA-staged line buffers.
//
// Same arithmetic as sws_fused_rgb24_v3_kernel (swscale.cu): yuv2rgb24_X_c with a one-tap luma and a four-tap chroma bank
// (libswscale/output.c:936-995 + write :853-866), the line window of swscale() (swscale.c:450-616), the colour tables of yuv2rgb.c:633-658
// evaluated arithmetically.  What changes is how the bytes travel and how many instructions a pixel costs:
//   * a warp owns a 256 x 8 pixel tile (four row pairs).  Its luma rows (8 x 256 B) and the eight chroma lines of each plane that the
//     tile's vertical windows touch (8 x 128 B each) arrive by three cp.async.bulk.tensor (TMA, 3-D maps x / line / frame, zero fill
//     outside the plane) into the warp's own double-buffered stage and are awaited on the warp's own mbarrier: no CTA-wide barrier
//     exists after start-up, the tile after the current one is always in flight.  Warps are persistent and stride over the tiles of all
//     frames of the batch.
//   * a thread owns 16 pixels x 4 rows (row pairs `half` and `half + 2` of the tile).  The seven chroma lines they need are
//     byte-transposed once (window 0-3), the windows of the next rows are one PRMT away (drop the oldest line, append the next); the
//     4-tap FIR is two IDP.2A on int16 taps.
//   * the rows leave through shared memory: four 768-byte rows per cp.async.bulk.tensor store (whole 128-byte lines per request).
//   * the per-(U, V) additive terms  cy * ((V * crv) >> 16) + kr,  cy * ((V * cgv) >> 16),  cy * ((U * cbu) >> 16) + kb,
//     cy * ((U * cgu) >> 16) + kg  come from two 256-entry tables in shared memory, built by the CTA from the context's constants
//     with exactly these formulas, 16 copies of each 8-byte entry so that every lane of a half-warp reads its own bank pair
//     (conflict-free whatever the picture holds): 2 LDS.64 + 1 IADD instead of 12 integer instructions per chroma sample and row.
//   * tiles at the top / bottom of the plane (clamped lines, shifted filters) and ragged last tiles take a row-by-row path with the
//     same per-row routine.
#include "sws_fused.h"
#include "tma.h"

namespace avb {

namespace {

constexpr int FT_STAGE = 4096;                  // Y 8 rows x 256 B | U 8 lines x 128 B | V 8 lines x 128 B
constexpr int FT_LUT = 2 * 256 * 16 * 8;        // table by V {tr, tgv}, table by U {tb, tgu}: 256 entries x 16 lane copies x 8 B
constexpr int FT_OUT = 4 * 768;                 // output staging: four 768-byte rows (half a tile) per bulk tensor store

struct FusedTmaArgs {
    uint8_t *dst; int dstStride; size_t dstFrame;
    int dstW, chrSrcH;
    int rp0, rowEnd;                  // first row pair of the launch (multiple of 4), first output row past it
    int tilesX, nTy;                  // tiles per tile row, tile rows of the launch
    int nitems;                       // nframes * nTy * tilesX
    int stepX, stepTy, stepF;         // the (tx, ty, frame) decomposition of the warp stride gridDim.x * WARPS
};

__device__ __forceinline__ int dp2a_lo(uint32_t k, uint32_t b, int c)
{ int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(k), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi(uint32_t k, uint32_t b, int c)
{ int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(k), "r"(b), "r"(c)); return d; }

// lines r0..r3 hold 4 chroma columns each (one byte per column): out[c] = (r0.c, r1.c, r2.c, r3.c)
__device__ __forceinline__ void transpose4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t *out)
{
    const uint32_t t0 = __byte_perm(r0, r1, 0x5140), t1 = __byte_perm(r2, r3, 0x5140), t2 = __byte_perm(r0, r1, 0x7362), t3 = __byte_perm(r2, r3, 0x7362);
    out[0] = __byte_perm(t0, t1, 0x5410); out[1] = __byte_perm(t0, t1, 0x7632); out[2] = __byte_perm(t2, t3, 0x5410); out[3] = __byte_perm(t2, t3, 0x7632);
}
// the window one line lower: drop the oldest line, append column c of `next`
template <int C> __device__ __forceinline__ uint32_t slide(uint32_t w, uint32_t next) { return __byte_perm(w, next, ((4 + C) << 12) | 0x321); }
__device__ __forceinline__ void slide4(const uint32_t *w, uint32_t next, uint32_t *out)
{ out[0] = slide<0>(w[0], next); out[1] = slide<1>(w[1], next); out[2] = slide<2>(w[2], next); out[3] = slide<3>(w[3], next); }

// one output row of 16 pixels: TU / TV = the row's four chroma lines of each of the 8 chroma columns, yy = 16 luma bytes
template <bool BGR, bool LUT>
__device__ __forceinline__ void row16(const uint32_t (&TU)[8], const uint32_t (&TV)[8], uint32_t k01, uint32_t k23, uint4 yy,
                                      const RgbConstants &k, unsigned lutU_s, unsigned lutV_s, uint32_t (&o)[12])
{
    const int cy = k.cy;
#pragma unroll
    for (int q = 0; q < 4; q++) {                            // 4 pixels = 2 chroma columns per step
        const uint32_t yw = q == 0 ? yy.x : q == 1 ? yy.y : q == 2 ? yy.z : yy.w;
        int r[4], g[4], b[4];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int c = 2 * q + e;
            // (2^18 + sum (u << 7) * coef) >> 19  ==  (2^11 + sum u * coef) >> 12
            const int su = dp2a_hi(k23, TU[c], dp2a_lo(k01, TU[c], 2048));
            const int sv = dp2a_hi(k23, TV[c], dp2a_lo(k01, TV[c], 2048));
            const int U = sat255(su >> 12), V = sat255(sv >> 12);        // per-value clip == the reference's flagged clip (host-checked range)
            int tr, tg, tb;
            if (LUT) {
                const uint2 lv = lds64_ro(lutV_s + V * 128), lu = lds64_ro(lutU_s + U * 128);
                tr = (int)lv.x; tb = (int)lu.x; tg = (int)(lu.y + lv.y);
            } else {
                tr = cy * ((V * k.crv) >> 16) + k.kr;
                tg = cy * (((U * k.cgu) >> 16) + ((V * k.cgv) >> 16)) + k.kg;
                tb = cy * ((U * k.cbu) >> 16) + k.kb;
            }
            const int Ya = byte_of(yw, 2 * e), Yb = byte_of(yw, 2 * e + 1);
            r[2 * e] = cy * Ya + (BGR ? tb : tr); g[2 * e] = cy * Ya + tg; b[2 * e] = cy * Ya + (BGR ? tr : tb);
            r[2 * e + 1] = cy * Yb + (BGR ? tb : tr); g[2 * e + 1] = cy * Yb + tg; b[2 * e + 1] = cy * Yb + (BGR ? tr : tb);
        }
        o[3 * q + 0] = pack4_hi16_sat(r[0], g[0], b[0], r[1]);
        o[3 * q + 1] = pack4_hi16_sat(g[1], b[1], r[2], g[2]);
        o[3 * q + 2] = pack4_hi16_sat(b[2], r[3], g[3], b[3]);
    }
}

template <bool BGR, bool LUT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
sws_fused_rgb24_tma_kernel(const RgbConstants k, const FusedTmaArgs a, const SwsPairTapsT *__restrict__ taps,
                           const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
                           const __grid_constant__ CUtensorMap tmD)
{
    extern __shared__ __align__(1024) uint8_t ft_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(ft_smem);
    const unsigned stage_s = smem_s + (LUT ? FT_LUT : 0) + warp * 2 * FT_STAGE;
    const unsigned out_s = smem_s + (LUT ? FT_LUT : 0) + WARPS * 2 * FT_STAGE + warp * FT_OUT;
    const unsigned mbar_s = smem_s + (LUT ? FT_LUT : 0) + WARPS * (2 * FT_STAGE + FT_OUT) + warp * 16;

    if (LUT) {                          // the two term tables, every entry 16 times (one copy per lane of a half-warp)
        uint2 *lut = reinterpret_cast<uint2 *>(ft_smem);
        for (int e = threadIdx.x; e < 2 * 256 * 16; e += WARPS * 32) {
            const int i = (e >> 4) & 255;
            uint2 val;
            if (e < 256 * 16) { val.x = (uint32_t)(k.cy * ((i * k.crv) >> 16) + k.kr); val.y = (uint32_t)(k.cy * ((i * k.cgv) >> 16)); }
            else              { val.x = (uint32_t)(k.cy * ((i * k.cbu) >> 16) + k.kb); val.y = (uint32_t)(k.cy * ((i * k.cgu) >> 16) + k.kg); }
            lut[e] = val;
        }
    }
    if (lane == 0) { mbar_init(mbar_s, 1); mbar_init(mbar_s + 8, 1); mbar_init_fence(); }
    __syncthreads();

    const int cg = lane & 15, half = lane >> 4;
    const unsigned lutV_s = smem_s + cg * 8, lutU_s = smem_s + 256 * 128 + cg * 8;

    int item = blockIdx.x * WARPS + warp;
    if (item >= a.nitems) return;
    const int wstride = gridDim.x * WARPS;
    // three tiles are known at any time: the one being computed (c), the one in flight (n) and the one whose table words are being
    // fetched (m); every table read is issued a whole tile before its value is used
    int ctx = item % a.tilesX, cty, cf;
    { const int t2 = item / a.tilesX; cty = t2 % a.nTy; cf = t2 / a.nTy; }
    auto advance = [&](int &tx, int &ty, int &f) {
        tx += a.stepX; f += a.stepF;
        if (tx >= a.tilesX) { tx -= a.tilesX; ty++; }
        ty += a.stepTy;
        if (ty >= a.nTy) { ty -= a.nTy; f++; }
    };
    auto tile_words = [&](int ty) { return __ldg(reinterpret_cast<const int2 *>(&taps[a.rp0 + 4 * ty].base)); };   // (base, interior)
    auto issue = [&](int tx, int ty, int f, int base, unsigned st) {
        if (lane == 0) {
            const unsigned mb = mbar_s + (st ? 8u : 0u), dst = stage_s + st * FT_STAGE;
            fence_proxy_async();                              // the stage was last read by ordinary shared loads
            mbar_expect_tx(mb, FT_STAGE);
            tma_load_3d(dst, &tmY, tx * 256, 2 * (a.rp0 + 4 * ty), f, mb);
            tma_load_3d(dst + 2048, &tmU, tx * 128, base, f, mb);
            tma_load_3d(dst + 3072, &tmV, tx * 128, base, f, mb);
        }
    };
    int2 cw = tile_words(cty);
    issue(ctx, cty, cf, cw.x, 0);
    int ntx = ctx, nty = cty, nf = cf;
    advance(ntx, nty, nf);
    int2 nw = item + wstride < a.nitems ? tile_words(nty) : make_int2(0, 0);

    unsigned st = 0, uses = 0;
    for (; item < a.nitems; item += wstride, st ^= 1u, uses++) {
        if (item + wstride < a.nitems) issue(ntx, nty, nf, nw.x, st ^ 1u);       // the next tile's bytes start travelling before this one is touched
        int mtx = ntx, mty = nty, mf = nf;
        advance(mtx, mty, mf);
        const int2 mw = (long long)item + 2ll * wstride < a.nitems ? tile_words(mty) : make_int2(0, 0);

        const int rpT = a.rp0 + 4 * cty;
        mbar_wait(mbar_s + (st ? 8u : 0u), (uses >> 1) & 1u);                   // k-th use of a stage completes its phase k
        const unsigned sY = stage_s + st * FT_STAGE, sU = sY + 2048, sV = sY + 3072;
        // Rows leave the warp four at a time: the half-warps' first pairs are tile rows 0-3, their second pairs rows 4-7.  A row is staged
        // in the warp's 4 x 768-byte slab; one bulk tensor store writes the four rows (whole 128-byte lines: three 16-byte STG per lane
        // 48 bytes apart cost twice the L1 -> L2 write transactions, measured).  Columns past the width / rows past the range are clipped
        // by the tensor map.
        auto stage_row = [&](int slot, const uint32_t (&o)[12]) {
#pragma unroll
            for (int j = 0; j < 3; j++) sts128(out_s + slot * 768 + cg * 48 + 16 * j, make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
        };
        auto slab_free = [&]() { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); };       // the previous store has read the slab
        auto flush = [&](int g) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_3d(&tmD, ctx * 192, 2 * rpT + 4 * g, cf, out_s); bulk_commit(); }
        };

        if (cw.y) {
            // interior tile: pair q of the tile filters chroma lines q .. q + 3 (first row) and q + 1 .. q + 4 (second row) of the stage.
            // This thread's pairs are `half` and `half + 2`: seven lines, four windows one line apart
            uint4 yy[4];
            uint2 ul[7], vl[7];
#pragma unroll
            for (int r = 0; r < 4; r++) yy[r] = lds128(sY + (2 * half + (r & 1) + 4 * (r >> 1)) * 256 + cg * 16);
#pragma unroll
            for (int j = 0; j < 7; j++) { ul[j] = lds64(sU + (half + j) * 128 + cg * 8); vl[j] = lds64(sV + (half + j) * 128 + cg * 8); }
            const uint4 tA = __ldg(reinterpret_cast<const uint4 *>(taps + rpT + half)), tB = __ldg(reinterpret_cast<const uint4 *>(taps + rpT + half + 2));
            __syncwarp();
            uint32_t WU[8], WV[8], XU[8], XV[8], o[12];
            transpose4(ul[0].x, ul[1].x, ul[2].x, ul[3].x, WU); transpose4(ul[0].y, ul[1].y, ul[2].y, ul[3].y, WU + 4);
            transpose4(vl[0].x, vl[1].x, vl[2].x, vl[3].x, WV); transpose4(vl[0].y, vl[1].y, vl[2].y, vl[3].y, WV + 4);
            slab_free();
            row16<BGR, LUT>(WU, WV, tA.x, tA.y, yy[0], k, lutU_s, lutV_s, o); stage_row(2 * half, o);
            slide4(WU, ul[4].x, XU); slide4(WU + 4, ul[4].y, XU + 4); slide4(WV, vl[4].x, XV); slide4(WV + 4, vl[4].y, XV + 4);
            row16<BGR, LUT>(XU, XV, tA.z, tA.w, yy[1], k, lutU_s, lutV_s, o); stage_row(2 * half + 1, o);
            flush(0);
            slide4(XU, ul[5].x, WU); slide4(XU + 4, ul[5].y, WU + 4); slide4(XV, vl[5].x, WV); slide4(XV + 4, vl[5].y, WV + 4);
            row16<BGR, LUT>(WU, WV, tB.x, tB.y, yy[2], k, lutU_s, lutV_s, o);
            slab_free();
            stage_row(2 * half, o);
            slide4(WU, ul[6].x, XU); slide4(WU + 4, ul[6].y, XU + 4); slide4(WV, vl[6].x, XV); slide4(WV + 4, vl[6].y, XV + 4);
            row16<BGR, LUT>(XU, XV, tB.z, tB.w, yy[3], k, lutU_s, lutV_s, o); stage_row(2 * half + 1, o);
            flush(1);
        } else {
            // plane edges and ragged tiles: every row fetches its own four (clamped) lines
#pragma unroll 1
            for (int r = 0; r < 4; r++) {
                const int trow = 2 * half + (r & 1) + 4 * (r >> 1);
                const int row = min(2 * rpT + trow, a.rowEnd - 1);               // rows past the launch's range compute a discarded duplicate: the loop stays warp-uniform
                const SwsPairTapsT *t = taps + (row >> 1);
                const uint2 kk = __ldg(reinterpret_cast<const uint2 *>(t) + (row & 1));
                const int first = __ldg(&t->first0 + (row & 1));
                uint2 ul[4], vl[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int line = min(max(min(max(first + j, 0), a.chrSrcH - 1) - cw.x, 0), 7);   // 0..7 (host-checked for rows inside the range)
                    ul[j] = lds64(sU + line * 128 + cg * 8); vl[j] = lds64(sV + line * 128 + cg * 8);
                }
                const uint4 yy = lds128(sY + trow * 256 + cg * 16);
                uint32_t TU[8], TV[8], o[12];
                transpose4(ul[0].x, ul[1].x, ul[2].x, ul[3].x, TU); transpose4(ul[0].y, ul[1].y, ul[2].y, ul[3].y, TU + 4);
                transpose4(vl[0].x, vl[1].x, vl[2].x, vl[3].x, TV); transpose4(vl[0].y, vl[1].y, vl[2].y, vl[3].y, TV + 4);
                row16<BGR, LUT>(TU, TV, kk.x, kk.y, yy, k, lutU_s, lutV_s, o);
                if (!(r & 1)) slab_free();
                stage_row(trow & 3, o);
                if (r & 1) flush(r >> 1);
            }
            __syncwarp();
        }
        ctx = ntx; cty = nty; cf = nf; cw = nw;
        ntx = mtx; nty = mty; nf = mf; nw = mw;
    }
    if (lane == 0) bulk_wait<0>();
}

}  // namespace

template <bool BGR, bool LUT, int WARPS>
static int launch_variant(const RgbConstants &k, const FusedTmaArgs &fa, const SwsPairTapsT *taps, const CUtensorMap &tmY, const CUtensorMap &tmU,
                          const CUtensorMap &tmV, const CUtensorMap &tmD, int grid, cudaStream_t st)
{
    const size_t smem = (LUT ? FT_LUT : 0) + (size_t)WARPS * (2 * FT_STAGE + FT_OUT) + WARPS * 16;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(sws_fused_rgb24_tma_kernel<BGR, LUT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("sws_scale:fused tma", cudaGetLastError()); return -1;
        }
        attr_done = true;
    }
    sws_fused_rgb24_tma_kernel<BGR, LUT, WARPS><<<grid, WARPS * 32, smem, st>>>(k, fa, taps, tmY, tmU, tmV, tmD);
    return check_launch("sws_scale:fused tma");
}

int sws_fused_tma_launch(const RgbConstants &k, int bgr, int dstW, int dstH, int chrSrcW, int chrSrcH, const FusedArgs &a,
                         const SwsPairTapsT *taps, int nframes, cudaStream_t st)
{
    const int pairs = dstH / 2;
    const int rp0 = a.rp0, rp1 = a.rp1 < pairs ? a.rp1 : pairs;
    if (nframes <= 0 || rp1 <= rp0) return 0;
    // what a tensor map can describe: 16-byte aligned planes and pitches; tiles are aligned to multiples of four row pairs
    if ((rp0 & 3) || ((rp1 & 3) && rp1 != pairs)) return 1;
    if (((uintptr_t)a.y | (uintptr_t)a.u | (uintptr_t)a.v | (uintptr_t)a.dst) & 15) return 1;
    if ((a.yStride | a.uStride | a.vStride | a.dstStride) & 15) return 1;
    if (nframes > 1 && ((a.yFrame | a.uFrame | a.vFrame | a.dstFrame) & 15)) return 1;
    if (a.yStride <= 0 || a.uStride <= 0 || a.vStride <= 0) return 1;
    CUtensorMap tmY, tmU, tmV, tmD;
    {
        // the destination as 32-bit words: a 256-pixel row of a tile is 768 bytes = 192 words (a box side is at most 256 elements)
        const cuuint64_t dD[3] = { (cuuint64_t)dstW * 3 / 4, (cuuint64_t)(2 * rp1), (cuuint64_t)nframes };
        const cuuint64_t sD[2] = { (cuuint64_t)a.dstStride, nframes > 1 ? (cuuint64_t)a.dstFrame : (cuuint64_t)a.dstStride * dstH };
        const cuuint32_t bD[3] = { 192, 4, 1 };
        if (a.dstStride <= 0 || !tma_encode(&tmD, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, a.dst, dD, sD, bD)) return 1;
        const cuuint64_t dY[3] = { (cuuint64_t)dstW, (cuuint64_t)dstH, (cuuint64_t)nframes };
        const cuuint64_t dC[3] = { (cuuint64_t)chrSrcW, (cuuint64_t)chrSrcH, (cuuint64_t)nframes };
        const cuuint64_t sY[2] = { (cuuint64_t)a.yStride, nframes > 1 ? (cuuint64_t)a.yFrame : (cuuint64_t)a.yStride * dstH };
        const cuuint64_t sU[2] = { (cuuint64_t)a.uStride, nframes > 1 ? (cuuint64_t)a.uFrame : (cuuint64_t)a.uStride * chrSrcH };
        const cuuint64_t sV[2] = { (cuuint64_t)a.vStride, nframes > 1 ? (cuuint64_t)a.vFrame : (cuuint64_t)a.vStride * chrSrcH };
        const cuuint32_t bY[3] = { 256, 8, 1 }, bC[3] = { 128, 8, 1 };
        if (!tma_encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, a.y, dY, sY, bY) ||
            !tma_encode(&tmU, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, a.u, dC, sU, bC) ||
            !tma_encode(&tmV, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, a.v, dC, sV, bC)) return 1;
    }
    FusedTmaArgs fa;
    fa.dst = a.dst; fa.dstStride = a.dstStride; fa.dstFrame = a.dstFrame;
    fa.dstW = dstW; fa.chrSrcH = chrSrcH;
    fa.rp0 = rp0; fa.rowEnd = 2 * rp1;
    fa.tilesX = (dstW + 255) / 256; fa.nTy = (rp1 - rp0 + 3) / 4;
    const long long nitems = (long long)nframes * fa.nTy * fa.tilesX;
    if (nitems > 0x7fffffff / 2) return 1;
    fa.nitems = (int)nitems;
    // tuning knobs (profiling): sws_tma_warps 4 / 12 / 14, sws_tma_lut 2 = arithmetic terms instead of the shared-memory tables
    const bool lut = tuning("sws_tma_lut") != 2;
    int warps = tuning("sws_tma_warps");
    if (warps != 4 && warps != 12 && warps != 14) warps = nitems >= 12 * 2 * sm_count() ? 12 : 4;
    if (!lut && warps == 14) warps = 12;
    const int ctas_per_sm = warps == 4 ? 2 : 1;
    long long grid = (nitems + warps - 1) / warps;
    if (grid > (long long)sm_count() * ctas_per_sm) grid = (long long)sm_count() * ctas_per_sm;
    const long long wstride = grid * warps;
    fa.stepX = (int)(wstride % fa.tilesX);
    { const long long t2 = wstride / fa.tilesX; fa.stepTy = (int)(t2 % fa.nTy); fa.stepF = (int)(t2 / fa.nTy); }
#define AVB_FT_GO(L, W) (bgr ? launch_variant<true, L, W>(k, fa, taps, tmY, tmU, tmV, tmD, (int)grid, st) : launch_variant<false, L, W>(k, fa, taps, tmY, tmU, tmV, tmD, (int)grid, st))
    if (lut) return warps == 4 ? AVB_FT_GO(true, 4) : warps == 14 ? AVB_FT_GO(true, 14) : AVB_FT_GO(true, 12);
    return AVB_FT_GO(false, 12);
#undef AVB_FT_GO
}

}  // namespace avb
