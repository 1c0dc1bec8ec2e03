// libav_b200/csrc/sws_slots.cu -- libswscale's per-line SwsContext slots (hyScale / hcScale, hyscale_fast / hcscale_fast, yuv2plane1 /
// yuv2planeX, yuv2nv12cX, yuv2packed1 / 2 / X) with the reference's signatures and HOST pointers, and ff_sws_init_swscale_cuda(), the
// arch hook that installs them (libswscale/swscale.c:723-783, swscale_internal.h:62-110,312-330,478-535).
//
// Like the DSP table slots (slots.cu) a call is a batch of one: the lines and coefficients the C function would read are staged into one
// pinned buffer, uploaded, one kernel evaluates the line (a thread per output sample / pixel pair, the same arithmetic as the frame
// kernels through sws_dev.cuh), the bytes the C function would write come back, and the call synchronises.  This is the drop-in /
// parity layer of boundary B; throughput comes from sws_scale_frames_cuda().  The kernel's threads never communicate, so the whole
// file also compiles for tests/hostsim/ (CPU suite).
#include "sws_dev.cuh"
#include "scratch.h"
#include "../../include/avdsp_b200.h"
#include <mutex>
#include <vector>
#include <utility>
#include <string.h>

namespace avb {

enum SwsLineOp { OP_HSCALE, OP_HFAST, OP_PLANE, OP_NV12, OP_PACKED };

struct SwsLineArgs {
    int op;
    int dstW;                        // output samples (pixels for the packed stage)
    int nl, nc;                      // lines per luma / chroma set (tap counts)
    int a, b, c, d;                  // op-specific
    const uint8_t *src8, *src8b;     // 8-bit source lines
    const int16_t *lf, *cf;          // coefficients
    const int32_t *pos;
    const uint8_t *lum, *chrU, *chrV;    // staged line sets: line j at + j * stride bytes
    int lumStride, chrStride;
    uint8_t *dst, *dst2;
    uint8_t dither[8];
    RgbConstants k;
};

__device__ __forceinline__ int line_at(const uint8_t *set, int stride, int j, int i, bool wide)
{
    return wide ? reinterpret_cast<const int32_t *>(set + (size_t)j * stride)[i] : reinterpret_cast<const int16_t *>(set + (size_t)j * stride)[i];
}

// packed stores: target 0 rgb24, 1 bgr24, 4 argb, 5 rgba, 6 abgr, 7 bgra (alpha 255: the contexts taken over have no alpha plane)
__device__ __forceinline__ void store_rgb(uint8_t *d, int target, int i, int R, int G, int B)
{
    if (target < 4) { d += 3 * (size_t)i; d[target ? 2 : 0] = (uint8_t)R; d[1] = (uint8_t)G; d[target ? 0 : 2] = (uint8_t)B; return; }
    d += 4 * (size_t)i;
    const int ro = target == 4 ? 1 : target == 5 ? 0 : target == 6 ? 3 : 2, bo = target == 4 ? 3 : target == 5 ? 2 : target == 6 ? 1 : 0;
    d[ro] = (uint8_t)R; d[target <= 5 ? ro + 1 : ro - 1] = (uint8_t)G; d[bo] = (uint8_t)B; d[(target & 1) ? 3 : 0] = 255;
}

__global__ void __launch_bounds__(256) sws_line_kernel(SwsLineArgs s)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    switch (s.op) {
    case OP_HSCALE: {                         // hScale8To15_c / hScale8To19_c (swscale.c:133-164); a = 19 for the 19-bit variant; b = first staged source byte
        if (i >= s.dstW) return;
        const uint8_t *sp = s.src8 + (s.pos[i] - s.b);
        const int16_t *f = s.lf + (size_t)i * s.nl;
        int val = 0;
        for (int j = 0; j < s.nl; j++) val += (int)sp[j] * f[j];
        if (s.a == 19) reinterpret_cast<int32_t *>(s.dst)[i] = min(val >> 3, (1 << 19) - 1);
        else           reinterpret_cast<int16_t *>(s.dst)[i] = (int16_t)min(val >> 7, (1 << 15) - 1);
    } break;
    case OP_HFAST: {                          // hyscale_fast_c / hcscale_fast_c (swscale.c:238-250, 286-299); a = xInc, b = chroma
        if (i >= s.dstW) return;
        const unsigned xpos = (unsigned)i * (unsigned)s.a, xx = xpos >> 16, xa = (xpos & 0xFFFF) >> 9;
        if (!s.b) reinterpret_cast<int16_t *>(s.dst)[i] = (int16_t)((s.src8[xx] << 7) + (s.src8[xx + 1] - s.src8[xx]) * (int)xa);
        else {
            reinterpret_cast<int16_t *>(s.dst)[i]  = (int16_t)(s.src8[xx] * (int)(xa ^ 127) + s.src8[xx + 1] * (int)xa);
            reinterpret_cast<int16_t *>(s.dst2)[i] = (int16_t)(s.src8b[xx] * (int)(xa ^ 127) + s.src8b[xx + 1] * (int)xa);
        }
    } break;
    case OP_PLANE: {                          // yuv2plane1 / yuv2planeX (output.c:136-265): nl = 0 plane1 else taps; a = bits, b = big endian, c = dither offset
        if (i >= s.dstW) return;
        const int bits = s.a;
        int v;
        if (bits == 8) {
            const int dth = s.dither[(i + s.c) & 7];
            if (!s.nl) v = clip_u8((line_at(s.lum, 0, 0, i, false) + dth) >> 7);
            else {
                int val = dth << 12;
                for (int j = 0; j < s.nl; j++) val += line_at(s.lum, s.lumStride, j, i, false) * s.lf[j];
                v = clip_u8(val >> 19);
            }
            s.dst[i] = (uint8_t)v;
        } else if (bits == 16) {              // 19-bit int32 lines; planeX biases the sum so that it stays inside 32 bits (wrap-around arithmetic here)
            if (!s.nl) v = min(max((line_at(s.lum, 0, 0, i, true) + 4) >> 3, 0), 65535);
            else {
                unsigned acc = (1u << 14) - 0x40000000u;
                for (int j = 0; j < s.nl; j++) acc += (unsigned)(line_at(s.lum, s.lumStride, j, i, true) * s.lf[j]);
                v = min(max((int)acc >> 15, -32768), 32767) + 0x8000;
            }
            reinterpret_cast<uint16_t *>(s.dst)[i] = (uint16_t)swap16_if(v, s.b);
        } else {                              // 9 / 10 bit
            if (!s.nl) v = (line_at(s.lum, 0, 0, i, false) + (1 << (14 - bits))) >> (15 - bits);
            else {
                int val = 1 << (26 - bits);
                for (int j = 0; j < s.nl; j++) val += line_at(s.lum, s.lumStride, j, i, false) * s.lf[j];
                v = val >> (27 - bits);
            }
            reinterpret_cast<uint16_t *>(s.dst)[i] = (uint16_t)swap16_if(plane_clip(v, bits), s.b);
        }
    } break;
    case OP_NV12: {                           // yuv2nv12cX_c (output.c:267-303); a = 1 for nv21; dither = chrDither8
        if (i >= s.dstW) return;
        int u = s.dither[i & 7] << 12, v = s.dither[(i + 3) & 7] << 12;
        for (int j = 0; j < s.nc; j++) { u += line_at(s.chrU, s.chrStride, j, i, false) * s.cf[j]; v += line_at(s.chrV, s.chrStride, j, i, false) * s.cf[j]; }
        s.dst[2 * i + (s.a ? 1 : 0)] = (uint8_t)clip_u8(u >> 19);
        s.dst[2 * i + (s.a ? 0 : 1)] = (uint8_t)clip_u8(v >> 19);
    } break;
    case OP_PACKED: {                         // a = kind (1, 2, 0 = X), b = target, c = SWS_FULL_CHR_H_INT, d = uvalpha (kind 1)
        const int kind = s.a, target = s.b;
        if (s.c) {                            // yuv2rgb_full_X_c_template (output.c:1165-1250): a thread per pixel
            if (i >= s.dstW) return;
            int Y = 0, U = -128 * (1 << 19), V = -128 * (1 << 19);
            for (int j = 0; j < s.nl; j++) Y += line_at(s.lum, s.lumStride, j, i, false) * s.lf[j];
            for (int j = 0; j < s.nc; j++) { U += line_at(s.chrU, s.chrStride, j, i, false) * s.cf[j]; V += line_at(s.chrV, s.chrStride, j, i, false) * s.cf[j]; }
            uint8_t px[3];
            full_pixel(Y >> 10, U >> 10, V >> 10, s.k, 0, px);
            store_rgb(s.dst, target, i, px[0], px[1], px[2]);
            return;
        }
        if (i >= ((s.dstW + 1) >> 1)) return; // a thread per pixel pair; the pair is always written whole, like the C loops
        auto L = [&](int j, int x) { return line_at(s.lum, s.lumStride, j, x, false); };
        auto CU = [&](int j) { return line_at(s.chrU, s.chrStride, j, i, false); };
        auto CV = [&](int j) { return line_at(s.chrV, s.chrStride, j, i, false); };
        int Y1, Y2, U, V;
        if (kind == 1) {                      // yuv2rgb_1 / yuv2422_1 (output.c:1042-1110, 531-576)
            Y1 = L(0, 2 * i) >> 7; Y2 = L(0, 2 * i + 1) >> 7;
            if (s.d < 2048) { U = CU(0) >> 7; V = CV(0) >> 7; }
            else            { U = (CU(0) + CU(1)) >> 8; V = (CV(0) + CV(1)) >> 8; }
            Y1 = clip_u8(Y1); Y2 = clip_u8(Y2); U = clip_u8(U); V = clip_u8(V);
        } else if (kind == 2) {               // yuv2rgb_2 / yuv2422_2 (output.c:997-1040, 498-529): lf = { 4096 - yalpha, yalpha }, cf likewise
            Y1 = (L(0, 2 * i) * s.lf[0] + L(1, 2 * i) * s.lf[1]) >> 19;
            Y2 = (L(0, 2 * i + 1) * s.lf[0] + L(1, 2 * i + 1) * s.lf[1]) >> 19;
            U = (CU(0) * s.cf[0] + CU(1) * s.cf[1]) >> 19;
            V = (CV(0) * s.cf[0] + CV(1) * s.cf[1]) >> 19;
            Y1 = clip_u8(Y1); Y2 = clip_u8(Y2); U = clip_u8(U); V = clip_u8(V);
        } else {                              // yuv2rgb_X / yuv2422_X (output.c:936-995, 456-496)
            Y1 = Y2 = U = V = 1 << 18;
            for (int j = 0; j < s.nl; j++) { Y1 += L(j, 2 * i) * s.lf[j]; Y2 += L(j, 2 * i + 1) * s.lf[j]; }
            for (int j = 0; j < s.nc; j++) { U += CU(j) * s.cf[j]; V += CV(j) * s.cf[j]; }
            Y1 >>= 19; Y2 >>= 19; U >>= 19; V >>= 19;
            clip_if_flagged(Y1, Y2, U, V);
        }
        if (target == 2 || target == 3) {     // output_pixels (output.c:448-467)
            uint8_t *d = s.dst + 4 * (size_t)i;
            if (target == 2) { d[0] = (uint8_t)Y1; d[1] = (uint8_t)U; d[2] = (uint8_t)Y2; d[3] = (uint8_t)V; }
            else             { d[0] = (uint8_t)U; d[1] = (uint8_t)Y1; d[2] = (uint8_t)V; d[3] = (uint8_t)Y2; }
            return;
        }
        const ChromaTerms t = chroma_terms(U, V, s.k);
        store_rgb(s.dst, target, 2 * i,     clip_u8((s.k.cy * Y1 + t.tr) >> 16), clip_u8((s.k.cy * Y1 + t.tg) >> 16), clip_u8((s.k.cy * Y1 + t.tb) >> 16));
        store_rgb(s.dst, target, 2 * i + 1, clip_u8((s.k.cy * Y2 + t.tr) >> 16), clip_u8((s.k.cy * Y2 + t.tg) >> 16), clip_u8((s.k.cy * Y2 + t.tb) >> 16));
    } break;
    }
}

// lumRangeFromJpeg_c / chrRangeFromJpeg_c / lumRangeToJpeg_c / chrRangeToJpeg_c (swscale.c:166-197), in place on 15-bit lines:
// the frame path runs it over its hscaled line planes, the lumConvertRange / chrConvertRange slots over one (two) staged line(s)
__device__ __forceinline__ int range_sample(int v, int kind)
{
    switch (kind) {
    case 0: return (v * 14071 + 33561947) >> 14;
    case 1: return (v * 1799 + 4081085) >> 11;
    case 2: return (min(v, 30189) * 19077 - 39057361) >> 14;
    default: return (min(v, 30775) * 4663 - 9289992) >> 12;
    }
}
__global__ void __launch_bounds__(256) sws_range_kernel(int16_t *plane, int stridePx, int w, int rows, int kind)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= rows) return;
    int16_t *p = plane + (size_t)y * stridePx + x;
    *p = (int16_t)range_sample(*p, kind);
}
int sws_launch_range(int16_t *plane, int stridePx, int w, int rows, int kind, cudaStream_t stream)
{
    if (w <= 0 || rows <= 0) return 0;
    AVB_LAUNCH(sws_range_kernel, dim3((w + 255) / 256, rows), dim3(256), 0, stream)(plane, stridePx, w, rows, kind);
    return check_launch("sws range conversion");
}

// ---- registry: struct SwsContext * (opaque key) -> the context made by sws_getContext_cuda ------------------
static std::mutex g_reg_mu;
static std::vector<std::pair<const void *, const void *>> g_reg;

void sws_slots_forget(const void *ctx)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (size_t n = g_reg.size(); n-- > 0;) if (g_reg[n].second == ctx) g_reg.erase(g_reg.begin() + n);
}
static bool view_of(const void *key, SwsSlotView &v, const char *who)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (auto &e : g_reg) if (e.first == key) return sws_slot_view(e.second, v);
    set_error_msg(who, "this SwsContext was not registered with ff_sws_init_swscale_cuda");
    return false;
}

// ---- staging: one pinned buffer and one device buffer sized for the call ---------------------------------------
struct LineStage {
    ScratchLock lk;
    uint8_t *h = nullptr, *d = nullptr;
    cudaStream_t s = nullptr;
    size_t used = 0, cap = 0;
    bool ok(size_t bytes) {
        Scratch &S = scratch();
        cap = bytes + 4096;
        h = (uint8_t *)S.pinned2(cap); d = (uint8_t *)S.dev(6, cap);
        cudaStream_t *st = S.streams();
        if (!h || !d || !st) return false;
        s = st[0];
        return true;
    }
    size_t take(size_t bytes) { size_t o = used; used += (bytes + 15) & ~(size_t)15; return o; }
    size_t put(const void *src, size_t bytes) { size_t o = take(bytes); memcpy(h + o, src, bytes); return o; }
    // `n` lines of `bytes` each, packed `stride` apart
    size_t put_lines(const int16_t *const *lines, int n, size_t bytes, size_t stride) {
        size_t o = take(stride * (size_t)(n > 0 ? n : 1));
        for (int j = 0; j < n; j++) memcpy(h + o + j * stride, lines[j], bytes);
        return o;
    }
    int run(const SwsLineArgs &a, int threads, size_t out_off, size_t out_bytes, size_t out2_off = 0, size_t out2_bytes = 0) {
        if (used > cap) { set_error_msg("sws line slot", "internal: staging overflow"); return -1; }
        AVB_CUDA(cudaMemcpyAsync(d, h, used, cudaMemcpyHostToDevice, s), "sws line slot:h2d");
        if (threads > 0) AVB_LAUNCH(sws_line_kernel, dim3((threads + 255) / 256), dim3(256), 0, s)(a);
        if (check_launch("sws line slot")) return -1;
        AVB_CUDA(cudaMemcpyAsync(h + out_off, d + out_off, out_bytes, cudaMemcpyDeviceToHost, s), "sws line slot:d2h");
        if (out2_bytes) AVB_CUDA(cudaMemcpyAsync(h + out2_off, d + out2_off, out2_bytes, cudaMemcpyDeviceToHost, s), "sws line slot:d2h");
        AVB_CUDA(cudaStreamSynchronize(s), "sws line slot:sync");
        return 0;
    }
};
static inline size_t al16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace avb

using namespace avb;

namespace {

// ---- horizontal ----
template <int BITS> void slot_hscale(struct SwsContext *, int16_t *dst, int dstW, const uint8_t *src, const int16_t *filter, const int32_t *filterPos, int filterSize)
{
    if (dstW <= 0 || filterSize <= 0) return;
    int lo = filterPos[0], hi = filterPos[0];
    for (int i = 1; i < dstW; i++) { lo = filterPos[i] < lo ? filterPos[i] : lo; hi = filterPos[i] > hi ? filterPos[i] : hi; }
    const size_t nsrc = (size_t)(hi - lo) + filterSize, nf = (size_t)dstW * filterSize * 2, np = (size_t)dstW * 4, nout = (size_t)dstW * (BITS == 19 ? 4 : 2);
    LineStage S; if (!S.ok(al16(nsrc) + al16(nf) + al16(np) + al16(nout))) return;
    SwsLineArgs a = {}; a.op = OP_HSCALE; a.dstW = dstW; a.nl = filterSize; a.a = BITS; a.b = lo;
    a.src8 = S.d + S.put(src + lo, nsrc); a.lf = (const int16_t *)(S.d + S.put(filter, nf)); a.pos = (const int32_t *)(S.d + S.put(filterPos, np));
    const size_t oo = S.take(nout); a.dst = S.d + oo;
    if (S.run(a, dstW, oo, nout)) return;
    memcpy(dst, S.h + oo, nout);
}
// the bytes the C loops read: src[0 .. ((dstWidth - 1) * xInc >> 16) + 1]
void slot_hyscale_fast(struct SwsContext *, int16_t *dst, int dstWidth, const uint8_t *src, int srcW, int xInc)
{
    (void)srcW;
    if (dstWidth <= 0) return;
    const size_t nsrc = (((unsigned)(dstWidth - 1) * (unsigned)xInc) >> 16) + 2, nout = (size_t)dstWidth * 2;
    LineStage S; if (!S.ok(al16(nsrc) + al16(nout))) return;
    SwsLineArgs a = {}; a.op = OP_HFAST; a.dstW = dstWidth; a.a = xInc; a.b = 0;
    a.src8 = S.d + S.put(src, nsrc);
    const size_t oo = S.take(nout); a.dst = S.d + oo;
    if (S.run(a, dstWidth, oo, nout)) return;
    memcpy(dst, S.h + oo, nout);
}
void slot_hcscale_fast(struct SwsContext *, int16_t *dst1, int16_t *dst2, int dstWidth, const uint8_t *src1, const uint8_t *src2, int srcW, int xInc)
{
    (void)srcW;
    if (dstWidth <= 0) return;
    const size_t nsrc = (((unsigned)(dstWidth - 1) * (unsigned)xInc) >> 16) + 2, nout = (size_t)dstWidth * 2;
    LineStage S; if (!S.ok(2 * al16(nsrc) + 2 * al16(nout))) return;
    SwsLineArgs a = {}; a.op = OP_HFAST; a.dstW = dstWidth; a.a = xInc; a.b = 1;
    a.src8 = S.d + S.put(src1, nsrc); a.src8b = S.d + S.put(src2, nsrc);
    const size_t o1 = S.take(nout), o2 = S.take(nout); a.dst = S.d + o1; a.dst2 = S.d + o2;
    if (S.run(a, dstWidth, o1, nout, o2, nout)) return;
    memcpy(dst1, S.h + o1, nout); memcpy(dst2, S.h + o2, nout);
}

// ---- planar output ----
template <int BITS, int BE> void plane_impl(const int16_t *filter, int fs, const int16_t *const *src, uint8_t *dest, int dstW, const uint8_t *dither, int offset)
{
    if (dstW <= 0) return;
    const size_t sample = BITS == 16 ? 4 : 2, line = (size_t)dstW * sample, stride = al16(line), nout = (size_t)dstW * (BITS == 8 ? 1 : 2);
    const int n = fs ? fs : 1;
    LineStage S; if (!S.ok(stride * n + al16((size_t)n * 2) + al16(nout))) return;
    SwsLineArgs a = {}; a.op = OP_PLANE; a.dstW = dstW; a.nl = fs; a.a = BITS; a.b = BE; a.c = offset;
    if (BITS == 8) { if (!dither) { set_error_msg("yuv2plane slot", "NULL dither"); return; } memcpy(a.dither, dither, 8); }
    a.lum = S.d + S.put_lines(src, n, line, stride); a.lumStride = (int)stride;
    if (fs) a.lf = (const int16_t *)(S.d + S.put(filter, (size_t)fs * 2));
    const size_t oo = S.take(nout); a.dst = S.d + oo;
    if (S.run(a, dstW, oo, nout)) return;
    memcpy(dest, S.h + oo, nout);
}
template <int BITS, int BE> void slot_plane1(const int16_t *src, uint8_t *dest, int dstW, const uint8_t *dither, int offset)
{ const int16_t *one[1] = { src }; plane_impl<BITS, BE>(nullptr, 0, one, dest, dstW, dither, offset); }
template <int BITS, int BE> void slot_planeX(const int16_t *filter, int filterSize, const int16_t **src, uint8_t *dest, int dstW, const uint8_t *dither, int offset)
{ if (filterSize > 0) plane_impl<BITS, BE>(filter, filterSize, src, dest, dstW, dither, offset); }

void slot_nv12cX(struct SwsContext *c, const int16_t *chrFilter, int chrFilterSize, const int16_t **chrUSrc, const int16_t **chrVSrc, uint8_t *dest, int chrDstW)
{
    SwsSlotView v;
    if (!view_of(c, v, "yuv2nv12cX slot") || chrDstW <= 0 || chrFilterSize <= 0) return;
    const size_t line = (size_t)chrDstW * 2, stride = al16(line), nout = (size_t)chrDstW * 2;
    LineStage S; if (!S.ok(2 * stride * chrFilterSize + al16((size_t)chrFilterSize * 2) + al16(nout))) return;
    SwsLineArgs a = {}; a.op = OP_NV12; a.dstW = chrDstW; a.nc = chrFilterSize; a.a = v.dstNV == 2;
    memset(a.dither, 64, 8);                 // c->chrDither8 = ff_sws_pb_64 for every 8-bit source (swscale.c:413-416,445-447)
    a.chrU = S.d + S.put_lines(chrUSrc, chrFilterSize, line, stride); a.chrV = S.d + S.put_lines(chrVSrc, chrFilterSize, line, stride); a.chrStride = (int)stride;
    a.cf = (const int16_t *)(S.d + S.put(chrFilter, (size_t)chrFilterSize * 2));
    const size_t oo = S.take(nout); a.dst = S.d + oo;
    if (S.run(a, chrDstW, oo, nout)) return;
    memcpy(dest, S.h + oo, nout);
}

// ---- packed output: kind 1 / 2 / 0 (X) ----
void packed_impl(struct SwsContext *c, int kind, const int16_t *lumFilter, const int16_t *const *lumSrc, int nl, const int16_t *chrFilter,
                 const int16_t *const *chrUSrc, const int16_t *const *chrVSrc, int nc, uint8_t *dest, int dstW, int uvalpha)
{
    SwsSlotView v;
    if (!view_of(c, v, "yuv2packed slot") || dstW <= 0) return;
    if (v.target < 0) { set_error_msg("yuv2packed slot", "the context has a planar destination"); return; }
    const bool full = (v.flags & SWS_FULL_CHR_H_INT) && v.target != 2 && v.target != 3;
    if (full && kind) { set_error_msg("yuv2packed slot", "SWS_FULL_CHR_H_INT only has yuv2packedX"); return; }
    const int pairs = (dstW + 1) >> 1;
    const int lumN = full ? dstW : 2 * pairs, chrN = full ? dstW : pairs;           // samples the C loops read per line
    const int px = full ? dstW : 2 * pairs, bpp = v.target == 2 || v.target == 3 ? 2 : v.target >= 4 ? 4 : 3;
    const size_t lline = (size_t)lumN * 2, lstride = al16(lline), cline = (size_t)chrN * 2, cstride = al16(cline), nout = (size_t)px * bpp;
    LineStage S; if (!S.ok(lstride * nl + 2 * cstride * nc + al16((size_t)nl * 2) + al16((size_t)nc * 2) + al16(nout))) return;
    SwsLineArgs a = {}; a.op = OP_PACKED; a.dstW = dstW; a.nl = nl; a.nc = nc; a.a = kind; a.b = v.target; a.c = full; a.d = uvalpha; a.k = v.k;
    a.lum = S.d + S.put_lines(lumSrc, nl, lline, lstride); a.lumStride = (int)lstride;
    a.chrU = S.d + S.put_lines(chrUSrc, nc, cline, cstride); a.chrV = S.d + S.put_lines(chrVSrc, nc, cline, cstride); a.chrStride = (int)cstride;
    if (lumFilter) a.lf = (const int16_t *)(S.d + S.put(lumFilter, (size_t)nl * 2));
    if (chrFilter) a.cf = (const int16_t *)(S.d + S.put(chrFilter, (size_t)nc * 2));
    const size_t oo = S.take(nout); a.dst = S.d + oo;
    if (S.run(a, full ? dstW : pairs, oo, nout)) return;
    memcpy(dest, S.h + oo, nout);
}
void slot_packed1(struct SwsContext *c, const int16_t *lumSrc, const int16_t *chrUSrc[2], const int16_t *chrVSrc[2], const int16_t *, uint8_t *dest, int dstW, int uvalpha, int)
{
    const int16_t *l[1] = { lumSrc };
    packed_impl(c, 1, nullptr, l, 1, nullptr, chrUSrc, chrVSrc, uvalpha < 2048 ? 1 : 2, dest, dstW, uvalpha);     // the second chroma line is only read from 2048 up
}
void slot_packed2(struct SwsContext *c, const int16_t *lumSrc[2], const int16_t *chrUSrc[2], const int16_t *chrVSrc[2], const int16_t *[2], uint8_t *dest, int dstW, int yalpha, int uvalpha, int)
{
    const int16_t lf[2] = { (int16_t)(4096 - yalpha), (int16_t)yalpha }, cf[2] = { (int16_t)(4096 - uvalpha), (int16_t)uvalpha };
    packed_impl(c, 2, lf, lumSrc, 2, cf, chrUSrc, chrVSrc, 2, dest, dstW, uvalpha);
}
void slot_packedX(struct SwsContext *c, const int16_t *lumFilter, const int16_t **lumSrc, int lumFilterSize, const int16_t *chrFilter, const int16_t **chrUSrc,
                  const int16_t **chrVSrc, int chrFilterSize, const int16_t **, uint8_t *dest, int dstW, int)
{
    if (lumFilterSize <= 0 || chrFilterSize <= 0) return;
    packed_impl(c, 0, lumFilter, lumSrc, lumFilterSize, chrFilter, chrUSrc, chrVSrc, chrFilterSize, dest, dstW, 0);
}

// ---- range conversion: c->lumConvertRange(dst, width) / c->chrConvertRange(dstU, dstV, width), in place ----
template <int KIND> void range_impl(int16_t *d1, int16_t *d2, int width)
{
    if (width <= 0) return;
    const size_t line = (size_t)width * 2, stride = al16(line);
    LineStage S; if (!S.ok(2 * stride)) return;
    const size_t o1 = S.put(d1, line), o2 = d2 ? S.put(d2, line) : o1;
    if (S.used > S.cap) return;
    if (cudaMemcpyAsync(S.d, S.h, S.used, cudaMemcpyHostToDevice, S.s) != cudaSuccess) { set_error("sws range slot:h2d", cudaGetLastError()); return; }
    if (sws_launch_range((int16_t *)(S.d + o1), (int)(stride / 2), width, d2 ? 2 : 1, KIND, S.s)) return;
    if (cudaMemcpyAsync(S.h, S.d, S.used, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) { set_error("sws range slot:d2h", cudaGetLastError()); return; }
    memcpy(d1, S.h + o1, line);
    if (d2) memcpy(d2, S.h + o2, line);
}
template <int KIND> void slot_lum_range(int16_t *dst, int width) { range_impl<KIND>(dst, nullptr, width); }
template <int KIND> void slot_chr_range(int16_t *dstU, int16_t *dstV, int width) { range_impl<KIND>(dstU, dstV, width); }

}  // namespace

extern "C" int ff_sws_init_swscale_cuda(struct SwsContext *c, SwsContextCUDA *cuda, SwsLineSlotsCUDA *t)
{
    avb::enter();
    SwsSlotView v;
    if (!c || !t || !sws_slot_view(cuda, v)) { set_error_msg("ff_sws_init_swscale_cuda", "NULL argument"); return -1; }
    {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        bool found = false;
        for (auto &e : g_reg) if (e.first == c) { e.second = cuda; found = true; }
        if (!found) g_reg.emplace_back(c, cuda);
    }
    memset(t, 0, sizeof(*t));
    // 9 / 10 / 16-bit sources: the reference installs hScale16To15_c / hScale16To19_c (swscale.c:744-745) and feeds the 8-bit output
    // functions an ordered-dither row per output line that lives in its own SwsContext (chrDither8 / lumDither8, swscale.c:554-555),
    // which this hook cannot see.  Those contexts are NOT taken over line by line: every slot stays NULL (the caller keeps its C
    // functions) and the refusal is reported; the whole-frame calls sws_scale_cuda / sws_scale_frames_cuda serve them.
    if (v.srcBits > 8) { set_error_msg("ff_sws_init_swscale_cuda", "high-bit-depth source: the per-line slots are not taken over (use the frame calls)"); return -1; }
    const bool d16 = v.planar && v.dstBits == 16;
    t->hyScale = t->hcScale = d16 ? slot_hscale<19> : slot_hscale<15>;                    // swscale.c:733-743
    if ((v.flags & SWS_FAST_BILINEAR) && !d16) { t->hyscale_fast = slot_hyscale_fast; t->hcscale_fast = slot_hcscale_fast; }
    const int bits = v.planar ? v.dstBits : 8, be = v.planar ? v.dstBE : 0;               // output.c:1369-1390
    if (bits == 16)      { t->yuv2plane1 = be ? slot_plane1<16, 1> : slot_plane1<16, 0>; t->yuv2planeX = be ? slot_planeX<16, 1> : slot_planeX<16, 0>; }
    else if (bits == 10) { t->yuv2plane1 = be ? slot_plane1<10, 1> : slot_plane1<10, 0>; t->yuv2planeX = be ? slot_planeX<10, 1> : slot_planeX<10, 0>; }
    else if (bits == 9)  { t->yuv2plane1 = be ? slot_plane1<9, 1> : slot_plane1<9, 0>;   t->yuv2planeX = be ? slot_planeX<9, 1> : slot_planeX<9, 0>; }
    else                 { t->yuv2plane1 = slot_plane1<8, 0>; t->yuv2planeX = slot_planeX<8, 0>; if (v.dstNV) t->yuv2nv12cX = slot_nv12cX; }
    if (v.target >= 0) {                                                                  // output.c:1392-1580
        const bool full = (v.flags & SWS_FULL_CHR_H_INT) && v.target != 2 && v.target != 3;
        t->yuv2packedX = slot_packedX;
        if (!full) { t->yuv2packed1 = slot_packed1; t->yuv2packed2 = slot_packed2; }
    }
    if (v.rangeConv == 1)      { t->lumConvertRange = slot_lum_range<0>; t->chrConvertRange = slot_chr_range<1>; }     // swscale.c:748-757
    else if (v.rangeConv == 2) { t->lumConvertRange = slot_lum_range<2>; t->chrConvertRange = slot_chr_range<3>; }
    return 0;
}
