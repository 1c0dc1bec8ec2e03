// libav_b200/csrc/slots_hbd.cu -- the 9 / 10-bit instances of H264DSPContext, H264QpelContext and H264ChromaContext as per-call slots:
// what ff_h264dsp_init(c, 9 | 10, idc), ff_h264qpel_init(c, 9 | 10) and ff_h264chroma_init(c, 9 | 10) install (libavcodec/h264dsp.c:57-143,
// h264qpel.c:36-89, h264chroma.c:32-55), with the C signatures (uint8_t * to 16-bit samples, int16_t * to int32 coefficients, strides in
// bytes) and HOST pointers.  Like slots.cu every call is a batch of one: stage the rectangles the C function touches, copy, one kernel,
// copy back, synchronise -- drop-in plumbing and parity for High 10 streams; there is no batched high-bit-depth path yet.
// The kernel's threads never communicate, so this file also compiles for tests/hostsim/ (CPU suite).
#include "h264dsp_hbd.cuh"
#include "scratch.h"
#include "../../include/avdsp_b200.h"
#include <string.h>

namespace avb {

constexpr int HSP = 32;           // pitch of every staged rectangle, in SAMPLES (64 bytes)
enum HbdOp { HOP_IDCT, HOP_IDCT_MB, HOP_DC, HOP_ADD_PIXELS, HOP_LOOP, HOP_QPEL, HOP_CHROMA_MC, HOP_WEIGHT };
struct HbdArgs {
    int op, bits, a, b, c, d, e, f, g;
    hbd::px *p0, *p1;
    int32_t *blk;
    int off[48];                  // block offsets in samples (idct_mb)
    uint8_t nnzc[120];
    int8_t tc0[4];
};
__device__ __forceinline__ int hbd_scan8(int i)
{
    const int plane = i >> 4, k = i & 15;
    return 4 + (k & 1) + 2 * ((k >> 2) & 1) + 8 * (1 + ((k >> 1) & 1) + 2 * (k >> 3) + 5 * plane);
}

__global__ void __launch_bounds__(32) slot_hbd_kernel(HbdArgs s)
{
    using namespace hbd;
    const int lane = threadIdx.x, bits = s.bits;
    switch (s.op) {
    case HOP_IDCT:                            // a = which (0 idct4, 1 idct8, 2 dc4, 3 dc8)
        if (lane == 0) {
            if (s.a == 0) idct4_add(bits, s.p0, s.blk, HSP);
            else if (s.a == 1) idct8_add(bits, s.p0, s.blk, HSP);
            else dc_add(bits, s.p0, s.blk, HSP, s.a == 2 ? 4 : 8);
        }
        break;
    case HOP_IDCT_MB: {                       // a = which (0 add16, 1 add16intra, 2 idct8_add4, 3 add8, 4 add8_422), b = pitch in samples
        const int which = s.a, pitch = s.b;
        if (which < 3 && lane < 16) {
            int32_t *blk = s.blk + 16 * lane;
            px *d = s.p0 + s.off[lane];
            const int nnz = s.nnzc[hbd_scan8(lane)];
            if (which == 0) { if (nnz) { if (nnz == 1 && blk[0]) dc_add(bits, d, blk, pitch, 4); else idct4_add(bits, d, blk, pitch); } }
            else if (which == 1) { if (nnz) idct4_add(bits, d, blk, pitch); else if (blk[0]) dc_add(bits, d, blk, pitch, 4); }
            else if ((lane & 3) == 0 && nnz) { if (nnz == 1 && blk[0]) dc_add(bits, d, blk, pitch, 8); else idct8_add(bits, d, blk, pitch); }
        } else if (which >= 3 && lane < (which == 4 ? 16 : 8)) {
            const int per = which == 4 ? 8 : 4, plane = lane / per, k = lane % per;
            const int i = 16 + 16 * plane + k, e = k >= 4 ? i + 4 : i;
            int32_t *blk = s.blk + 16 * i;
            px *d = (plane ? s.p1 : s.p0) + s.off[e];
            if (s.nnzc[hbd_scan8(e)]) idct4_add(bits, d, blk, pitch); else if (blk[0]) dc_add(bits, d, blk, pitch, 4);
        }
    } break;
    case HOP_DC:                              // a = qmul, b = kind (0 luma: blk = 256 out + 16 in, 1 chroma 4:2:0, 2 chroma 4:2:2)
        if (lane == 0) {
            if (s.b == 0) luma_dc_dequant(s.blk, s.blk + 256, s.a);
            else if (s.b == 1) chroma_dc_dequant(s.blk, s.a);
            else chroma422_dc_dequant(s.blk, s.a);
        }
        break;
    case HOP_ADD_PIXELS: {                    // a = n (4 / 8): dst += block (wraps in 16 bits, no clip), block cleared
        const int n = s.a;
        for (int i = lane; i < n * n; i += 32) { px *d = s.p0 + (i / n) * HSP + i % n; *d = (px)(*d + s.blk[i]); s.blk[i] = 0; }
    } break;
    case HOP_LOOP: {                          // a = kind bits (1 horizontal edge, 2 chroma, 4 intra), b = alpha, c = beta (both already << (bits - 8)), d = lines
        const bool horiz_edge = s.a & 1, chroma = s.a & 2, intra = s.a & 4;
        const int across = horiz_edge ? HSP : 1, along = horiz_edge ? 1 : HSP, lines = s.d, sh = bits - 8;
        if (lane < lines) {
            px *q = s.p0 + lane * along;
            const int t0 = s.tc0[lane / (lines >> 2)];
            if (!chroma) { if (intra) luma_intra_line(q, across, s.b, s.c); else if (t0 >= 0) luma_line(bits, q, across, s.b, s.c, t0 << sh); }   // h264dsp_template.c:113
            else if (intra) chroma_line(bits, q, across, s.b, s.c, 0, 1);
            else { const int tc = ((t0 - 1) << sh) + 1; if (tc > 0) chroma_line(bits, q, across, s.b, s.c, tc, 0); }                             // :240
        }
    } break;
    case HOP_QPEL:                            // a = avg, b = size, c = mc; p0 = dst, p1 = src (both pitch HSP)
        for (int i = lane; i < s.b * s.b; i += 32) {
            const int x = i % s.b, y = i / s.b, v = qpel_sample(bits, s.p1 + y * HSP + x, HSP, s.c & 3, s.c >> 2);
            px *d = s.p0 + y * HSP + x;
            *d = (px)(s.a ? (*d + v + 1) >> 1 : v);
        }
        break;
    case HOP_CHROMA_MC:                       // a = avg, b = w, c = h, d = x, e = y
        for (int i = lane; i < s.b * s.c; i += 32) {
            const int x = i % s.b, y = i / s.b, v = chroma_sample(s.p1 + y * HSP + x, HSP, s.d, s.e);
            px *d = s.p0 + y * HSP + x;
            *d = (px)(s.a ? (*d + v + 1) >> 1 : v);
        }
        break;
    case HOP_WEIGHT:                          // a = bi, b = w, c = h, d = log2_denom, e = weight (dst), f = weight (src), g = offset (already combined)
        for (int i = lane; i < s.b * s.c; i += 32) {
            px *d = s.p0 + (i / s.b) * HSP + i % s.b;
            if (!s.a) *d = (px)clipb((*d * s.e + s.g) >> s.d, bits);
            else *d = (px)clipb((s.p1[(i / s.b) * HSP + i % s.b] * s.f + *d * s.e + s.g) >> (s.d + 1), bits);
        }
        break;
    }
}

struct HStage {
    ScratchLock lk;
    uint8_t *h = nullptr, *d = nullptr;
    cudaStream_t s = nullptr;
    size_t used = 0;
    static constexpr size_t CAP = 128 * 1024;
    bool ok() {
        Scratch &S = scratch();
        h = (uint8_t *)S.pinned2(CAP); d = (uint8_t *)S.dev(7, CAP);
        cudaStream_t *st = S.streams();
        if (!h || !d || !st) return false;
        s = st[0];
        return true;
    }
    size_t take(size_t bytes) { size_t o = used; used += (bytes + 15) & ~(size_t)15; return o; }
    // rows of `w` SAMPLES -> pitch HSP samples
    size_t rect_in(const uint8_t *src, ptrdiff_t stride, int w, int hgt) {
        size_t o = take((size_t)HSP * 2 * hgt + HSP * 2);
        for (int y = 0; y < hgt; y++) memcpy(h + o + (size_t)y * HSP * 2, src + y * stride, (size_t)w * 2);
        return o;
    }
    void rect_out(uint8_t *dst, ptrdiff_t stride, int w, int hgt, size_t o) { for (int y = 0; y < hgt; y++) memcpy(dst + y * stride, h + o + (size_t)y * HSP * 2, (size_t)w * 2); }
    hbd::px *dp(size_t o) { return (hbd::px *)(d + o); }
    int go(const HbdArgs &a) {
        AVB_CUDA(cudaMemcpyAsync(d, h, used, cudaMemcpyHostToDevice, s), "hbd slot:h2d");
        AVB_LAUNCH(slot_hbd_kernel, 1, 32, 0, s)(a);
        if (check_launch("hbd slot")) return -1;
        AVB_CUDA(cudaMemcpyAsync(h, d, used, cudaMemcpyDeviceToHost, s), "hbd slot:d2h");
        AVB_CUDA(cudaStreamSynchronize(s), "hbd slot:sync");
        return 0;
    }
};

}  // namespace avb

using namespace avb;

namespace {

template <int BITS, int WHICH> void hslot_idct(uint8_t *dst, int16_t *block, int stride)
{
    HStage S; if (!S.ok()) return;
    const int n = (WHICH & 1) ? 8 : 4, coefs = WHICH == 1 ? 64 : WHICH == 0 ? 16 : 1;
    size_t od = S.rect_in(dst, stride, n, n), ob = S.take(256);
    memcpy(S.h + ob, block, (WHICH & 1) ? 256 : 64);
    HbdArgs a = {}; a.op = HOP_IDCT; a.bits = BITS; a.a = WHICH; a.p0 = S.dp(od); a.blk = (int32_t *)(S.d + ob);
    if (S.go(a)) return;
    S.rect_out(dst, stride, n, n, od);
    memcpy(block, S.h + ob, (size_t)coefs * 4);                // the transforms zero the block, dc_add only block[0]
}
// WHICH: 0 add16, 1 add16intra, 2 idct8_add4, 3 add8, 4 add8_422; offsets and stride in bytes
template <int BITS, int WHICH> void hidct_mb_impl(uint8_t *dst, uint8_t **dst2, const int *bo, int16_t *block, int stride, const uint8_t *nnzc)
{
    HStage S; if (!S.ok()) return;
    if (stride <= 0 || (stride & 1)) { set_error_msg("h264_idct_add16 slot (high bit depth)", "stride must be positive and even"); return; }
    constexpr int nb = WHICH == 2 ? 8 : 4, planes = WHICH >= 3 ? 2 : 1;
    constexpr int ncoef = (WHICH == 3 ? 36 : WHICH == 4 ? 40 : 16) * 16;
    int ent[16], n_ent = 0;
    if (WHICH <= 1) for (int i = 0; i < 16; i++) ent[n_ent++] = i;
    else if (WHICH == 2) for (int i = 0; i < 16; i += 4) ent[n_ent++] = i;
    else for (int p = 0; p < 2; p++) for (int k = 0; k < (WHICH == 4 ? 8 : 4); k++) ent[n_ent++] = 16 + 16 * p + (k >= 4 ? k + 4 : k);
    int W = 0, H = 0;                                          // bounding rectangle in samples / rows
    for (int n = 0; n < n_ent; n++) {
        const int i = ent[n];
        if (bo[i] < 0 || (bo[i] & 1)) { set_error_msg("h264_idct_add16 slot (high bit depth)", "block offsets must be non-negative and even"); return; }
        const int x = (bo[i] % stride) / 2 + nb, y = bo[i] / stride + nb;
        if (x > W) W = x;
        if (y > H) H = y;
    }
    const int pitch = (W + 7) & ~7;                            // samples
    if ((size_t)pitch * 2 * H * planes + 4096 > HStage::CAP) { set_error_msg("h264_idct_add16 slot (high bit depth)", "block offsets span too large a rectangle"); return; }
    size_t od[2];
    for (int p = 0; p < planes; p++) {
        od[p] = S.take((size_t)pitch * 2 * H);
        const uint8_t *src = WHICH >= 3 ? dst2[p] : dst;
        for (int y = 0; y < H; y++) memcpy(S.h + od[p] + (size_t)y * pitch * 2, src + (size_t)y * stride, (size_t)W * 2);
    }
    size_t ob = S.take(48 * 16 * 4);
    memcpy(S.h + ob, block, (size_t)ncoef * 4);
    HbdArgs a = {}; a.op = HOP_IDCT_MB; a.bits = BITS; a.a = WHICH; a.b = pitch; a.p0 = S.dp(od[0]); a.p1 = planes == 2 ? S.dp(od[1]) : nullptr;
    a.blk = (int32_t *)(S.d + ob);
    for (int n = 0; n < n_ent; n++) { const int i = ent[n]; a.off[i] = (bo[i] / stride) * pitch + (bo[i] % stride) / 2; }
    memcpy(a.nnzc, nnzc, 120);
    if (S.go(a)) return;
    for (int p = 0; p < planes; p++) {
        uint8_t *d = WHICH >= 3 ? dst2[p] : dst;
        for (int y = 0; y < H; y++) memcpy(d + (size_t)y * stride, S.h + od[p] + (size_t)y * pitch * 2, (size_t)W * 2);
    }
    memcpy(block, S.h + ob, (size_t)ncoef * 4);
}
template <int BITS, int WHICH> void hslot_idct_mb(uint8_t *dst, const int *bo, int16_t *block, int stride, const uint8_t nnzc[15 * 8]) { hidct_mb_impl<BITS, WHICH>(dst, nullptr, bo, block, stride, nnzc); }
template <int BITS, int WHICH> void hslot_idct_add8(uint8_t **dst, const int *bo, int16_t *block, int stride, const uint8_t nnzc[15 * 8]) { hidct_mb_impl<BITS, WHICH>(nullptr, dst, bo, block, stride, nnzc); }

template <int BITS> void hslot_luma_dc(int16_t *output, int16_t *input, int qmul)
{
    HStage S; if (!S.ok()) return;
    size_t o = S.take(1024 + 64);
    memcpy(S.h + o, output, 1024); memcpy(S.h + o + 1024, input, 64);        // 256 int32 of the macroblock's blocks, 16 int32 DC values
    HbdArgs a = {}; a.op = HOP_DC; a.bits = BITS; a.a = qmul; a.b = 0; a.blk = (int32_t *)(S.d + o);
    if (S.go(a)) return;
    memcpy(output, S.h + o, 1024);
}
template <int BITS, int C422> void hslot_chroma_dc(int16_t *block, int qmul)
{
    HStage S; if (!S.ok()) return;
    constexpr size_t bytes = (C422 ? 113 : 49) * 4;              // up to the last coefficient the C function touches
    size_t o = S.take(512);
    memcpy(S.h + o, block, bytes);
    HbdArgs a = {}; a.op = HOP_DC; a.bits = BITS; a.a = qmul; a.b = 1 + C422; a.blk = (int32_t *)(S.d + o);
    if (S.go(a)) return;
    memcpy(block, S.h + o, bytes);
}
template <int BITS, int N> void hslot_add_pixels_clear(uint8_t *dst, int16_t *block, int stride)
{
    HStage S; if (!S.ok()) return;
    size_t od = S.rect_in(dst, stride, N, N), ob = S.take(256);
    memcpy(S.h + ob, block, N * N * 4);
    HbdArgs a = {}; a.op = HOP_ADD_PIXELS; a.bits = BITS; a.a = N; a.p0 = S.dp(od); a.blk = (int32_t *)(S.d + ob);
    if (S.go(a)) return;
    S.rect_out(dst, stride, N, N, od);
    memcpy(block, S.h + ob, N * N * 4);
}
template <int BITS, int WIDX> void hslot_weight(uint8_t *block, int stride, int height, int log2_denom, int weight, int offset)
{
    HStage S; if (!S.ok()) return;
    const int w = 16 >> WIDX;
    size_t od = S.rect_in(block, stride, w, height);
    offset <<= log2_denom + (BITS - 8);                          // h264dsp_template.c:39-40
    if (log2_denom) offset += 1 << (log2_denom - 1);
    HbdArgs a = {}; a.op = HOP_WEIGHT; a.bits = BITS; a.a = 0; a.b = w; a.c = height; a.d = log2_denom; a.e = weight; a.g = offset; a.p0 = S.dp(od);
    if (S.go(a)) return;
    S.rect_out(block, stride, w, height, od);
}
template <int BITS, int WIDX> void hslot_biweight(uint8_t *dst, uint8_t *src, int stride, int height, int log2_denom, int weightd, int weights, int offset)
{
    HStage S; if (!S.ok()) return;
    const int w = 16 >> WIDX;
    size_t od = S.rect_in(dst, stride, w, height), os = S.rect_in(src, stride, w, height);
    offset <<= BITS - 8;                                         // :70-71
    offset = ((offset + 1) | 1) << log2_denom;
    HbdArgs a = {}; a.op = HOP_WEIGHT; a.bits = BITS; a.a = 1; a.b = w; a.c = height; a.d = log2_denom; a.e = weightd; a.f = weights; a.g = offset;
    a.p0 = S.dp(od); a.p1 = S.dp(os);
    if (S.go(a)) return;
    S.rect_out(dst, stride, w, height, od);
}
// WHICH: the numbering of slots.cu's loop filters (0..15)
template <int BITS, int WHICH> void hloop_impl(uint8_t *pix, int stride, int alpha, int beta, const int8_t *tc0)
{
    HStage S; if (!S.ok()) return;
    constexpr bool ext = WHICH >= 8;
    constexpr bool horiz_edge = !ext && !(WHICH & 1), chroma = ext ? WHICH >= 10 : WHICH >= 4, intra = ext ? (WHICH & 1) != 0 : (WHICH & 2) != 0;
    constexpr int ext_lines[8] = { 8, 8, 4, 4, 16, 16, 8, 8 };
    constexpr int lines = ext ? ext_lines[WHICH & 7] : chroma ? 8 : 16, reach = chroma ? 2 : 4;
    const int w = horiz_edge ? lines : 2 * reach, h = horiz_edge ? 2 * reach : lines;
    uint8_t *org = horiz_edge ? pix - reach * stride : pix - reach * 2;
    size_t o = S.rect_in(org, stride, w, h);
    HbdArgs a = {}; a.op = HOP_LOOP; a.bits = BITS; a.a = (horiz_edge ? 1 : 0) | (chroma ? 2 : 0) | (intra ? 4 : 0);
    a.b = alpha << (BITS - 8); a.c = beta << (BITS - 8); a.d = lines;                  // h264dsp_template.c:110-111
    a.p0 = S.dp(o) + (horiz_edge ? reach * HSP : reach);
    if (tc0) memcpy(a.tc0, tc0, 4);
    if (S.go(a)) return;
    S.rect_out(org, stride, w, h, o);
}
template <int BITS, int WHICH> void hslot_loop(uint8_t *pix, int stride, int alpha, int beta, int8_t *tc0) { hloop_impl<BITS, WHICH>(pix, stride, alpha, beta, tc0); }
template <int BITS, int WHICH> void hslot_loop_intra(uint8_t *pix, int stride, int alpha, int beta) { hloop_impl<BITS, WHICH>(pix, stride, alpha, beta, nullptr); }

template <int BITS, int AVG, int SIDX, int MC> void hslot_qpel(uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    HStage S; if (!S.ok()) return;
    const int n = 16 >> SIDX;
    size_t od = S.rect_in(dst, stride, n, n), os = S.rect_in(src - 2 * stride - 4, stride, n + 5, n + 5);
    HbdArgs a = {}; a.op = HOP_QPEL; a.bits = BITS; a.a = AVG; a.b = n; a.c = MC; a.p0 = S.dp(od); a.p1 = S.dp(os) + 2 * HSP + 2;
    if (S.go(a)) return;
    S.rect_out(dst, stride, n, n, od);
}
template <int BITS, int AVG, int WIDX> void hslot_chroma(uint8_t *dst, uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    HStage S; if (!S.ok()) return;
    const int w = 8 >> WIDX;
    size_t od = S.rect_in(dst, stride, w, h), os = S.rect_in(src, stride, w + (x != 0), h + (y != 0));      // like the C code, no read of a zero-weight column / row
    HbdArgs a = {}; a.op = HOP_CHROMA_MC; a.bits = BITS; a.a = AVG; a.b = w; a.c = h; a.d = x; a.e = y; a.p0 = S.dp(od); a.p1 = S.dp(os);
    if (S.go(a)) return;
    S.rect_out(dst, stride, w, h, od);
}

template <int BITS, int AVG, int SIDX> void hfill_qpel(qpel_mc_func *t)
{
#define Q(mc) t[mc] = hslot_qpel<BITS, AVG, SIDX, mc>;
    Q(0) Q(1) Q(2) Q(3) Q(4) Q(5) Q(6) Q(7) Q(8) Q(9) Q(10) Q(11) Q(12) Q(13) Q(14) Q(15)
#undef Q
}

template <int BITS> void init_dsp(H264DSPContext *c, int chroma_format_idc)
{
    const bool c420 = chroma_format_idc <= 1;                             // h264dsp.c:81-122
    c->weight_h264_pixels_tab[0] = hslot_weight<BITS, 0>; c->weight_h264_pixels_tab[1] = hslot_weight<BITS, 1>;
    c->weight_h264_pixels_tab[2] = hslot_weight<BITS, 2>; c->weight_h264_pixels_tab[3] = hslot_weight<BITS, 3>;
    c->biweight_h264_pixels_tab[0] = hslot_biweight<BITS, 0>; c->biweight_h264_pixels_tab[1] = hslot_biweight<BITS, 1>;
    c->biweight_h264_pixels_tab[2] = hslot_biweight<BITS, 2>; c->biweight_h264_pixels_tab[3] = hslot_biweight<BITS, 3>;
    c->h264_v_loop_filter_luma = hslot_loop<BITS, 0>; c->h264_h_loop_filter_luma = hslot_loop<BITS, 1>;
    c->h264_h_loop_filter_luma_mbaff = hslot_loop<BITS, 8>;
    c->h264_v_loop_filter_luma_intra = hslot_loop_intra<BITS, 2>; c->h264_h_loop_filter_luma_intra = hslot_loop_intra<BITS, 3>;
    c->h264_h_loop_filter_luma_mbaff_intra = hslot_loop_intra<BITS, 9>;
    c->h264_v_loop_filter_chroma = hslot_loop<BITS, 4>;
    c->h264_h_loop_filter_chroma = c420 ? hslot_loop<BITS, 5> : hslot_loop<BITS, 12>;
    c->h264_h_loop_filter_chroma_mbaff = c420 ? hslot_loop<BITS, 10> : hslot_loop<BITS, 14>;
    c->h264_v_loop_filter_chroma_intra = hslot_loop_intra<BITS, 6>;
    c->h264_h_loop_filter_chroma_intra = c420 ? hslot_loop_intra<BITS, 7> : hslot_loop_intra<BITS, 13>;
    c->h264_h_loop_filter_chroma_mbaff_intra = c420 ? hslot_loop_intra<BITS, 11> : hslot_loop_intra<BITS, 15>;
    c->h264_idct_add = hslot_idct<BITS, 0>; c->h264_idct8_add = hslot_idct<BITS, 1>;
    c->h264_idct_dc_add = hslot_idct<BITS, 2>; c->h264_idct8_dc_add = hslot_idct<BITS, 3>;
    c->h264_idct_add16 = hslot_idct_mb<BITS, 0>; c->h264_idct_add16intra = hslot_idct_mb<BITS, 1>; c->h264_idct8_add4 = hslot_idct_mb<BITS, 2>;
    c->h264_idct_add8 = c420 ? hslot_idct_add8<BITS, 3> : hslot_idct_add8<BITS, 4>;
    c->h264_luma_dc_dequant_idct = hslot_luma_dc<BITS>;
    c->h264_chroma_dc_dequant_idct = c420 ? hslot_chroma_dc<BITS, 0> : hslot_chroma_dc<BITS, 1>;
    c->h264_add_pixels8_clear = hslot_add_pixels_clear<BITS, 8>; c->h264_add_pixels4_clear = hslot_add_pixels_clear<BITS, 4>;
}
template <int BITS> void init_qpel(H264QpelContext *c)
{
    hfill_qpel<BITS, 0, 0>(c->put_h264_qpel_pixels_tab[0]); hfill_qpel<BITS, 0, 1>(c->put_h264_qpel_pixels_tab[1]);
    hfill_qpel<BITS, 0, 2>(c->put_h264_qpel_pixels_tab[2]); hfill_qpel<BITS, 0, 3>(c->put_h264_qpel_pixels_tab[3]);
    hfill_qpel<BITS, 1, 0>(c->avg_h264_qpel_pixels_tab[0]); hfill_qpel<BITS, 1, 1>(c->avg_h264_qpel_pixels_tab[1]);
    hfill_qpel<BITS, 1, 2>(c->avg_h264_qpel_pixels_tab[2]);              // no avg 2x2 row (h264qpel.c:60-67)
}
template <int BITS> void init_chroma(H264ChromaContext *c)
{
    c->put_h264_chroma_pixels_tab[0] = hslot_chroma<BITS, 0, 0>; c->put_h264_chroma_pixels_tab[1] = hslot_chroma<BITS, 0, 1>; c->put_h264_chroma_pixels_tab[2] = hslot_chroma<BITS, 0, 2>;
    c->avg_h264_chroma_pixels_tab[0] = hslot_chroma<BITS, 1, 0>; c->avg_h264_chroma_pixels_tab[1] = hslot_chroma<BITS, 1, 1>; c->avg_h264_chroma_pixels_tab[2] = hslot_chroma<BITS, 1, 2>;
}

}  // namespace

namespace avb {
// called by ff_h264dsp_init_cuda / ff_h264qpel_init_cuda / ff_h264chroma_init_cuda (slots.cu) for bit_depth 9 and 10
void h264dsp_init_hbd(H264DSPContext *c, int bits, int chroma_format_idc) { if (bits == 9) init_dsp<9>(c, chroma_format_idc); else init_dsp<10>(c, chroma_format_idc); }
void h264qpel_init_hbd(H264QpelContext *c, int bits) { if (bits == 9) init_qpel<9>(c); else init_qpel<10>(c); }
void h264chroma_init_hbd(H264ChromaContext *c, int bits) { if (bits == 9) init_chroma<9>(c); else init_chroma<10>(c); }
}  // namespace avb
