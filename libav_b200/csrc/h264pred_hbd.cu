// libav_b200/csrc/h264pred_hbd.cu -- the 9 / 10-bit instances of H264PredContext (libavcodec/h264pred.h:91-110; h264pred_template.c with
// BIT_DEPTH > 8: uint16 samples, int32 residual in the lossless *_add functions) for codec H.264, 4:2:0, as per-call table slots.
// Same edge-array formulas as h264pred.cuh (its directional expression is reused as is); what the depth changes: the DC_128 family
// predicts 1 << (bits - 1), plane prediction clips to `bits` bits, the lossless running sums wrap in 16 bits.  A slot gathers the
// neighbours the C function of that mode reads (host pointers, strides in bytes), runs one small kernel and scatters the block back.
// Every thread derives its sample from the raw neighbours on its own (no shared memory, no barrier): the file also compiles for
// tests/hostsim/.  There is no batched high-bit-depth intra path yet.
#include "h264pred.cuh"
#include "h264pred_hbd.cuh"
#include "scratch.h"
#include "../../include/avdsp_b200.h"
#include <string.h>

namespace avb {

// out: n x n samples (pitch n); tab 5: 8 x 16 samples (pitch 8)
__global__ void __launch_bounds__(256) pred_hbd_kernel(PredJobH j, uint16_t *__restrict__ out)
{
    const int n = j.tab == 0 ? 4 : j.tab == 3 ? 16 : 8, t = threadIdx.x;
    if (j.tab == 5) { if (t < 128) out[t] = (uint16_t)hbd_sample_8x16(j, t & 7, t >> 3); return; }
    if (t >= n * n) return;
    const int x = t % n, y = t / n;
    int v;
    if (j.tab < 2) {
        if (j.mode == 11) v = 1 << (j.bits - 1);
        else { IntraEdges e; hbd_edges(e, j); v = intra_directional(e, n, j.mode, x, y); }
    } else v = hbd_big_sample(j, n, x, y);
    out[t] = (uint16_t)v;
}

// lossless vertical / horizontal prediction + residual (h264pred_template.c:1123-1354).  rect: staged 17 x 17 neighbourhood (row 0 = the
// row above, column 0 = the column to the left), pitch 32 samples.  The C code runs the blocks one after the other (a block below /
// right of another starts from that block's output), so one thread walks them in that order.
__global__ void __launch_bounds__(32) pred_add_hbd_kernel(PredJobH j, uint16_t *__restrict__ rect, int32_t *__restrict__ block)
{
    if (threadIdx.x) return;
    const int kind = j.tab - 4, horizontal = j.mode, n = (kind == 1 || kind == 2) ? 8 : 4;
    const int mask = j.bits > 8 ? 0xffff : 0xff;               // the running sum wraps in the sample type (8-bit jobs only come from the 4:2:2 entries)
    IntraEdges e;
    if (kind == 2) hbd_edges(e, j);
    for (int b = 0; b < j.nblocks; b++) {
        uint16_t *p = rect + 32 + 1 + (kind >= 3 ? j.off[b] : 0);
        int32_t *blk = block + b * n * n;
        for (int i = 0; i < n; i++) {
            int v = kind == 2 ? (horizontal ? e.l[i + 1] : e.t[i + 1]) : (horizontal ? p[-1 + i * 32] : p[i - 32]);
            for (int k = 0; k < n; k++) {
                const int ci = horizontal ? i * n + k : k * n + i;
                v = (v + blk[ci]) & mask;
                p[horizontal ? k + i * 32 : i + k * 32] = (uint16_t)v;
                blk[ci] = 0;
            }
        }
    }
}

namespace {

struct PHStage {
    ScratchLock lk;
    uint8_t *h = nullptr, *d = nullptr;
    cudaStream_t s = nullptr;
    bool ok() {
        Scratch &S = scratch();
        h = (uint8_t *)S.pinned2(64 * 1024); d = (uint8_t *)S.dev(9, 64 * 1024);
        cudaStream_t *st = S.streams();
        if (!h || !d || !st) return false;
        s = st[0];
        return true;
    }
};
inline const uint16_t *px(const uint8_t *p) { return (const uint16_t *)p; }

template <int BITS> void predict(int tab, int mode, uint8_t *src, const uint8_t *topright, int has_tl, int has_tr, ptrdiff_t st)
{
    PHStage S; if (!S.ok()) return;
    PredJobH j; memset(&j, 0, sizeof(j));
    j.bits = BITS; j.tab = tab; j.mode = mode; j.has_tl = has_tl != 0; j.has_tr = has_tr != 0;
    const int n = tab == 0 ? 4 : tab == 3 ? 16 : 8;
    bool top = false, left = false, corner = false, tr = false;
    int nleft = n;
    if (tab < 2) {
        const int nd = intra_needs(mode);
        top = nd & 1; left = nd & 2; corner = nd & 4; tr = nd & 8;
        if (tab == 1) {                                       // the edge filters look one sample further only when told they may
            if ((top || left) && has_tl) corner = true;
            tr = top && has_tr;
        }
    } else if (tab == 2) {
        top = mode == 0 || mode == 2 || mode == 3 || mode == 5 || mode == 7 || mode == 8;
        left = mode == 0 || mode == 1 || mode == 3 || mode == 4 || mode >= 7;
        corner = mode == 3;
        if (mode == 7) nleft = 4;
    } else {
        top = mode == 0 || mode == 2 || mode == 3 || mode == 5;
        left = mode == 0 || mode == 1 || mode == 3 || mode == 4;
        corner = mode == 3;
    }
    if (top) memcpy(j.top, src - st, (size_t)n * 2);
    if (tr) { if (tab == 0) memcpy(j.top + 4, topright, 8); else memcpy(j.top + 8, src - st + 16, (intra_needs(mode) & 8) ? 16 : 2); }
    if (left) for (int i = 0; i < nleft; i++) j.left[i] = *px(src - 2 + i * st);
    if (corner) j.corner = *px(src - 2 - st);
    AVB_LAUNCH(pred_hbd_kernel, 1, 256, 0, S.s)(j, (uint16_t *)S.d);
    if (check_launch("h264_pred slot (high bit depth)")) return;
    if (cudaMemcpyAsync(S.h, S.d, 512, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) {
        set_error("h264_pred slot (high bit depth)", cudaGetLastError()); return;
    }
    for (int y = 0; y < n; y++) memcpy(src + y * st, S.h + (size_t)y * n * 2, (size_t)n * 2);
}

template <int BITS> void predict_add(int kind, int horizontal, uint8_t *pix, const int *block_offset, int16_t *block, int has_tl, int has_tr, ptrdiff_t st)
{
    PHStage S; if (!S.ok()) return;
    PredJobH j; memset(&j, 0, sizeof(j));
    j.bits = BITS; j.tab = 4 + kind; j.mode = horizontal; j.has_tl = has_tl != 0; j.has_tr = has_tr != 0;
    const int n = (kind == 1 || kind == 2) ? 8 : 4, nb = kind == 3 ? 4 : kind == 4 ? 16 : 1, span = kind == 4 ? 16 : 8;
    j.nblocks = nb;
    for (int b = 0; b < nb; b++) {                         // byte offsets in the caller's pitch -> sample offsets in the staged pitch (32)
        const int o = kind >= 3 ? block_offset[b] : 0, oy = (int)(o / st), ox = (int)(o - oy * st) / 2;
        j.off[b] = oy * 32 + ox;
    }
    uint16_t *rect = (uint16_t *)S.h;                      // 17 rows x 32 samples: row 0 / column 0 = the neighbours
    memset(rect, 0, 17 * 32 * 2);
    const int ext = kind >= 3 ? span : n;
    if (kind == 2) {                                       // filtered edges: the same gather as pred8x8l vertical / horizontal
        if (!horizontal) { memcpy(j.top, pix - st, 16); if (has_tr) j.top[8] = *px(pix + 16 - st); }
        else for (int i = 0; i < 8; i++) j.left[i] = *px(pix - 2 + i * st);
        if (has_tl) j.corner = *px(pix - 2 - st);
    }
    if (!horizontal) memcpy(rect + 1, pix - st, (size_t)ext * 2);
    else for (int i = 0; i < ext; i++) rect[32 * (i + 1)] = *px(pix - 2 + i * st);
    for (int y = 0; y < ext; y++) memcpy(rect + 32 * (y + 1) + 1, pix + y * st, (size_t)ext * 2);
    const size_t rbytes = 17 * 32 * 2, cbytes = (size_t)nb * n * n * 4;
    memcpy(S.h + 2048, block, cbytes);
    if (cudaMemcpyAsync(S.d, S.h, 2048 + cbytes, cudaMemcpyHostToDevice, S.s) != cudaSuccess) { set_error("h264_pred_add slot (high bit depth)", cudaGetLastError()); return; }
    AVB_LAUNCH(pred_add_hbd_kernel, 1, 32, 0, S.s)(j, (uint16_t *)S.d, (int32_t *)(S.d + 2048));
    if (check_launch("h264_pred_add slot (high bit depth)")) return;
    if (cudaMemcpyAsync(S.h, S.d, 2048 + cbytes, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) {
        set_error("h264_pred_add slot (high bit depth)", cudaGetLastError()); return;
    }
    (void)rbytes;
    for (int y = 0; y < ext; y++) memcpy(pix + y * st, rect + 32 * (y + 1) + 1, (size_t)ext * 2);
    memcpy(block, S.h + 2048, cbytes);
}

// ---- chroma_format_idc 2: the 8 x 16 pred8x8[] entries and the two pred8x8_add[] entries, bit depth 8 / 9 / 10 ----
// (samples are widened to 16 bits in the job / staged rectangle and narrowed on the way back for 8-bit pictures)
template <int BITS> inline int lds(const uint8_t *p, ptrdiff_t byte_off) { return BITS > 8 ? *(const uint16_t *)(p + byte_off) : p[byte_off]; }
template <int BITS> inline void sts(uint8_t *p, ptrdiff_t byte_off, int v) { if (BITS > 8) *(uint16_t *)(p + byte_off) = (uint16_t)v; else p[byte_off] = (uint8_t)v; }

template <int BITS, int MODE> void s_pred8x16(uint8_t *src, ptrdiff_t st)
{
    PHStage S; if (!S.ok()) return;
    constexpr int SB = BITS > 8 ? 2 : 1;
    PredJobH j; memset(&j, 0, sizeof(j));
    j.bits = BITS; j.tab = 5; j.mode = MODE;
    constexpr bool top = MODE == 0 || MODE == 2 || MODE == 3 || MODE == 5 || MODE == 7 || MODE == 8;
    constexpr bool left = MODE == 0 || MODE == 1 || MODE == 3 || MODE == 4 || MODE >= 7;
    constexpr int nleft = MODE == 7 ? 4 : 16;                 // L0T only looks at the first block's left column
    if (top) for (int i = 0; i < 8; i++) j.top[i] = (uint16_t)lds<BITS>(src, -st + i * SB);
    if (left) for (int i = 0; i < nleft; i++) j.left[i] = (uint16_t)lds<BITS>(src, i * st - SB);
    if (MODE == 3) j.corner = (uint16_t)lds<BITS>(src, -st - SB);
    AVB_LAUNCH(pred_hbd_kernel, 1, 256, 0, S.s)(j, (uint16_t *)S.d);
    if (check_launch("h264_pred slot (4:2:2)")) return;
    if (cudaMemcpyAsync(S.h, S.d, 256, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) {
        set_error("h264_pred slot (4:2:2)", cudaGetLastError()); return;
    }
    const uint16_t *o = (const uint16_t *)S.h;
    for (int y = 0; y < 16; y++) for (int x = 0; x < 8; x++) sts<BITS>(src, y * st + x * SB, o[8 * y + x]);
}
template <int BITS, int HOR> void s_add8x16(uint8_t *pix, const int *block_offset, int16_t *block, ptrdiff_t st)
{
    PHStage S; if (!S.ok()) return;
    constexpr int SB = BITS > 8 ? 2 : 1;
    PredJobH j; memset(&j, 0, sizeof(j));
    j.bits = BITS; j.tab = 4 + 3; j.mode = HOR; j.nblocks = 8;          // the kernel's kind 3: 4 x 4 blocks at j.off[]
    int maxx = 0, maxy = 0;
    for (int b = 0; b < 8; b++) {
        const int o = block_offset[b < 4 ? b : b + 4], oy = (int)(o / st), ox = (int)(o - oy * st) / SB;
        if (o < 0 || ox + 4 > 31 || oy + 4 > 16) { set_error_msg("h264_pred8x16_add slot", "block offsets outside the 8 x 16 chroma block are not taken over"); return; }
        j.off[b] = oy * 32 + ox;
        if (ox + 4 > maxx) maxx = ox + 4;
        if (oy + 4 > maxy) maxy = oy + 4;
    }
    uint16_t *rect = (uint16_t *)S.h;                                   // 17 rows x 32 samples: row 0 / column 0 = the neighbours
    memset(rect, 0, 17 * 32 * 2);
    if (!HOR) for (int x = 0; x < maxx; x++) rect[1 + x] = (uint16_t)lds<BITS>(pix, -st + x * SB);
    else for (int y = 0; y < maxy; y++) rect[32 * (y + 1)] = (uint16_t)lds<BITS>(pix, y * st - SB);
    for (int y = 0; y < maxy; y++) for (int x = 0; x < maxx; x++) rect[32 * (y + 1) + 1 + x] = (uint16_t)lds<BITS>(pix, y * st + x * SB);
    int32_t *cb = (int32_t *)(S.h + 2048);
    for (int i = 0; i < 128; i++) cb[i] = BITS > 8 ? ((const int32_t *)block)[i] : block[i];
    if (cudaMemcpyAsync(S.d, S.h, 2048 + 512, cudaMemcpyHostToDevice, S.s) != cudaSuccess) { set_error("h264_pred8x16_add slot", cudaGetLastError()); return; }
    AVB_LAUNCH(pred_add_hbd_kernel, 1, 32, 0, S.s)(j, (uint16_t *)S.d, (int32_t *)(S.d + 2048));
    if (check_launch("h264_pred8x16_add slot")) return;
    if (cudaMemcpyAsync(S.h, S.d, 2048 + 512, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) {
        set_error("h264_pred8x16_add slot", cudaGetLastError()); return;
    }
    for (int y = 0; y < maxy; y++) for (int x = 0; x < maxx; x++) sts<BITS>(pix, y * st + x * SB, rect[32 * (y + 1) + 1 + x]);
    for (int i = 0; i < 128; i++) { if (BITS > 8) ((int32_t *)block)[i] = cb[i]; else block[i] = (int16_t)cb[i]; }
}
template <int B, int M> struct Fill422 { static void go(H264PredContext *h) { h->pred8x8[M] = s_pred8x16<B, M>; Fill422<B, M - 1>::go(h); } };
template <int B> struct Fill422<B, -1> { static void go(H264PredContext *) {} };
template <int B> void install422(H264PredContext *h)
{
    Fill422<B, 10>::go(h);
    h->pred8x8_add[2] = s_add8x16<B, 0>; h->pred8x8_add[1] = s_add8x16<B, 1>;        // [VERT_PRED8x8 = 2], [HOR_PRED8x8 = 1]
}

template <int B, int MODE> void s_pred4x4(uint8_t *src, const uint8_t *topright, ptrdiff_t stride) { predict<B>(0, MODE, src, topright, 0, 0, stride); }
template <int B, int MODE> void s_pred8x8l(uint8_t *src, int tl, int tr, ptrdiff_t stride) { predict<B>(1, MODE, src, nullptr, tl, tr, stride); }
template <int B, int MODE> void s_pred8x8(uint8_t *src, ptrdiff_t stride) { predict<B>(2, MODE, src, nullptr, 0, 0, stride); }
template <int B, int MODE> void s_pred16x16(uint8_t *src, ptrdiff_t stride) { predict<B>(3, MODE, src, nullptr, 0, 0, stride); }
template <int B, int HOR> void s_add4(uint8_t *pix, int16_t *block, ptrdiff_t stride) { predict_add<B>(0, HOR, pix, nullptr, block, 0, 0, stride); }
template <int B, int HOR> void s_add8l(uint8_t *pix, int16_t *block, ptrdiff_t stride) { predict_add<B>(1, HOR, pix, nullptr, block, 0, 0, stride); }
template <int B, int HOR> void s_add8lf(uint8_t *pix, int16_t *block, int tl, int tr, ptrdiff_t stride) { predict_add<B>(2, HOR, pix, nullptr, block, tl, tr, stride); }
template <int B, int HOR> void s_add8(uint8_t *pix, const int *bo, int16_t *block, ptrdiff_t stride) { predict_add<B>(3, HOR, pix, bo, block, 0, 0, stride); }
template <int B, int HOR> void s_add16(uint8_t *pix, const int *bo, int16_t *block, ptrdiff_t stride) { predict_add<B>(4, HOR, pix, bo, block, 0, 0, stride); }

template <int B, int M> struct Fill12 {
    static void go(H264PredContext *h) { h->pred4x4[M] = s_pred4x4<B, M>; h->pred8x8l[M] = s_pred8x8l<B, M>; Fill12<B, M - 1>::go(h); }
};
template <int B> struct Fill12<B, -1> { static void go(H264PredContext *) {} };
template <int B, int M> struct Fill11 { static void go(H264PredContext *h) { h->pred8x8[M] = s_pred8x8<B, M>; Fill11<B, M - 1>::go(h); } };
template <int B> struct Fill11<B, -1> { static void go(H264PredContext *) {} };
template <int B, int M> struct Fill7 { static void go(H264PredContext *h) { h->pred16x16[M] = s_pred16x16<B, M>; Fill7<B, M - 1>::go(h); } };
template <int B> struct Fill7<B, -1> { static void go(H264PredContext *) {} };

template <int B> void install(H264PredContext *h)
{
    Fill12<B, 11>::go(h); Fill11<B, 10>::go(h); Fill7<B, 6>::go(h);
    h->pred4x4_add[0] = s_add4<B, 0>;   h->pred4x4_add[1] = s_add4<B, 1>;
    h->pred8x8l_add[0] = s_add8l<B, 0>; h->pred8x8l_add[1] = s_add8l<B, 1>;
    h->pred8x8l_filter_add[0] = s_add8lf<B, 0>; h->pred8x8l_filter_add[1] = s_add8lf<B, 1>;
    h->pred8x8_add[2] = s_add8<B, 0>;   h->pred8x8_add[1] = s_add8<B, 1>;          // [VERT_PRED8x8 = 2], [HOR_PRED8x8 = 1]
    h->pred16x16_add[2] = s_add16<B, 0>; h->pred16x16_add[1] = s_add16<B, 1>;
}

}  // namespace

// ff_h264_pred_init_cuda (h264pred.cu) for bit_depth 9 and 10
void h264pred_init_hbd(H264PredContext *h, int bits) { if (bits == 9) install<9>(h); else install<10>(h); }
// chroma_format_idc > 1 (h264pred.c:477-563): the chroma entries become the 8 x 16 functions; bits 8 / 9 / 10
void h264pred_install_422(H264PredContext *h, int bits) { if (bits == 8) install422<8>(h); else if (bits == 9) install422<9>(h); else install422<10>(h); }

}  // namespace avb
