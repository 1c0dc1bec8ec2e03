// libav_b200/csrc/h264pred.cu -- H264PredContext (libavcodec/h264pred.h:91-110) for codec H.264, 8 bit, 4:2:0:
// per-call table slots + ff_h264_pred_init_cuda.  A slot gathers exactly the neighbour samples the C function of that
// mode reads (host pointers), runs the shared device arithmetic (h264pred.cuh) as a batch of one, and scatters the
// predicted block back.  The batched intra reconstruction path lives in h264intra.cu.
#include "h264pred.cuh"
#include "scratch.h"
#include "../../include/avdsp_b200.h"
#include <string.h>

namespace avb {

struct PredJob {
    int tab, mode, has_tl, has_tr;     // tab 0 pred4x4, 1 pred8x8l, 2 pred8x8, 3 pred16x16; 4.. lossless add kinds
    IntraRaw raw;
    int nblocks;                        // lossless: blocks in this call
    int off[16];                        // lossless 8x8 / 16x16: block offsets inside the staged 16 x 16 rectangle (pitch 16)
};

// out: n x n samples (pitch n)
__global__ void __launch_bounds__(256) pred_slot_kernel(PredJob j, uint8_t *__restrict__ out)
{
    __shared__ IntraEdges e;
    __shared__ IntraBig big;
    const int n = j.tab == 0 ? 4 : j.tab == 3 ? 16 : 8;
    if (threadIdx.x == 0) {
        if (j.tab == 0) intra_edges4(e, j.raw);
        else if (j.tab == 1) intra_edges8(e, j.raw, j.has_tl, j.has_tr);
        else intra_big_prepare(big, j.raw, n);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= n * n) return;
    const int x = t % n, y = t / n;
    out[t] = (uint8_t)(j.tab < 2 ? intra_directional(e, n, j.mode, x, y) : intra_big_sample(big, j.raw, n, j.mode, x, y));
}

// lossless vertical / horizontal prediction + residual (h264pred_template.c:1123-1354): thread = one line of one block.
// rect: staged 17 x 17 neighbourhood (row 0 = the row above, column 0 = the column to the left), pitch 32.
__global__ void __launch_bounds__(128) pred_add_slot_kernel(PredJob j, uint8_t *__restrict__ rect, int16_t *__restrict__ block)
{
    __shared__ IntraEdges e;
    const int kind = j.tab - 4, horizontal = j.mode;            // kind 0 4x4, 1 8x8l, 2 8x8l filtered, 3 8x8 (4 blocks), 4 16x16
    const int n = (kind == 1 || kind == 2) ? 8 : 4;
    if (kind == 2 && threadIdx.x == 0) intra_edges8(e, j.raw, j.has_tl, j.has_tr);
    __syncthreads();
    // the C code runs the blocks one after the other: a block below (right of) another starts from that block's output
    for (int b = 0; b < j.nblocks; b++) {
        const int i = threadIdx.x;
        if (i < n) {
            uint8_t *p = rect + 32 + 1 + (kind >= 3 ? (j.off[b] >> 4) * 32 + (j.off[b] & 15) : 0);
            int16_t *blk = block + b * n * n;
            int v = kind == 2 ? (horizontal ? e.l[i + 1] : e.t[i + 1]) : (horizontal ? p[-1 + i * 32] : p[i - 32]);
            for (int k = 0; k < n; k++) {
                const int ci = horizontal ? i * n + k : k * n + i;
                v = (v + blk[ci]) & 255;
                p[horizontal ? k + i * 32 : i + k * 32] = (uint8_t)v;
                blk[ci] = 0;
            }
        }
        __syncthreads();
    }
}

namespace {

struct PStage {
    ScratchLock lk;
    uint8_t *h = nullptr, *d = nullptr;
    cudaStream_t s = nullptr;
    bool ok() {
        Scratch &S = scratch();
        h = (uint8_t *)S.pinned2(64 * 1024); d = (uint8_t *)S.dev(4, 64 * 1024);
        cudaStream_t *st = S.streams();
        if (!h || !d || !st) return false;
        s = st[0];
        return true;
    }
};

// gather what mode `mode` of table `tab` reads, predict on the device, write the block back
void predict(int tab, int mode, uint8_t *src, const uint8_t *topright, int has_tl, int has_tr, ptrdiff_t st)
{
    PStage S; if (!S.ok()) return;
    PredJob j; memset(&j, 0, sizeof(j));
    j.tab = tab; j.mode = mode; j.has_tl = has_tl != 0; j.has_tr = has_tr != 0;
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
    if (top) memcpy(j.raw.top, src - st, n);
    if (tr) { if (tab == 0) memcpy(j.raw.top + 4, topright, 4); else memcpy(j.raw.top + 8, src - st + 8, (intra_needs(mode) & 8) ? 8 : 1); }
    if (left) for (int i = 0; i < nleft; i++) j.raw.left[i] = src[-1 + i * st];
    if (corner) j.raw.corner = src[-1 - st];
    pred_slot_kernel<<<1, 256, 0, S.s>>>(j, S.d);
    if (check_launch("h264_pred slot")) return;
    if (cudaMemcpyAsync(S.h, S.d, 256, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) {
        set_error("h264_pred slot", cudaGetLastError()); return;
    }
    for (int y = 0; y < n; y++) memcpy(src + y * st, S.h + y * n, n);
}

void predict_add(int kind, int horizontal, uint8_t *pix, const int *block_offset, int16_t *block, int has_tl, int has_tr, ptrdiff_t st)
{
    PStage S; if (!S.ok()) return;
    PredJob j; memset(&j, 0, sizeof(j));
    j.tab = 4 + kind; j.mode = horizontal; j.has_tl = has_tl != 0; j.has_tr = has_tr != 0;
    const int n = (kind == 1 || kind == 2) ? 8 : 4, nb = kind == 3 ? 4 : kind == 4 ? 16 : 1, span = kind == 4 ? 16 : 8;
    j.nblocks = nb;
    // block offsets are relative to `pix` in the caller's pitch; restate them in the staged pitch
    for (int b = 0; b < nb; b++) {
        const int o = kind >= 3 ? block_offset[b] : 0, oy = (int)(o / st), ox = (int)(o - oy * st);
        j.off[b] = oy * 16 + ox;
    }
    uint8_t *rect = S.h;                                   // 17 rows x 32: row 0 / column 0 = the neighbours
    memset(rect, 0, 17 * 32);
    const int ext = kind >= 3 ? span : n;
    if (kind == 2) {
        // filtered edges: the same gather as pred8x8l vertical / horizontal
        const int mode = horizontal ? 1 : 0;
        if (!horizontal) { memcpy(j.raw.top, pix - st, 8); if (has_tr) j.raw.top[8] = pix[8 - st]; }
        else for (int i = 0; i < 8; i++) j.raw.left[i] = pix[-1 + i * st];
        if (has_tl) j.raw.corner = pix[-1 - st];
        (void)mode;
    }
    if (!horizontal) memcpy(rect + 1, pix - st, ext);
    else for (int i = 0; i < ext; i++) rect[32 * (i + 1)] = pix[-1 + i * st];
    for (int y = 0; y < ext; y++) memcpy(rect + 32 * (y + 1) + 1, pix + y * st, ext);
    const size_t cbytes = (size_t)nb * n * n * 2;
    memcpy(S.h + 1024, block, cbytes);
    if (cudaMemcpyAsync(S.d, S.h, 1024 + cbytes, cudaMemcpyHostToDevice, S.s) != cudaSuccess) { set_error("h264_pred_add slot", cudaGetLastError()); return; }
    pred_add_slot_kernel<<<1, 128, 0, S.s>>>(j, S.d, (int16_t *)(S.d + 1024));
    if (check_launch("h264_pred_add slot")) return;
    if (cudaMemcpyAsync(S.h, S.d, 1024 + cbytes, cudaMemcpyDeviceToHost, S.s) != cudaSuccess || cudaStreamSynchronize(S.s) != cudaSuccess) {
        set_error("h264_pred_add slot", cudaGetLastError()); return;
    }
    for (int y = 0; y < ext; y++) memcpy(pix + y * st, rect + 32 * (y + 1) + 1, ext);
    memcpy(block, S.h + 1024, cbytes);
}

template <int MODE> void s_pred4x4(uint8_t *src, const uint8_t *topright, ptrdiff_t stride) { predict(0, MODE, src, topright, 0, 0, stride); }
template <int MODE> void s_pred8x8l(uint8_t *src, int tl, int tr, ptrdiff_t stride) { predict(1, MODE, src, nullptr, tl, tr, stride); }
template <int MODE> void s_pred8x8(uint8_t *src, ptrdiff_t stride) { predict(2, MODE, src, nullptr, 0, 0, stride); }
template <int MODE> void s_pred16x16(uint8_t *src, ptrdiff_t stride) { predict(3, MODE, src, nullptr, 0, 0, stride); }
template <int HOR> void s_add4(uint8_t *pix, int16_t *block, ptrdiff_t stride) { predict_add(0, HOR, pix, nullptr, block, 0, 0, stride); }
template <int HOR> void s_add8l(uint8_t *pix, int16_t *block, ptrdiff_t stride) { predict_add(1, HOR, pix, nullptr, block, 0, 0, stride); }
template <int HOR> void s_add8lf(uint8_t *pix, int16_t *block, int tl, int tr, ptrdiff_t stride) { predict_add(2, HOR, pix, nullptr, block, tl, tr, stride); }
template <int HOR> void s_add8(uint8_t *pix, const int *bo, int16_t *block, ptrdiff_t stride) { predict_add(3, HOR, pix, bo, block, 0, 0, stride); }
template <int HOR> void s_add16(uint8_t *pix, const int *bo, int16_t *block, ptrdiff_t stride) { predict_add(4, HOR, pix, bo, block, 0, 0, stride); }

template <int M> struct Fill12 {
    static void go(H264PredContext *h) { h->pred4x4[M] = s_pred4x4<M>; h->pred8x8l[M] = s_pred8x8l<M>; Fill12<M - 1>::go(h); }
};
template <> struct Fill12<-1> { static void go(H264PredContext *) {} };
template <int M> struct Fill11 { static void go(H264PredContext *h) { h->pred8x8[M] = s_pred8x8<M>; Fill11<M - 1>::go(h); } };
template <> struct Fill11<-1> { static void go(H264PredContext *) {} };
template <int M> struct Fill7 { static void go(H264PredContext *h) { h->pred16x16[M] = s_pred16x16<M>; Fill7<M - 1>::go(h); } };
template <> struct Fill7<-1> { static void go(H264PredContext *) {} };

}  // namespace
}  // namespace avb

namespace avb { void h264pred_init_hbd(H264PredContext *h, int bits); void h264pred_install_422(H264PredContext *h, int bits); }

using namespace avb;

// libavcodec/h264pred.h:112-113 / the per-arch hooks :114-123.  Takes over codec H.264 (AV_CODEC_ID_H264 = 27), 8 bit,
// chroma_format_idc <= 1 -- and the 9 / 10-bit instances through h264pred_hbd.cu; anything else leaves the table as the C init filled it.
extern "C" void ff_h264_pred_init_cuda(H264PredContext *h, int codec_id, const int bit_depth, const int chroma_format_idc)
{
    avb::enter();
    if (!h || codec_id != 27 || (bit_depth != 8 && bit_depth != 9 && bit_depth != 10)) return;
    if (bit_depth != 8) {                                     // h264pred_hbd.cu
        h264pred_init_hbd(h, bit_depth);
        if (chroma_format_idc > 1) h264pred_install_422(h, bit_depth);
        return;
    }
    Fill12<11>::go(h); Fill11<10>::go(h); Fill7<6>::go(h);
    h->pred4x4_add[0] = s_add4<0>;   h->pred4x4_add[1] = s_add4<1>;
    h->pred8x8l_add[0] = s_add8l<0>; h->pred8x8l_add[1] = s_add8l<1>;
    h->pred8x8l_filter_add[0] = s_add8lf<0>; h->pred8x8l_filter_add[1] = s_add8lf<1>;
    h->pred8x8_add[2] = s_add8<0>;   h->pred8x8_add[1] = s_add8<1>;        // [VERT_PRED8x8 = 2], [HOR_PRED8x8 = 1]
    h->pred16x16_add[2] = s_add16<0>; h->pred16x16_add[1] = s_add16<1>;
    if (chroma_format_idc > 1) h264pred_install_422(h, 8);    // h264pred.c:477-563: the pred8x8[] / pred8x8_add[] entries become the 8 x 16 functions
}
