// libav_b200/csrc/h264pred_hbd.cuh -- the neighbour set of one intra block at 9 / 10 bit (also the 8-bit 4:2:2 jobs) and the predictors that work
// from it: the edge arrays of the 4x4 / 8x8 luma predictors (h264pred_template.c:840-875 filtering for pred8x8l) and the DC / plane family of
// pred8x8 / pred16x16 (:330-797).  Shared by the table slots (h264pred_hbd.cu) and the batched intra reconstruction (h264_hbd_batch.cu).
// Thread-local arithmetic only.
#pragma once
#include "h264pred.cuh"

namespace avb {

struct PredJobH {
    int bits, tab, mode, has_tl, has_tr;   // tab 0 pred4x4, 1 pred8x8l, 2 pred8x8, 3 pred16x16; 4.. lossless add kinds
    uint16_t top[16], left[16], corner, pad;
    int nblocks;
    int off[16];                           // lossless 8x8 / 16x16: block offsets (samples) inside the staged rectangle (pitch 32)
};

__device__ inline void hbd_edges(IntraEdges &e, const PredJobH &j)
{
    const uint16_t *t = j.top, *l = j.left;
    const int c = j.corner;
    if (j.tab == 0 || j.tab == 4) {                        // raw edges of a 4x4 block
        e.t[0] = e.l[0] = c;
        for (int i = 0; i < 8; i++) e.t[i + 1] = t[i];
        for (int i = 0; i < 4; i++) e.l[i + 1] = l[i];
        return;
    }
    // 8x8 luma: low-pass filtered with the availability rules of PREDICT_8x8_LOAD_* (h264pred_template.c:840-875)
    e.t[1] = ip_f3(j.has_tl ? c : t[0], t[0], t[1]);
    for (int i = 1; i < 7; i++) e.t[i + 1] = ip_f3(t[i - 1], t[i], t[i + 1]);
    e.t[8] = ip_f3(j.has_tr ? t[8] : t[7], t[7], t[6]);
    if (j.has_tr) {
        for (int i = 8; i < 15; i++) e.t[i + 1] = ip_f3(t[i - 1], t[i], t[i + 1]);
        e.t[16] = (t[14] + 3 * t[15] + 2) >> 2;
    } else {
        for (int i = 8; i < 16; i++) e.t[i + 1] = t[7];
    }
    e.l[1] = ip_f3(j.has_tl ? c : l[0], l[0], l[1]);
    for (int i = 1; i < 7; i++) e.l[i + 1] = ip_f3(l[i - 1], l[i], l[i + 1]);
    e.l[8] = (l[6] + 3 * l[7] + 2) >> 2;
    e.t[0] = e.l[0] = ip_f3(l[0], c, t[0]);
}

// pred8x8 (chroma, n = 8, modes 0..10) and pred16x16 (n = 16, modes 0..6) from raw edges (h264pred_template.c:330-797)
__device__ inline int hbd_big_sample(const PredJobH &j, int n, int x, int y)
{
    const int mode = j.mode, mid = 1 << (j.bits - 1);
    if (mode == 1) return j.left[y];
    if (mode == 2) return j.top[x];
    if (mode == 6) return mid;
    if (mode == 3) {
        const int h = n / 2;
        int H = 0, V = 0;
        for (int k = 1; k <= h; k++) {
            const int tl = h - 1 - k < 0 ? j.corner : j.top[h - 1 - k], ll = h - 1 - k < 0 ? j.corner : j.left[h - 1 - k];
            H += k * (j.top[h - 1 + k] - tl); V += k * (j.left[h - 1 + k] - ll);
        }
        if (n == 8) { H = (17 * H + 16) >> 5; V = (17 * V + 16) >> 5; } else { H = (5 * H + 32) >> 6; V = (5 * V + 32) >> 6; }
        const int a = 16 * (j.left[n - 1] + j.top[n - 1] + 1) - (h - 1) * (V + H);
        return min(max((a + x * H + y * V) >> 5, 0), (1 << j.bits) - 1);
    }
    int st[4] = { 0, 0, 0, 0 }, sl[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < n; i++) { st[i >> 2] += j.top[i]; sl[i >> 2] += j.left[i]; }
    if (n == 16) {
        const int T = st[0] + st[1] + st[2] + st[3], L = sl[0] + sl[1] + sl[2] + sl[3];
        return mode == 0 ? (T + L + 16) >> 5 : mode == 4 ? (L + 8) >> 4 : (T + 8) >> 4;
    }
    const int q = (x >> 2) + 2 * (y >> 2), t0 = st[0], t1 = st[1], l0 = sl[0], l1 = sl[1];
    const int dc_q[4] = { (t0 + l0 + 4) >> 3, (t1 + 2) >> 2, (l1 + 2) >> 2, (t1 + l1 + 4) >> 3 };
    const int top_q = (((q & 1) ? t1 : t0) + 2) >> 2, left_q = (((q >> 1) ? l1 : l0) + 2) >> 2;
    switch (mode) {
    case 0: return dc_q[q];
    case 4: return left_q;
    case 5: return top_q;
    case 7: return q == 0 ? dc_q[0] : top_q;
    case 8: return q == 0 ? top_q : dc_q[q];
    case 9: return q < 2 ? left_q : mid;
    default: return q < 2 ? mid : left_q;
    }
}

// chroma_format_idc 2: sample (x, y) of an 8 x 16 chroma block, modes 0..10 of pred8x8[] (h264pred_template.c:502-838): the DC family per
// 4 x 4 quadrant (2 across, 4 down), plane prediction with an 8-tap vertical gradient
__device__ inline int hbd_sample_8x16(const PredJobH &j, int x, int y)
{
    const int mode = j.mode, mid = 1 << (j.bits - 1);
    if (mode == 1) return j.left[y];
    if (mode == 2) return j.top[x];
    if (mode == 6) return mid;
    if (mode == 3) {
        int H = 0, V = 0;
        for (int k = 1; k <= 4; k++) H += k * (j.top[3 + k] - (3 - k < 0 ? j.corner : j.top[3 - k]));
        for (int k = 1; k <= 8; k++) V += k * (j.left[7 + k] - (7 - k < 0 ? j.corner : j.left[7 - k]));
        H = (17 * H + 16) >> 5; V = (5 * V + 32) >> 6;
        const int a = 16 * (j.left[15] + j.top[7] + 1) - 7 * V - 3 * H;
        return min(max((a + x * H + y * V) >> 5, 0), (1 << j.bits) - 1);
    }
    int t0 = 0, t1 = 0, l[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < 4; i++) { t0 += j.top[i]; t1 += j.top[4 + i]; }
    for (int i = 0; i < 16; i++) l[i >> 2] += j.left[i];
    const int c = x >> 2, r = y >> 2;
    const int dc = !c ? (!r ? (t0 + l[0] + 4) >> 3 : (l[r] + 2) >> 2) : (!r ? (t1 + 2) >> 2 : (t1 + l[r] + 4) >> 3);
    const int ldc = (l[r] + 2) >> 2, tdc = ((c ? t1 : t0) + 2) >> 2;
    switch (mode) {
    case 0: return dc;
    case 4: return ldc;
    case 5: return tdc;
    case 7: return (!c && !r) ? (t0 + l[0] + 4) >> 3 : tdc;          // top_dc, then pred4x4_dc on the first block
    case 8: return (!c && !r) ? (t0 + 2) >> 2 : dc;                  // dc, then pred4x4_top_dc on the first block
    case 9: return r == 1 ? mid : ldc;                               // left_dc, then 128 on the two blocks of the second row
    default: return r == 0 ? mid : ldc;                              // left_dc, then 128 on the two blocks of the first row
    }
}

}  // namespace avb
