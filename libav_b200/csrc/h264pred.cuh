// libav_b200/csrc/h264pred.cuh -- H.264 intra prediction arithmetic (8 bit), shared by the table slots and the batched
// intra reconstruction kernel.  Follows libavcodec/h264pred_template.c in VALUES, not in shape: every directional mode of
// pred4x4 / pred8x8l is one expression over two edge arrays T[-1..2N-1] (row above, T[-1] = corner) and L[-1..N-1]
// (column to the left, L[-1] = corner) -- raw samples for 4x4, low-pass filtered with the availability rules of
// PREDICT_8x8_LOAD_* (:840-875) for 8x8 -- so one thread can produce any sample independently.
#pragma once
#include "common.cuh"

namespace avb {

struct IntraEdges {                    // T(i) = t[i + 1], L(i) = l[i + 1]
    int t[17], l[9];
};

__device__ __forceinline__ int ip_f3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }
__device__ __forceinline__ int ip_f2(int a, int b) { return (a + b + 1) >> 1; }

// bit 0 top, 1 left, 2 corner, 3 top-right: what mode m of pred4x4 / pred8x8l reads
__device__ __host__ __forceinline__ int intra_needs(int mode)
{
    return (int)((0x123977792210ull >> (4 * (11 - mode))) & 15);     // modes 0..11: 1,2,3,1|8,7,7,7,1|8,2,2,1,0
}

// sample (x, y) of an N x N block, modes 0..11 (h264pred.h:32-46)
__device__ __forceinline__ int intra_directional(const IntraEdges &e, int n, int mode, int x, int y)
{
#define IT(i) e.t[(i) + 1]
#define IL(i) e.l[(i) + 1]
    int s = 0, z, i;
    switch (mode) {
    case 0: return IT(x);
    case 1: return IL(y);
    case 2: for (i = 0; i < n; i++) s += IT(i) + IL(i); return (s + n) >> (n == 4 ? 3 : 4);
    case 9: for (i = 0; i < n; i++) s += IL(i); return (s + n / 2) >> (n == 4 ? 2 : 3);
    case 10: for (i = 0; i < n; i++) s += IT(i); return (s + n / 2) >> (n == 4 ? 2 : 3);
    case 11: return 128;
    case 3: return (x == n - 1 && y == n - 1) ? (IT(2 * n - 2) + 3 * IT(2 * n - 1) + 2) >> 2 : ip_f3(IT(x + y), IT(x + y + 1), IT(x + y + 2));
    case 4: return x > y ? ip_f3(IT(x - y - 2), IT(x - y - 1), IT(x - y)) : x < y ? ip_f3(IL(y - x - 2), IL(y - x - 1), IL(y - x)) : ip_f3(IT(0), IT(-1), IL(0));
    case 5:
        z = 2 * x - y; i = x - (y >> 1);
        if (z >= 0) return (z & 1) ? ip_f3(IT(i - 2), IT(i - 1), IT(i)) : ip_f2(IT(i - 1), IT(i));
        return z == -1 ? ip_f3(IL(0), IT(-1), IT(0)) : ip_f3(IL(y - 2 * x - 1), IL(y - 2 * x - 2), IL(y - 2 * x - 3));
    case 6:
        z = 2 * y - x; i = y - (x >> 1);
        if (z >= 0) return (z & 1) ? ip_f3(IL(i - 2), IL(i - 1), IL(i)) : ip_f2(IL(i - 1), IL(i));
        return z == -1 ? ip_f3(IL(0), IT(-1), IT(0)) : ip_f3(IT(x - 2 * y - 1), IT(x - 2 * y - 2), IT(x - 2 * y - 3));
    case 7:
        i = x + (y >> 1);
        return (y & 1) ? ip_f3(IT(i), IT(i + 1), IT(i + 2)) : ip_f2(IT(i), IT(i + 1));
    default:
        z = x + 2 * y; i = y + (x >> 1);
        if (z > 2 * n - 3) return IL(n - 1);
        if (z == 2 * n - 3) return (IL(n - 2) + 3 * IL(n - 1) + 2) >> 2;
        return (z & 1) ? ip_f3(IL(i), IL(i + 1), IL(i + 2)) : ip_f2(IL(i), IL(i + 1));
    }
#undef IT
#undef IL
}

// raw neighbours of a block as the caller gathered them: top[i] = sample (i, -1) for i = 0..2N-1 (4x4: i >= 4 from
// the `topright` pointer), left[i] = sample (-1, i), corner = (-1, -1).  Entries a mode does not read are don't-care.
struct IntraRaw { uint8_t top[16], left[16], corner, pad[3]; };

__device__ __forceinline__ void intra_edges4(IntraEdges &e, const IntraRaw &r)
{
    e.t[0] = e.l[0] = r.corner;
    for (int i = 0; i < 8; i++) e.t[i + 1] = r.top[i];
    for (int i = 0; i < 4; i++) e.l[i + 1] = r.left[i];
}

__device__ __forceinline__ void intra_edges8(IntraEdges &e, const IntraRaw &r, bool has_tl, bool has_tr)
{
    const uint8_t *t = r.top, *l = r.left;
    const int c = r.corner;
    e.t[1] = ip_f3(has_tl ? c : t[0], t[0], t[1]);
    for (int i = 1; i < 7; i++) e.t[i + 1] = ip_f3(t[i - 1], t[i], t[i + 1]);
    e.t[8] = ip_f3(has_tr ? t[8] : t[7], t[7], t[6]);
    if (has_tr) {
        for (int i = 8; i < 15; i++) e.t[i + 1] = ip_f3(t[i - 1], t[i], t[i + 1]);
        e.t[16] = (t[14] + 3 * t[15] + 2) >> 2;
    } else {
        for (int i = 8; i < 16; i++) e.t[i + 1] = t[7];
    }
    e.l[1] = ip_f3(has_tl ? c : l[0], l[0], l[1]);
    for (int i = 1; i < 7; i++) e.l[i + 1] = ip_f3(l[i - 1], l[i], l[i + 1]);
    e.l[8] = (l[6] + 3 * l[7] + 2) >> 2;
    e.t[0] = e.l[0] = ip_f3(l[0], c, t[0]);
}

// pred8x8 (chroma, n = 8, modes 0..10) and pred16x16 (n = 16, modes 0..6): sample (x, y) from raw edges
// (h264pred_template.c:330-797).  `sums` are the per-4-sample sums of the top row (0..3) and the left column (4..7).
struct IntraBig { int st[4], sl[4]; int H, V, a; };

__device__ __forceinline__ void intra_big_prepare(IntraBig &b, const IntraRaw &r, int n)
{
    for (int k = 0; k < 4; k++) { b.st[k] = 0; b.sl[k] = 0; }
    for (int i = 0; i < n; i++) { b.st[i >> 2] += r.top[i]; b.sl[i >> 2] += r.left[i]; }
    const int h = n / 2;
    int H = 0, V = 0;
    for (int k = 1; k <= h; k++) {
        const int tl = h - 1 - k < 0 ? r.corner : r.top[h - 1 - k], ll = h - 1 - k < 0 ? r.corner : r.left[h - 1 - k];
        H += k * (r.top[h - 1 + k] - tl);
        V += k * (r.left[h - 1 + k] - ll);
    }
    if (n == 8) { H = (17 * H + 16) >> 5; V = (17 * V + 16) >> 5; } else { H = (5 * H + 32) >> 6; V = (5 * V + 32) >> 6; }
    b.H = H; b.V = V;
    b.a = 16 * (r.left[n - 1] + r.top[n - 1] + 1) - (h - 1) * (V + H);
}

__device__ __forceinline__ int intra_big_sample(const IntraBig &b, const IntraRaw &r, int n, int mode, int x, int y)
{
    if (mode == 1) return r.left[y];
    if (mode == 2) return r.top[x];
    if (mode == 3) return clip_u8((b.a + x * b.H + y * b.V) >> 5);
    if (mode == 6) return 128;
    if (n == 16) {
        const int T = b.st[0] + b.st[1] + b.st[2] + b.st[3], L = b.sl[0] + b.sl[1] + b.sl[2] + b.sl[3];
        return mode == 0 ? (T + L + 16) >> 5 : mode == 4 ? (L + 8) >> 4 : (T + 8) >> 4;
    }
    const int q = (x >> 2) + 2 * (y >> 2), t0 = b.st[0], t1 = b.st[1], l0 = b.sl[0], l1 = b.sl[1];
    const int dc_q[4] = { (t0 + l0 + 4) >> 3, (t1 + 2) >> 2, (l1 + 2) >> 2, (t1 + l1 + 4) >> 3 };       // pred8x8_dc
    const int top_q = ((q & 1) ? t1 : t0) + 2 >> 2, left_q = ((q >> 1) ? l1 : l0) + 2 >> 2;
    switch (mode) {
    case 0: return dc_q[q];
    case 4: return left_q;
    case 5: return top_q;
    case 7: return q == 0 ? dc_q[0] : top_q;                        // L0T: top_dc, then the 4x4 DC of quadrant 0
    case 8: return q == 0 ? top_q : dc_q[q];                        // 0LT: dc, then the 4x4 top DC of quadrant 0
    case 9: return q < 2 ? left_q : 128;                            // L00
    default: return q < 2 ? 128 : left_q;                           // 0L0
    }
}

}  // namespace avb
