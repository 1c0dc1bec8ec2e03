/*
 * oracle/port/orc_h264pred.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the H.264 intra predictors (H264PredContext for codec H.264, 8 bit, 4:2:0;
 * libavcodec/h264pred_template.c, table layout libavcodec/h264pred.h:91-110, h264pred.c:411-585).
 * Instead of the reference's one unrolled function per mode and size, every directional mode is ONE formula over two
 * edge arrays T[-1 .. 2N-1] (row above, T[-1] = corner) and L[-1 .. N-1] (column to the left, L[-1] = corner), for
 * N = 4 (raw edges, h264pred_template.c:34-328) and N = 8 (edges low-pass filtered with the availability rules of
 * PREDICT_8x8_LOAD_*, :840-875).  Plane prediction (:415-482, :760-797) and the chroma DC family (:560-757) follow
 * the standard's formulas.  Pinned byte-for-byte against oracle/_ref in tests/test_oracle_h264pred_cpu.py.
 */
#include <stdint.h>
#include <string.h>
#include "../oracle_api.h"

static inline int clip_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

typedef struct Edges { int t[17], l[9]; } Edges;            /* t[1 + i] = T[i], l[1 + i] = L[i]; index 0 = corner */
#define T(i) (e->t[(i) + 1])
#define L(i) (e->l[(i) + 1])
static inline int f3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }
static inline int f2(int a, int b) { return (a + b + 1) >> 1; }

/* value of sample (x, y) of an N x N block, modes 0..11 of pred4x4 / pred8x8l */
static int directional(const Edges *e, int n, int mode, int x, int y)
{
    int s = 0, z, i;
    switch (mode) {
    case 0: return T(x);
    case 1: return L(y);
    case 2: for (i = 0; i < n; i++) s += T(i) + L(i); return (s + n) >> (n == 4 ? 3 : 4);
    case 9: for (i = 0; i < n; i++) s += L(i); return (s + n / 2) >> (n == 4 ? 2 : 3);
    case 10: for (i = 0; i < n; i++) s += T(i); return (s + n / 2) >> (n == 4 ? 2 : 3);
    case 11: return 128;
    case 3: return (x == n - 1 && y == n - 1) ? (T(2 * n - 2) + 3 * T(2 * n - 1) + 2) >> 2 : f3(T(x + y), T(x + y + 1), T(x + y + 2));
    case 4: return x > y ? f3(T(x - y - 2), T(x - y - 1), T(x - y)) : x < y ? f3(L(y - x - 2), L(y - x - 1), L(y - x)) : f3(T(0), T(-1), L(0));
    case 5:
        z = 2 * x - y; i = x - (y >> 1);
        if (z >= 0) return (z & 1) ? f3(T(i - 2), T(i - 1), T(i)) : f2(T(i - 1), T(i));
        return z == -1 ? f3(L(0), T(-1), T(0)) : f3(L(y - 2 * x - 1), L(y - 2 * x - 2), L(y - 2 * x - 3));
    case 6:
        z = 2 * y - x; i = y - (x >> 1);
        if (z >= 0) return (z & 1) ? f3(L(i - 2), L(i - 1), L(i)) : f2(L(i - 1), L(i));
        return z == -1 ? f3(L(0), T(-1), T(0)) : f3(T(x - 2 * y - 1), T(x - 2 * y - 2), T(x - 2 * y - 3));
    case 7:
        i = x + (y >> 1);
        return (y & 1) ? f3(T(i), T(i + 1), T(i + 2)) : f2(T(i), T(i + 1));
    default:
        z = x + 2 * y; i = y + (x >> 1);
        if (z > 2 * n - 3) return L(n - 1);
        if (z == 2 * n - 3) return (L(n - 2) + 3 * L(n - 1) + 2) >> 2;
        return (z & 1) ? f3(L(i), L(i + 1), L(i + 2)) : f2(L(i), L(i + 1));
    }
}

/* which neighbours a mode reads: bit 0 top, 1 left, 2 corner, 3 top-right */
static int needs(int mode)
{
    static const uint8_t tab[12] = { 1, 2, 3, 1 | 8, 7, 7, 7, 1 | 8, 2, 2, 1, 0 };
    return tab[mode];
}

static void load_edges4(Edges *e, const uint8_t *src, const uint8_t *topright, ptrdiff_t st, int mode)
{
    const int nd = needs(mode);
    memset(e, 0, sizeof(*e));
    if (nd & 1) for (int i = 0; i < 4; i++) T(i) = src[i - st];
    if (nd & 2) for (int i = 0; i < 4; i++) L(i) = src[-1 + i * st];
    if (nd & 4) T(-1) = L(-1) = src[-1 - st];
    if (nd & 8) for (int i = 0; i < 4; i++) T(4 + i) = topright[i];
}

/* filtered edges of an 8x8 luma block, h264pred_template.c:840-875 */
static void load_edges8(Edges *e, const uint8_t *src, ptrdiff_t st, int mode, int has_tl, int has_tr)
{
    const int nd = needs(mode);
    memset(e, 0, sizeof(*e));
#define P(x, y) ((int)src[(x) + (y) * st])
    if (nd & 1) {
        T(0) = f3(has_tl ? P(-1, -1) : P(0, -1), P(0, -1), P(1, -1));
        for (int i = 1; i < 7; i++) T(i) = f3(P(i - 1, -1), P(i, -1), P(i + 1, -1));
        T(7) = f3(has_tr ? P(8, -1) : P(7, -1), P(7, -1), P(6, -1));
    }
    if (nd & 8) {
        if (has_tr) {
            for (int i = 8; i < 15; i++) T(i) = f3(P(i - 1, -1), P(i, -1), P(i + 1, -1));
            T(15) = (P(14, -1) + 3 * P(15, -1) + 2) >> 2;
        } else {
            for (int i = 8; i < 16; i++) T(i) = P(7, -1);
        }
    }
    if (nd & 2) {
        L(0) = f3(has_tl ? P(-1, -1) : P(-1, 0), P(-1, 0), P(-1, 1));
        for (int i = 1; i < 7; i++) L(i) = f3(P(-1, i - 1), P(-1, i), P(-1, i + 1));
        L(7) = (P(-1, 6) + 3 * P(-1, 7) + 2) >> 2;
    }
    if (nd & 4) T(-1) = L(-1) = f3(P(-1, 0), P(-1, -1), P(0, -1));
#undef P
}

static void fill(uint8_t *d, ptrdiff_t st, int w, int h, int v)
{
    for (int y = 0; y < h; y++) memset(d + y * st, v, w);
}

/* plane prediction of an N x N block (N = 8 chroma, 16 luma) */
static void plane(uint8_t *src, ptrdiff_t st, int n)
{
    const int h = n / 2;
    int H = 0, V = 0;
    for (int k = 1; k <= h; k++) {
        H += k * (src[h - 1 + k - st] - src[h - 1 - k - st]);
        V += k * (src[-1 + (h - 1 + k) * st] - src[-1 + (h - 1 - k) * st]);
    }
    if (n == 8) { H = (17 * H + 16) >> 5; V = (17 * V + 16) >> 5; }
    else        { H = (5 * H + 32) >> 6;  V = (5 * V + 32) >> 6; }
    const int a = 16 * (src[-1 + (n - 1) * st] + src[n - 1 - st] + 1) - (h - 1) * (V + H);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) src[x + y * st] = clip_u8((a + x * H + y * V) >> 5);
}

void orc_h264_pred(int tab, int mode, uint8_t *src, const uint8_t *topright, int has_topleft, int has_topright, ptrdiff_t st)
{
    Edges e;
    if (tab == 0 || tab == 1) {
        const int n = tab ? 8 : 4;
        uint8_t out[64];
        if (tab) load_edges8(&e, src, st, mode, has_topleft, has_topright);
        else     load_edges4(&e, src, topright, st, mode);
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) out[x + n * y] = (uint8_t)directional(&e, n, mode, x, y);
        for (int y = 0; y < n; y++) memcpy(src + y * st, out + n * y, n);
        return;
    }
    const int n = tab == 2 ? 8 : 16;
    int sum_t[4] = { 0, 0, 0, 0 }, sum_l[4] = { 0, 0, 0, 0 };           /* per group of 4 samples */
    const int chroma_dc = tab == 2 && (mode == 0 || mode == 4 || mode == 5 || mode >= 7);
    const int read_t = tab == 2 ? (mode == 0 || mode == 5 || mode == 7 || mode == 8) : (mode == 0 || mode == 5);
    const int read_l = tab == 2 ? (mode == 0 || mode == 4 || mode >= 7) : (mode == 0 || mode == 4);
    if (read_t) for (int i = 0; i < n; i++) sum_t[i >> 2] += src[i - st];
    const int nl = (tab == 2 && mode == 7) ? 4 : n;          /* L0T reads the upper half of the left column only */
    if (read_l) for (int i = 0; i < nl; i++) sum_l[i >> 2] += src[-1 + i * st];
    if (mode == 1) { for (int y = 0; y < n; y++) memset(src + y * st, src[-1 + y * st], n); return; }
    if (mode == 2) { uint8_t top[16]; memcpy(top, src - st, n); for (int y = 0; y < n; y++) memcpy(src + y * st, top, n); return; }
    if (mode == 3) { plane(src, st, n); return; }
    if (mode == 6) { fill(src, st, n, n, 128); return; }
    if (tab == 3) {
        const int T4 = sum_t[0] + sum_t[1] + sum_t[2] + sum_t[3], L4 = sum_l[0] + sum_l[1] + sum_l[2] + sum_l[3];
        fill(src, st, 16, 16, mode == 0 ? (T4 + L4 + 16) >> 5 : mode == 4 ? (L4 + 8) >> 4 : (T4 + 8) >> 4);
        return;
    }
    if (chroma_dc) {
        /* quadrant DCs, h264pred_template.c:560-757: q0 top-left, q1 top-right, q2 bottom-left, q3 bottom-right */
        int q[4];
        const int t0 = sum_t[0], t1 = sum_t[1], l0 = sum_l[0], l1 = sum_l[1];
        switch (mode) {
        case 0:  q[0] = (t0 + l0 + 4) >> 3; q[1] = (t1 + 2) >> 2; q[2] = (l1 + 2) >> 2; q[3] = (t1 + l1 + 4) >> 3; break;
        case 4:  q[0] = q[1] = (l0 + 2) >> 2; q[2] = q[3] = (l1 + 2) >> 2; break;
        case 5:  q[0] = q[2] = (t0 + 2) >> 2; q[1] = q[3] = (t1 + 2) >> 2; break;
        case 7:  /* L0T: top_dc, then the 4x4 DC of the top-left quadrant */
                 q[0] = (t0 + l0 + 4) >> 3; q[2] = (t0 + 2) >> 2; q[1] = q[3] = (t1 + 2) >> 2; break;
        case 8:  /* 0LT: full dc, then the 4x4 top DC of the top-left quadrant */
                 q[0] = (t0 + 2) >> 2; q[1] = (t1 + 2) >> 2; q[2] = (l1 + 2) >> 2; q[3] = (t1 + l1 + 4) >> 3; break;
        case 9:  /* L00: left_dc, bottom half 128 */
                 q[0] = q[1] = (l0 + 2) >> 2; q[2] = q[3] = 128; break;
        default: /* 0L0: left_dc, top half 128 */
                 q[0] = q[1] = 128; q[2] = q[3] = (l1 + 2) >> 2; break;
        }
        for (int k = 0; k < 4; k++) fill(src + 4 * (k & 1) + 4 * (k >> 1) * st, st, 4, 4, q[k]);
    }
}

/* lossless vertical / horizontal prediction with the running uint8 sum of the C code (h264pred_template.c:1123-1354) */
static void add_block(uint8_t *pix, int16_t *block, ptrdiff_t st, int n, int horizontal, const int *first /* n start values */)
{
    for (int i = 0; i < n; i++) {
        uint8_t v = (uint8_t)first[i];
        for (int k = 0; k < n; k++) {
            v = (uint8_t)(v + block[horizontal ? i * n + k : k * n + i]);
            pix[horizontal ? k + i * st : i + k * st] = v;
        }
    }
    memset(block, 0, sizeof(int16_t) * n * n);
}

void orc_h264_pred_add(int tab, int mode, uint8_t *pix, const int *block_offset, int16_t *block, int has_topleft, int has_topright,
                       ptrdiff_t st)
{
    int first[8];
    if (tab == 0 || tab == 1) {
        const int n = tab ? 8 : 4;
        for (int i = 0; i < n; i++) first[i] = mode ? pix[-1 + i * st] : pix[i - st];
        add_block(pix, block, st, n, mode, first);
    } else if (tab == 2) {
        Edges e;
        load_edges8(&e, pix, st, mode ? 1 : 0, has_topleft, has_topright);
        for (int i = 0; i < 8; i++) first[i] = mode ? e.l[1 + i] : e.t[1 + i];
        add_block(pix, block, st, 8, mode, first);
    } else {
        const int nb = tab == 3 ? 4 : 16;
        for (int b = 0; b < nb; b++) {
            uint8_t *p = pix + block_offset[b];
            for (int i = 0; i < 4; i++) first[i] = mode ? p[-1 + i * st] : p[i - st];
            add_block(p, block + 16 * b, st, 4, mode, first);
        }
    }
}
