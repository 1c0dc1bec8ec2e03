/*
 * oracle/port/orc_h264pred_hbd.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The 9 / 10-bit instances of the H.264 intra predictors (ff_h264_pred_init(h, AV_CODEC_ID_H264, bits, 1): h264pred_template.c with
 * BIT_DEPTH > 8, i.e. uint16 samples, int32 residual for the lossless *_add functions; table layout h264pred.h:91-110, h264pred.c:411-585).
 * Same formulas as orc_h264pred.c (one expression per directional mode over two edge arrays); what the depth changes: the DC_128 family
 * predicts 1 << (bits - 1), plane prediction clips to `bits` bits, the lossless running sum wraps in 16 bits.  `st` is in BYTES.
 * Pinned against oracle/_ref in tests/test_oracle_h264_hbd_cpu.py.
 */
#include <stdint.h>
#include <string.h>
#include "../oracle_api.h"

typedef uint16_t px;
typedef struct Edges { int t[17], l[9]; } Edges;
#define T(i) (e->t[(i) + 1])
#define L(i) (e->l[(i) + 1])
static inline int f3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }
static inline int f2(int a, int b) { return (a + b + 1) >> 1; }

static int directional(const Edges *e, int n, int mode, int x, int y, int mid)
{
    int s = 0, z, i;
    switch (mode) {
    case 0: return T(x);
    case 1: return L(y);
    case 2: for (i = 0; i < n; i++) s += T(i) + L(i); return (s + n) >> (n == 4 ? 3 : 4);
    case 9: for (i = 0; i < n; i++) s += L(i); return (s + n / 2) >> (n == 4 ? 2 : 3);
    case 10: for (i = 0; i < n; i++) s += T(i); return (s + n / 2) >> (n == 4 ? 2 : 3);
    case 11: return mid;
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
static int needs(int mode)
{
    static const uint8_t tab[12] = { 1, 2, 3, 1 | 8, 7, 7, 7, 1 | 8, 2, 2, 1, 0 };
    return tab[mode];
}
static void load_edges4(Edges *e, const px *src, const px *topright, int st, int mode)
{
    const int nd = needs(mode);
    memset(e, 0, sizeof(*e));
    if (nd & 1) for (int i = 0; i < 4; i++) T(i) = src[i - st];
    if (nd & 2) for (int i = 0; i < 4; i++) L(i) = src[-1 + i * st];
    if (nd & 4) T(-1) = L(-1) = src[-1 - st];
    if (nd & 8) for (int i = 0; i < 4; i++) T(4 + i) = topright[i];
}
static void load_edges8(Edges *e, const px *src, int st, int mode, int has_tl, int has_tr)
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
        if (has_tr) { for (int i = 8; i < 15; i++) T(i) = f3(P(i - 1, -1), P(i, -1), P(i + 1, -1)); T(15) = (P(14, -1) + 3 * P(15, -1) + 2) >> 2; }
        else for (int i = 8; i < 16; i++) T(i) = P(7, -1);
    }
    if (nd & 2) {
        L(0) = f3(has_tl ? P(-1, -1) : P(-1, 0), P(-1, 0), P(-1, 1));
        for (int i = 1; i < 7; i++) L(i) = f3(P(-1, i - 1), P(-1, i), P(-1, i + 1));
        L(7) = (P(-1, 6) + 3 * P(-1, 7) + 2) >> 2;
    }
    if (nd & 4) T(-1) = L(-1) = f3(P(-1, 0), P(-1, -1), P(0, -1));
#undef P
}
static void fill(px *d, int st, int w, int h, int v) { for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) d[y * st + x] = (px)v; }

void orc_h264_hbd_pred(int bits, int tab, int mode, uint8_t *srcp, const uint8_t *toprightp, int has_topleft, int has_topright, ptrdiff_t stride)
{
    px *src = (px *)srcp; const px *topright = (const px *)toprightp;
    const int st = (int)(stride / 2), mid = 1 << (bits - 1), top = (1 << bits) - 1;
    Edges e;
    if (tab == 0 || tab == 1) {
        const int n = tab ? 8 : 4;
        px out[64];
        if (tab) load_edges8(&e, src, st, mode, has_topleft, has_topright); else load_edges4(&e, src, topright, st, mode);
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) out[x + n * y] = (px)directional(&e, n, mode, x, y, mid);
        for (int y = 0; y < n; y++) memcpy(src + y * st, out + n * y, (size_t)n * 2);
        return;
    }
    const int n = tab == 2 ? 8 : 16;
    int sum_t[4] = { 0, 0, 0, 0 }, sum_l[4] = { 0, 0, 0, 0 };
    const int chroma_dc = tab == 2 && (mode == 0 || mode == 4 || mode == 5 || mode >= 7);
    const int read_t = tab == 2 ? (mode == 0 || mode == 5 || mode == 7 || mode == 8) : (mode == 0 || mode == 5);
    const int read_l = tab == 2 ? (mode == 0 || mode == 4 || mode >= 7) : (mode == 0 || mode == 4);
    if (read_t) for (int i = 0; i < n; i++) sum_t[i >> 2] += src[i - st];
    const int nl = (tab == 2 && mode == 7) ? 4 : n;
    if (read_l) for (int i = 0; i < nl; i++) sum_l[i >> 2] += src[-1 + i * st];
    if (mode == 1) { for (int y = 0; y < n; y++) { const px v = src[-1 + y * st]; for (int x = 0; x < n; x++) src[x + y * st] = v; } return; }
    if (mode == 2) { px t[16]; memcpy(t, src - st, (size_t)n * 2); for (int y = 0; y < n; y++) memcpy(src + y * st, t, (size_t)n * 2); return; }
    if (mode == 3) {
        const int h = n / 2;
        int H = 0, V = 0;
        for (int k = 1; k <= h; k++) { H += k * (src[h - 1 + k - st] - src[h - 1 - k - st]); V += k * (src[-1 + (h - 1 + k) * st] - src[-1 + (h - 1 - k) * st]); }
        if (n == 8) { H = (17 * H + 16) >> 5; V = (17 * V + 16) >> 5; } else { H = (5 * H + 32) >> 6; V = (5 * V + 32) >> 6; }
        const int a = 16 * (src[-1 + (n - 1) * st] + src[n - 1 - st] + 1) - (h - 1) * (V + H);
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) { int v = (a + x * H + y * V) >> 5; src[x + y * st] = (px)(v < 0 ? 0 : v > top ? top : v); }
        return;
    }
    if (mode == 6) { fill(src, st, n, n, mid); return; }
    if (tab == 3) {
        const int T4 = sum_t[0] + sum_t[1] + sum_t[2] + sum_t[3], L4 = sum_l[0] + sum_l[1] + sum_l[2] + sum_l[3];
        fill(src, st, 16, 16, mode == 0 ? (T4 + L4 + 16) >> 5 : mode == 4 ? (L4 + 8) >> 4 : (T4 + 8) >> 4);
        return;
    }
    if (chroma_dc) {
        int q[4];
        const int t0 = sum_t[0], t1 = sum_t[1], l0 = sum_l[0], l1 = sum_l[1];
        switch (mode) {
        case 0:  q[0] = (t0 + l0 + 4) >> 3; q[1] = (t1 + 2) >> 2; q[2] = (l1 + 2) >> 2; q[3] = (t1 + l1 + 4) >> 3; break;
        case 4:  q[0] = q[1] = (l0 + 2) >> 2; q[2] = q[3] = (l1 + 2) >> 2; break;
        case 5:  q[0] = q[2] = (t0 + 2) >> 2; q[1] = q[3] = (t1 + 2) >> 2; break;
        case 7:  q[0] = (t0 + l0 + 4) >> 3; q[2] = (t0 + 2) >> 2; q[1] = q[3] = (t1 + 2) >> 2; break;
        case 8:  q[0] = (t0 + 2) >> 2; q[1] = (t1 + 2) >> 2; q[2] = (l1 + 2) >> 2; q[3] = (t1 + l1 + 4) >> 3; break;
        case 9:  q[0] = q[1] = (l0 + 2) >> 2; q[2] = q[3] = mid; break;
        default: q[0] = q[1] = mid; q[2] = q[3] = (l1 + 2) >> 2; break;
        }
        for (int k = 0; k < 4; k++) fill(src + 4 * (k & 1) + 4 * (k >> 1) * st, st, 4, 4, q[k]);
    }
}

static void add_block(px *pix, int32_t *block, int st, int n, int horizontal, const int *first)
{
    for (int i = 0; i < n; i++) {
        px v = (px)first[i];
        for (int k = 0; k < n; k++) { v = (px)(v + block[horizontal ? i * n + k : k * n + i]); pix[horizontal ? k + i * st : i + k * st] = v; }
    }
    memset(block, 0, sizeof(int32_t) * n * n);
}
/* block_offset[] in BYTES like the decoder's for 16-bit samples */
void orc_h264_hbd_pred_add(int bits, int tab, int mode, uint8_t *pixp, const int *block_offset, int32_t *block, int has_topleft, int has_topright, ptrdiff_t stride)
{
    (void)bits;
    px *pix = (px *)pixp; const int st = (int)(stride / 2);
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
            px *p = (px *)(pixp + block_offset[b]);
            for (int i = 0; i < 4; i++) first[i] = mode ? p[-1 + i * st] : p[i - st];
            add_block(p, block + 16 * b, st, 4, mode, first);
        }
    }
}

/* ---- chroma_format_idc 2: 8 x 16 chroma blocks (h264pred_template.c:502-838), bits 8 / 9 / 10 ------------------------------------------
 * One expression per sample from the row above (top[0..7], corner) and the column to the left (left[0..15]): the DC family works per
 * 4 x 4 quadrant (2 across, 4 down), plane prediction has an 8-tap vertical gradient (:804-838). */
static int ld(const uint8_t *p, int bits, ptrdiff_t byte_off) { return bits > 8 ? *(const uint16_t *)(p + byte_off) : p[byte_off]; }
static void stp(uint8_t *p, int bits, ptrdiff_t byte_off, int v) { if (bits > 8) *(uint16_t *)(p + byte_off) = (uint16_t)v; else p[byte_off] = (uint8_t)v; }

void orc_h264_pred422(int bits, int mode, uint8_t *src, ptrdiff_t st)
{
    const int sb = bits > 8 ? 2 : 1, mid = 1 << (bits - 1), maxv = (1 << bits) - 1;
    int top[8], left[16], corner, out[16][8];
    const int rt = mode == 0 || mode == 2 || mode == 3 || mode == 5 || mode == 7 || mode == 8;
    const int rl = mode == 0 || mode == 1 || mode == 3 || mode == 4 || mode >= 7;
    const int nl = mode == 7 ? 4 : 16;
    memset(top, 0, sizeof(top)); memset(left, 0, sizeof(left));
    if (rt) for (int i = 0; i < 8; i++) top[i] = ld(src, bits, -st + i * sb);
    if (rl) for (int i = 0; i < nl; i++) left[i] = ld(src, bits, i * st - sb);
    corner = mode == 3 ? ld(src, bits, -st - sb) : 0;
    int t0 = 0, t1 = 0, l[4] = { 0, 0, 0, 0 }, H = 0, V = 0, a = 0;
    for (int i = 0; i < 4; i++) { t0 += top[i]; t1 += top[4 + i]; }
    for (int i = 0; i < 16; i++) l[i >> 2] += left[i];
    if (mode == 3) {
        for (int k = 1; k <= 4; k++) H += k * (top[3 + k] - (3 - k < 0 ? corner : top[3 - k]));
        for (int k = 1; k <= 8; k++) V += k * (left[7 + k] - (7 - k < 0 ? corner : left[7 - k]));
        H = (17 * H + 16) >> 5; V = (5 * V + 32) >> 6;
        a = 16 * (left[15] + top[7] + 1) - 7 * V - 3 * H;
    }
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 8; x++) {
            const int c = x >> 2, r = y >> 2;
            const int dc = !c ? (!r ? (t0 + l[0] + 4) >> 3 : (l[r] + 2) >> 2) : (!r ? (t1 + 2) >> 2 : (t1 + l[r] + 4) >> 3);     /* pred8x16_dc, :673-720 */
            const int ldc = (l[r] + 2) >> 2, tdc = ((c ? t1 : t0) + 2) >> 2;
            int v;
            switch (mode) {
            case 0: v = dc; break;
            case 1: v = left[y]; break;
            case 2: v = top[x]; break;
            case 3: v = (a + x * H + y * V) >> 5; v = v < 0 ? 0 : v > maxv ? maxv : v; break;
            case 4: v = ldc; break;
            case 5: v = tdc; break;
            case 6: v = mid; break;
            case 7: v = (!c && !r) ? (t0 + l[0] + 4) >> 3 : tdc; break;      /* top_dc, then pred4x4_dc on the first block */
            case 8: v = (!c && !r) ? (t0 + 2) >> 2 : dc; break;              /* dc, then pred4x4_top_dc on the first block */
            case 9: v = r == 1 ? mid : ldc; break;                           /* left_dc, then 128 on the two blocks of the second row */
            default: v = r == 0 ? mid : ldc; break;                          /* left_dc, then 128 on the two blocks of the first row */
            }
            out[y][x] = v;
        }
    for (int y = 0; y < 16; y++) for (int x = 0; x < 8; x++) stp(src, bits, y * st + x * sb, out[y][x]);
}

/* pred8x16_vertical_add / _horizontal_add (:1326-1354): eight 4 x 4 blocks, the lower four through block_offset[i + 4] */
void orc_h264_pred422_add(int bits, int add_mode, uint8_t *pix, const int *block_offset, void *block, ptrdiff_t st)
{
    const int sb = bits > 8 ? 2 : 1, mask = bits > 8 ? 0xffff : 0xff;
    for (int b = 0; b < 8; b++) {
        uint8_t *p = pix + block_offset[b < 4 ? b : b + 4];
        for (int i = 0; i < 4; i++) {
            int v = add_mode ? ld(p, bits, i * st - sb) : ld(p, bits, -st + i * sb);
            for (int k = 0; k < 4; k++) {
                const int ci = 16 * b + (add_mode ? i * 4 + k : k * 4 + i);
                v = (v + (bits > 8 ? ((int32_t *)block)[ci] : ((int16_t *)block)[ci])) & mask;
                stp(p, bits, add_mode ? i * st + k * sb : k * st + i * sb, v);
            }
        }
        if (bits > 8) memset((int32_t *)block + 16 * b, 0, 64); else memset((int16_t *)block + 16 * b, 0, 32);
    }
}
