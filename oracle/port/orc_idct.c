/*
 * oracle/port/orc_idct.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C) of the reference's 8x8 "simple" inverse DCT family and the
 * clamped pixel helpers that sit in IDCTDSPContext / BlockDSPContext.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or call this file.
 *
 * Pinned (tests/test_oracle_cpu.py) against oracle/_ref (the unmodified reference
 * compiled from /root/reference by oracle/refbuild/build_ref.sh) and against the
 * committed fixtures in tests/golden/.
 *
 * What is restated (reference file:line):
 *   row pass            libavcodec/simple_idct_template.c:88-173  (idctRowCondDC_8)
 *   column pass         libavcodec/simple_idct_template.c:175-223 (IDCT_COLS)
 *   put / add / plain   libavcodec/simple_idct_template.c:225-326
 *   clamp helpers       libavcodec/idctdsp.c:85-145
 *   clear / fill        libavcodec/blockdsp.c:29-58
 * Arithmetic is done in uint32_t where the reference uses int so that the (never
 * exercised by real streams) overflow case wraps like the GPU does instead of being UB.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../oracle_api.h"

enum { C1 = 22725, C2 = 21407, C3 = 19266, C4 = 16383, C5 = 12873, C6 = 8867, C7 = 4520 };
enum { ROW_SH = 11, COL_SH = 20 };

static inline int sra(uint32_t v, int s) { return (int32_t)v >> s; }
static inline uint8_t clamp_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : (uint8_t)v; }

/* One row, in place.  A row whose seven AC terms are all zero takes the exact
 * "dc << 3" shortcut of simple_idct_template.c:94-106 -- value relevant. */
static void row_pass(int16_t *r)
{
    if (!(r[1] | r[2] | r[3] | r[4] | r[5] | r[6] | r[7])) {
        int16_t dc = (int16_t)((uint32_t)r[0] << 3);
        for (int k = 0; k < 8; k++) r[k] = dc;
        return;
    }
    uint32_t x0 = r[0], x1 = r[1], x2 = r[2], x3 = r[3];
    uint32_t x4 = r[4], x5 = r[5], x6 = r[6], x7 = r[7];
    uint32_t base = C4 * x0 + (1u << (ROW_SH - 1));
    uint32_t e[4], o[4];
    e[0] = base + C2 * x2 + C4 * x4 + C6 * x6;
    e[1] = base + C6 * x2 - C4 * x4 - C2 * x6;
    e[2] = base - C6 * x2 - C4 * x4 + C2 * x6;
    e[3] = base - C2 * x2 + C4 * x4 - C6 * x6;
    o[0] = C1 * x1 + C3 * x3 + C5 * x5 + C7 * x7;
    o[1] = C3 * x1 - C7 * x3 - C1 * x5 - C5 * x7;
    o[2] = C5 * x1 - C1 * x3 + C7 * x5 + C3 * x7;
    o[3] = C7 * x1 - C5 * x3 + C3 * x5 - C1 * x7;
    for (int k = 0; k < 4; k++) {
        r[k]     = (int16_t)sra(e[k] + o[k], ROW_SH);   /* int16 truncation kept, :165-172 */
        r[7 - k] = (int16_t)sra(e[k] - o[k], ROW_SH);
    }
}

/* One column of the (row-transformed) block -> eight 32-bit results already >> 20. */
static void col_pass(const int16_t *c, int out[8])
{
    uint32_t x0 = c[0], x1 = c[8], x2 = c[16], x3 = c[24];
    uint32_t x4 = c[32], x5 = c[40], x6 = c[48], x7 = c[56];
    /* rounding folded into the DC term: (1<<19)/16383 == 32, :176 */
    uint32_t base = C4 * (x0 + ((1u << (COL_SH - 1)) / C4));
    uint32_t e[4], o[4];
    e[0] = base + C2 * x2 + C4 * x4 + C6 * x6;
    e[1] = base + C6 * x2 - C4 * x4 - C2 * x6;
    e[2] = base - C6 * x2 - C4 * x4 + C2 * x6;
    e[3] = base - C2 * x2 + C4 * x4 - C6 * x6;
    o[0] = C1 * x1 + C3 * x3 + C5 * x5 + C7 * x7;
    o[1] = C3 * x1 - C7 * x3 - C1 * x5 - C5 * x7;
    o[2] = C5 * x1 - C1 * x3 + C7 * x5 + C3 * x7;
    o[3] = C7 * x1 - C5 * x3 + C3 * x5 - C1 * x7;
    for (int k = 0; k < 4; k++) {
        out[k]     = sra(e[k] + o[k], COL_SH);
        out[7 - k] = sra(e[k] - o[k], COL_SH);
    }
}

void orc_simple_idct_put(uint8_t *dst, ptrdiff_t stride, int16_t *block)
{
    int v[8];
    for (int y = 0; y < 8; y++) row_pass(block + 8 * y);
    for (int x = 0; x < 8; x++) {
        col_pass(block + x, v);
        for (int y = 0; y < 8; y++) dst[y * stride + x] = clamp_u8(v[y]);
    }
}

void orc_simple_idct_add(uint8_t *dst, ptrdiff_t stride, int16_t *block)
{
    int v[8];
    for (int y = 0; y < 8; y++) row_pass(block + 8 * y);
    for (int x = 0; x < 8; x++) {
        col_pass(block + x, v);
        for (int y = 0; y < 8; y++) dst[y * stride + x] = clamp_u8(dst[y * stride + x] + v[y]);
    }
}

void orc_simple_idct(int16_t *block)
{
    int v[8];
    for (int y = 0; y < 8; y++) row_pass(block + 8 * y);
    for (int x = 0; x < 8; x++) {
        col_pass(block + x, v);
        for (int y = 0; y < 8; y++) block[8 * y + x] = (int16_t)v[y];
    }
}

/* idctdsp.c:85-145 */
void orc_put_pixels_clamped(const int16_t *block, uint8_t *pixels, ptrdiff_t stride)
{
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) pixels[y * stride + x] = clamp_u8(block[8 * y + x]);
}

void orc_put_signed_pixels_clamped(const int16_t *block, uint8_t *pixels, ptrdiff_t stride)
{
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) pixels[y * stride + x] = clamp_u8(block[8 * y + x] + 128);
}

void orc_add_pixels_clamped(const int16_t *block, uint8_t *pixels, ptrdiff_t stride)
{
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            pixels[y * stride + x] = clamp_u8(pixels[y * stride + x] + block[8 * y + x]);
}

/* blockdsp.c:29-58 */
void orc_clear_block(int16_t *block)   { memset(block, 0, 64 * sizeof(int16_t)); }
void orc_clear_blocks(int16_t *blocks) { memset(blocks, 0, 6 * 64 * sizeof(int16_t)); }
void orc_fill_block(int w16, uint8_t *block, uint8_t value, ptrdiff_t stride, int h)
{
    for (int y = 0; y < h; y++) memset(block + y * stride, value, w16 ? 16 : 8);
}

/* ---- batch driver (parity at scale + CPU timing); splits the block range over pthreads ---- */
#include <pthread.h>
struct idct_span { int mode; int16_t *blocks; uint8_t *frame; const uint32_t *off; ptrdiff_t stride; size_t lo, hi; };
static void *idct_span_run(void *arg)
{
    struct idct_span *s = arg;
    for (size_t i = s->lo; i < s->hi; i++) {
        int16_t *b = s->blocks + 64 * i;
        if (s->mode == 0)      orc_simple_idct_put(s->frame + s->off[i], s->stride, b);
        else if (s->mode == 1) orc_simple_idct_add(s->frame + s->off[i], s->stride, b);
        else                   orc_simple_idct(b);
    }
    return NULL;
}
void orc_idct_batch(int mode, int16_t *blocks, uint8_t *frame, const uint32_t *dst_off, ptrdiff_t stride,
                    size_t n, int nthreads)
{
    enum { MAXT = 256 };
    pthread_t th[MAXT];
    struct idct_span sp[MAXT];
    if (nthreads < 1) nthreads = 1;
    if (nthreads > MAXT) nthreads = MAXT;
    for (int t = 0; t < nthreads; t++) {
        sp[t] = (struct idct_span){ mode, blocks, frame, dst_off, stride, n * t / nthreads, n * (t + 1) / nthreads };
        if (nthreads > 1) pthread_create(&th[t], NULL, idct_span_run, &sp[t]);
        else idct_span_run(&sp[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ---- the 10-bit instance: simple_idct_template.c with BIT_DEPTH 10 (:63-78: 17-bit constants, ROW_SHIFT 15, COL_SHIFT 20, DC_SHIFT 1,
 * plain int multiplies); 16-bit samples clipped to 10 bits.  Same structure as the 8-bit functions: the DC-only row shortcut
 * (row[0] << 1 broadcast, :94-106), int16 write-back between the passes, column rounding folded into the DC term (:176). */
static void idct10_rows(int16_t *b)
{
    enum { W1 = 90901, W2 = 85627, W3 = 77062, W4 = 65535, W5 = 51491, W6 = 35468, W7 = 18081 };
    for (int r = 0; r < 8; r++) {
        int16_t *row = b + 8 * r;
        if (!(row[1] | row[2] | row[3] | row[4] | row[5] | row[6] | row[7])) {
            const int16_t v = (int16_t)((row[0] * 2) & 0xffff);
            for (int k = 0; k < 8; k++) row[k] = v;
            continue;
        }
        int a0 = W4 * row[0] + (1 << 14), a1 = a0, a2 = a0, a3 = a0;
        a0 += W2 * row[2]; a1 += W6 * row[2]; a2 -= W6 * row[2]; a3 -= W2 * row[2];
        int b0 = W1 * row[1] + W3 * row[3], b1 = W3 * row[1] - W7 * row[3], b2 = W5 * row[1] - W1 * row[3], b3 = W7 * row[1] - W5 * row[3];
        if (row[4] | row[5] | row[6] | row[7]) {
            a0 += W4 * row[4] + W6 * row[6]; a1 += -W4 * row[4] - W2 * row[6]; a2 += -W4 * row[4] + W2 * row[6]; a3 += W4 * row[4] - W6 * row[6];
            b0 += W5 * row[5] + W7 * row[7]; b1 += -W1 * row[5] - W5 * row[7]; b2 += W7 * row[5] + W3 * row[7]; b3 += W3 * row[5] - W1 * row[7];
        }
        row[0] = (int16_t)((a0 + b0) >> 15); row[7] = (int16_t)((a0 - b0) >> 15); row[1] = (int16_t)((a1 + b1) >> 15); row[6] = (int16_t)((a1 - b1) >> 15);
        row[2] = (int16_t)((a2 + b2) >> 15); row[5] = (int16_t)((a2 - b2) >> 15); row[3] = (int16_t)((a3 + b3) >> 15); row[4] = (int16_t)((a3 - b3) >> 15);
    }
}
static void idct10_col(const int16_t *col, int out[8])
{
    enum { W1 = 90901, W2 = 85627, W3 = 77062, W4 = 65535, W5 = 51491, W6 = 35468, W7 = 18081 };
    int a0 = W4 * (col[0] + ((1 << 19) / W4)), a1 = a0, a2 = a0, a3 = a0;
    a0 += W2 * col[16]; a1 += W6 * col[16]; a2 -= W6 * col[16]; a3 -= W2 * col[16];
    int b0 = W1 * col[8] + W3 * col[24], b1 = W3 * col[8] - W7 * col[24], b2 = W5 * col[8] - W1 * col[24], b3 = W7 * col[8] - W5 * col[24];
    a0 += W4 * col[32]; a1 -= W4 * col[32]; a2 -= W4 * col[32]; a3 += W4 * col[32];
    b0 += W5 * col[40]; b1 -= W1 * col[40]; b2 += W7 * col[40]; b3 += W3 * col[40];
    a0 += W6 * col[48]; a1 -= W2 * col[48]; a2 += W2 * col[48]; a3 -= W6 * col[48];
    b0 += W7 * col[56]; b1 -= W5 * col[56]; b2 += W3 * col[56]; b3 -= W1 * col[56];
    out[0] = (a0 + b0) >> 20; out[1] = (a1 + b1) >> 20; out[2] = (a2 + b2) >> 20; out[3] = (a3 + b3) >> 20;
    out[4] = (a3 - b3) >> 20; out[5] = (a2 - b2) >> 20; out[6] = (a1 - b1) >> 20; out[7] = (a0 - b0) >> 20;
}
void orc_simple_idct10(int mode, uint8_t *dst, ptrdiff_t stride, int16_t *block)
{
    uint16_t *d = (uint16_t *)dst;
    const ptrdiff_t st = stride / 2;
    idct10_rows(block);
    for (int i = 0; i < 8; i++) {
        int o[8];
        idct10_col(block + i, o);
        for (int k = 0; k < 8; k++) {
            if (mode == 2) { block[i + 8 * k] = (int16_t)o[k]; continue; }
            int v = mode == 1 ? d[i + k * st] + o[k] : o[k];
            d[i + k * st] = (uint16_t)(v < 0 ? 0 : v > 1023 ? 1023 : v);
        }
    }
}
