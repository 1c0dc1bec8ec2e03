/*
 * oracle/refbuild/refapi_h264pic.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * BASELINE config 3 on the host cores: the reference's own H.264 DSP tables (ff_h264dsp_init, ff_h264qpel_init,
 * ff_h264chroma_init, ff_videodsp_init -- unmodified sources) driven over a batch of synthetic pictures in the order
 * hl_decode_mb() and loop_filter() impose (libavcodec/h264_mb_template.c:40-260, h264_mb.c:204-320 mc_dir_part / mc_part_std,
 * h264_slice.c:1972-2066): per macroblock motion compensation of its partitions (emulated_edge_mc when a block leaves the
 * picture, h264_mb.c:227-246), then h264_idct_add16 / add16intra / idct8_add4 and h264_idct_add8; per slice afterwards the loop
 * filter of its macroblocks in raster order, vertical edges before horizontal ones (h264_loopfilter.c:397-415, filter_mb_edge*
 * :104-236).  Slices are independent work units (disable_deblocking_filter_idc = 2 in the synthetic pictures: no edge crosses a
 * slice) and are handed to pthreads from one counter, the way the reference's slice threading hands them out.
 *
 * It consumes the very record arrays the CUDA batch calls take (include/avdsp_b200.h), so bench.py's `--impl reference
 * --workload h264` arm and tests/test_gpu_h264chain.py feed both sides the same bytes.  No arithmetic of its own.
 */
#define ORC_PREFIX ref_
#include "../oracle_api.h"

/* the record layouts of include/avdsp_b200.h (that header cannot be included next to the reference's: it declares layout-identical
 * tables under the reference's own struct names); tests/test_abi_cpu.py pins the sizes */
typedef struct FFH264MCRecord { int16_t x, y, mvx, mvy; uint8_t w, h, avg, ref; } FFH264MCRecord;
typedef struct FFH264ResidualMB { uint32_t luma_off, chroma_off; uint8_t luma_mode, chroma, pad[2]; } FFH264ResidualMB;
typedef struct FFH264DeblockMB {
    uint8_t alpha[2][4], beta[2][4]; int8_t tc0[2][4][4]; uint8_t intra[2];
    uint8_t calpha[2][2][2], cbeta[2][2][2]; int8_t ctc0[2][2][2][4]; uint8_t cintra[2][2]; uint8_t pad[2];
} FFH264DeblockMB;
typedef char pic_mc_size[sizeof(FFH264MCRecord) == 12 ? 1 : -1];
typedef char pic_res_size[sizeof(FFH264ResidualMB) == 12 ? 1 : -1];
typedef char pic_dbk_size[sizeof(FFH264DeblockMB) == 104 ? 1 : -1];

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "libavutil/cpu.h"
#include "libavutil/mem.h"
#include "libavcodec/h264dsp.h"
#include "libavcodec/h264qpel.h"
#include "libavcodec/h264chroma.h"
#include "libavcodec/videodsp.h"

static pthread_once_t once = PTHREAD_ONCE_INIT;
static H264DSPContext dsp;
static H264QpelContext qpel;
static H264ChromaContext chroma;
static VideoDSPContext vdsp;
static int block_offset_ls = -1;

static void init_tables(void)
{
#ifndef REF_SIMD
    av_set_cpu_flags_mask(0);
#endif
    ff_h264dsp_init(&dsp, 8, 1);
    ff_h264qpel_init(&qpel, 8);
    ff_h264chroma_init(&chroma, 8);
    ff_videodsp_init(&vdsp, 8);
}

struct pic_job {
    const FFH264MCRecord *mc; const uint32_t *mc_first;     /* one picture's records; mc_first[m] .. mc_first[m + 1] belong to macroblock m */
    const FFH264ResidualMB *res; const FFH264DeblockMB *dbk;
    const uint8_t *nnzc;                                    /* one picture: 120 bytes per macroblock */
    int16_t *coeffs;                                        /* the whole batch: 768 per macroblock, consumed */
    const uint8_t *const *refs; int nrefs;                  /* refs[3 * r + plane]: ONE reference picture each, shared by every picture of the batch */
    uint8_t *y, *cb, *cr; int ls, uvls;                     /* destination planes, pictures stacked vertically */
    int mb_w, mb_h, n_pictures, slices;
    int block_offset[48];
    volatile int next;                                      /* next (picture, slice) unit */
};

static void mc_record(const struct pic_job *j, const FFH264MCRecord *r, uint8_t *py, uint8_t *pcb, uint8_t *pcr, uint8_t *emu)
{
    const int pic_w = 16 * j->mb_w, pic_h = 16 * j->mb_h, ls = j->ls, uvls = j->uvls;
    const uint8_t *ry = j->refs[3 * r->ref], *rcb = j->refs[3 * r->ref + 1], *rcr = j->refs[3 * r->ref + 2];
    const int n = r->w < r->h ? r->w : r->h;                 /* the qpel functions are square: a 16x8 partition is two 8x8 calls (h264_mb.c:248-250) */
    const int sidx = n == 16 ? 0 : n == 8 ? 1 : 2;
    const qpel_mc_func *qop = (r->avg ? qpel.avg_h264_qpel_pixels_tab : qpel.put_h264_qpel_pixels_tab)[sidx];
    const int mx = r->mvx + 4 * r->x, my = r->mvy + 4 * r->y;
    const int luma_xy = (mx & 3) + ((my & 3) << 2);
    for (int oy = 0; oy < r->h; oy += n)
        for (int ox = 0; ox < r->w; ox += n) {
            const int fx = (mx >> 2) + ox, fy = (my >> 2) + oy;
            const uint8_t *src = ry + (ptrdiff_t)fy * ls + fx;
            const int ew = (mx & 3) ? 3 : 0, eh = (my & 3) ? 3 : 0;     /* the 6-tap filters reach 2 left / 3 right of the block */
            if (fx < (ew ? 2 : 0) || fy < (eh ? 2 : 0) || fx + n + ew > pic_w || fy + n + eh > pic_h) {
                vdsp.emulated_edge_mc(emu, src - 2 - 2 * ls, ls, ls, n + 5, n + 5, fx - 2, fy - 2, pic_w, pic_h);
                src = emu + 2 + 2 * ls;
            }
            qop[luma_xy](py + (ptrdiff_t)(r->y + oy) * ls + r->x + ox, src, ls);
        }
    /* chroma: one bilinear call per plane over the whole partition (mc_dir_part, h264_mb.c:287-318) */
    const int cw = r->w >> 1, ch = r->h >> 1, cx = mx >> 3, cy = my >> 3;
    const int widx = cw == 8 ? 0 : cw == 4 ? 1 : 2;
    h264_chroma_mc_func cop = (r->avg ? chroma.avg_h264_chroma_pixels_tab : chroma.put_h264_chroma_pixels_tab)[widx];
    const int cemu = cx < 0 || cy < 0 || cx + cw + 1 > (pic_w >> 1) || cy + ch + 1 > (pic_h >> 1);
    uint8_t *const dst[2] = { pcb, pcr };
    const uint8_t *const rp[2] = { rcb, rcr };
    for (int p = 0; p < 2; p++) {
        const uint8_t *src = rp[p] + (ptrdiff_t)cy * uvls + cx;
        if (cemu) {
            vdsp.emulated_edge_mc(emu, src, uvls, uvls, cw + 1, ch + 1, cx, cy, pic_w >> 1, pic_h >> 1);
            src = emu;
        }
        cop(dst[p] + (ptrdiff_t)(r->y >> 1) * uvls + (r->x >> 1), (uint8_t *)src, uvls, ch, mx & 7, my & 7);
    }
}

static void deblock_mb(const struct pic_job *j, const FFH264DeblockMB *r, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    const int ls = j->ls, uvls = j->uvls;
    uint8_t *const cpl[2] = { cb, cr };
    for (int d = 0; d < 2; d++)
        for (int e = 0; e < 4; e++) {
            if (r->alpha[d][e] && r->beta[d][e]) {
                uint8_t *pix = y + (d ? 4 * e * ls : 4 * e);
                int8_t tc[4]; memcpy(tc, r->tc0[d][e], 4);
                if ((r->intra[d] >> e) & 1) {
                    if (d) dsp.h264_v_loop_filter_luma_intra(pix, ls, r->alpha[d][e], r->beta[d][e]);
                    else   dsp.h264_h_loop_filter_luma_intra(pix, ls, r->alpha[d][e], r->beta[d][e]);
                } else {
                    if (d) dsp.h264_v_loop_filter_luma(pix, ls, r->alpha[d][e], r->beta[d][e], tc);
                    else   dsp.h264_h_loop_filter_luma(pix, ls, r->alpha[d][e], r->beta[d][e], tc);
                }
            }
            if (e & 1) continue;
            const int ce = e >> 1;
            for (int p = 0; p < 2; p++) {
                const int a = r->calpha[p][d][ce], b = r->cbeta[p][d][ce];
                if (!a || !b) continue;
                uint8_t *pix = cpl[p] + (d ? 4 * ce * uvls : 4 * ce);
                int8_t tc[4]; memcpy(tc, r->ctc0[p][d][ce], 4);
                if ((r->cintra[p][d] >> ce) & 1) {
                    if (d) dsp.h264_v_loop_filter_chroma_intra(pix, uvls, a, b);
                    else   dsp.h264_h_loop_filter_chroma_intra(pix, uvls, a, b);
                } else {
                    if (d) dsp.h264_v_loop_filter_chroma(pix, uvls, a, b, tc);
                    else   dsp.h264_h_loop_filter_chroma(pix, uvls, a, b, tc);
                }
            }
        }
}

static void *pic_worker(void *arg)
{
    struct pic_job *j = arg;
    const int n_mb = j->mb_w * j->mb_h, per = (n_mb + j->slices - 1) / j->slices, units = j->n_pictures * j->slices;
    uint8_t *emu = av_malloc((size_t)j->ls * 24 + 64);
    if (!emu) return NULL;
    for (;;) {
        const int u = __sync_fetch_and_add(&j->next, 1);
        if (u >= units) break;
        const int pic = u / j->slices, sl = u % j->slices;
        const int m0 = sl * per, m1 = m0 + per < n_mb ? m0 + per : n_mb;
        uint8_t *py = j->y + (size_t)pic * 16 * j->mb_h * j->ls;
        uint8_t *pcb = j->cb + (size_t)pic * 8 * j->mb_h * j->uvls, *pcr = j->cr + (size_t)pic * 8 * j->mb_h * j->uvls;
        for (int m = m0; m < m1; m++) {                       /* hl_decode_mb(): prediction, then the residual */
            for (uint32_t k = j->mc_first[m]; k < j->mc_first[m + 1]; k++) mc_record(j, j->mc + k, py, pcb, pcr, emu);
            const FFH264ResidualMB *r = j->res + m;
            int16_t *blk = j->coeffs + ((size_t)pic * n_mb + m) * 768;
            const uint8_t *nz = j->nnzc + (size_t)m * 120;
            if (r->luma_mode == 0)      dsp.h264_idct_add16(py + r->luma_off, j->block_offset, blk, j->ls, nz);
            else if (r->luma_mode == 1) dsp.h264_idct_add16intra(py + r->luma_off, j->block_offset, blk, j->ls, nz);
            else if (r->luma_mode == 2) dsp.h264_idct8_add4(py + r->luma_off, j->block_offset, blk, j->ls, nz);
            if (r->chroma) {
                uint8_t *d2[2] = { pcb + r->chroma_off, pcr + r->chroma_off };
                dsp.h264_idct_add8(d2, j->block_offset, blk, j->uvls, nz);
            }
        }
        if (j->dbk)
            for (int m = m0; m < m1; m++) {                   /* loop_filter() over the slice's macroblocks */
                const int mbx = m % j->mb_w, mby = m / j->mb_w;
                deblock_mb(j, j->dbk + m, py + (size_t)mby * 16 * j->ls + mbx * 16, pcb + (size_t)mby * 8 * j->uvls + mbx * 8,
                           pcr + (size_t)mby * 8 * j->uvls + mbx * 8);
            }
    }
    av_free(emu);
    return NULL;
}

/* One picture's work (mc / mc_first / res / dbk / nnzc: records of ONE 16 mb_w x 16 mb_h picture, luma_off / chroma_off relative to
 * the picture's own planes) applied to every one of n_pictures pictures stacked vertically in y / cb / cr, each with its own
 * coefficient arena (coeffs + picture * mb_w * mb_h * 768, consumed).  refs: nrefs x 3 plane pointers of single reference pictures
 * with the destination's pitches.  Returns 0, -1 on bad arguments. */
int ref_h264_pictures(const void *mc, const uint32_t *mc_first, const void *res, const void *dbk, const uint8_t *nnzc, int16_t *coeffs,
                      const uint8_t *const *refs, int nrefs, uint8_t *y, uint8_t *cb, uint8_t *cr, int ls, int uvls,
                      int mb_w, int mb_h, int n_pictures, int slices, int nthreads)
{
    pthread_once(&once, init_tables);
    if (!mc_first || !res || !coeffs || !nnzc || mb_w <= 0 || mb_h <= 0 || n_pictures <= 0 || slices <= 0 || nrefs <= 0) return -1;
    struct pic_job j;
    memset(&j, 0, sizeof(j));
    j.mc = mc; j.mc_first = mc_first; j.res = res; j.dbk = dbk; j.nnzc = nnzc; j.coeffs = coeffs; j.refs = refs; j.nrefs = nrefs;
    j.y = y; j.cb = cb; j.cr = cr; j.ls = ls; j.uvls = uvls; j.mb_w = mb_w; j.mb_h = mb_h; j.n_pictures = n_pictures; j.slices = slices;
    for (int i = 0; i < 16; i++)                            /* frame-macroblock block_offset[] (h264_slice.c:486-493) */
        j.block_offset[i] = 4 * ((i & 1) + 2 * ((i >> 2) & 1)) + 4 * (((i >> 1) & 1) + 2 * (i >> 3)) * ls;
    for (int i = 0; i < 4; i++)
        j.block_offset[16 + i] = j.block_offset[32 + i] = 4 * (i & 1) + 4 * ((i >> 1) & 1) * uvls;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) { pic_worker(&j); return 0; }
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pic_worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
