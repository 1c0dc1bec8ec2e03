/*
 * include/avdsp_b200.h -- C ABI of libavdsp_b200.so, the sm_100a back-end for libav's per-block DSP
 * tables and the libswscale scaler.  Plain C: pointers and sizes only, no C++/torch types.
 *
 * Three layers, all in this one header:
 *   1. runtime      avb200_*            device selection, sticky error, memory/stream helpers
 *   2. batched ops  ff_*_batch_cuda     DEVICE pointers, asynchronous on `stream` (a cudaStream_t passed
 *                                       as void *, NULL = default stream).  These are the fast path:
 *                                       thousands of blocks / whole frames per launch.
 *   3. table hooks  ff_*_init_cuda      fill the reference's own function-pointer tables (structs in
 *                                       avdsp_b200_tables.h are layout-identical to the reference's) with
 *                                       slot functions that take HOST pointers exactly like the C slots
 *                                       (one block per call: upload, launch, download, sync).  They exist
 *                                       so the CUDA path drops in behind ff_*dsp_init()'s arch dispatch
 *                                       (libavcodec/idctdsp.c:183-188 etc.); they are for plumbing and
 *                                       parity, not speed.
 * Error convention: the reference slots return void.  Every entry point here that can fail
 * returns 0 / -1 and records a message (avb200_last_error()); slot functions record it too.  The first CUDA FAILURE sticks until
 * avb200_clear_error(); a REFUSAL (a request this library does not take over: the caller keeps its C path) is held only until a
 * failure arrives, so routine refusals never hide a real fault.  There is NO CPU fallback: without a B200 every call fails loudly.
 */
#ifndef AVDSP_B200_H
#define AVDSP_B200_H

#include <stddef.h>
#include <stdint.h>
#include "avdsp_b200_tables.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ 1. runtime ------------------- */
int         avb200_device_count(void);
int         avb200_init(int device);              /* cudaSetDevice + context warm-up; -1 if no sm_100 GPU */
const char *avb200_last_error(void);              /* "" when no error is pending */
void        avb200_clear_error(void);
void        avb200_set_log_callback(void (*cb)(int av_log_level, const char *msg)); /* av_log-style sink */
void        avb200_set_tuning(const char *key, int value);   /* kernel-variant knobs for profiling; see DESIGN.md */
void       *avb200_malloc(size_t bytes);          /* device memory */
void        avb200_free(void *dptr);
void       *avb200_host_alloc(size_t bytes);      /* pinned host memory */
void        avb200_host_free(void *hptr);
int         avb200_host_register(void *hptr, size_t bytes);
int         avb200_host_unregister(void *hptr);
int         avb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int         avb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int         avb200_memcpy2d_h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream);
int         avb200_memcpy2d_d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream);
int         avb200_memset(void *dst, int value, size_t bytes, void *stream);
void       *avb200_stream_create(void);
void        avb200_stream_destroy(void *stream);
int         avb200_stream_sync(void *stream);
int         avb200_device_sync(void);
void       *avb200_event_create(void);
void        avb200_event_destroy(void *ev);
int         avb200_event_record(void *ev, void *stream);
int         avb200_event_sync(void *ev);
float       avb200_event_elapsed_ms(void *start, void *stop);

/* ------------------------------------------------------------------ 2. batched ops --------------- */
/*
 * Destination addressing shared by the 8x8 block ops: block i is written at
 *     frame + dst_off[i]                                  when dst_off != NULL (byte offsets), else at
 *     frame + (i / tiles_per_row) * 8 * stride + (i % tiles_per_row) * 8   (raster of 8x8 tiles).
 * `blocks` is n * 64 int16_t, block-major, row-major inside a block, 16-byte aligned -- the layout the
 * reference's callers hand to idsp.idct_put (libavcodec/mpegvideo.c:1401-1427).
 */

/* IDCTDSPContext.idct_put / idct_add / idct with idct_algo = FF_IDCT_SIMPLE
 * (libavcodec/simple_idct_template.c:289-326, selected at libavcodec/idctdsp.c:168-173).
 * mode 0 = put, 1 = add, 2 = in-place int16 (frame/dst_off/stride ignored).
 * clear != 0 additionally zeroes each coefficient block (fused BlockDSPContext.clear_block,
 * libavcodec/blockdsp.c:29-32); otherwise `blocks` is left untouched (the reference leaves its
 * row-pass intermediate there; callers must not rely on either). */
int ff_simple_idct_batch_cuda(int mode, int16_t *blocks, uint8_t *frame, const uint32_t *dst_off,
                              ptrdiff_t stride, size_t n, int tiles_per_row, int clear, void *stream);

/* IDCTDSPContext.put_pixels_clamped / put_signed_pixels_clamped / add_pixels_clamped
 * (libavcodec/idctdsp.c:85-145).  mode 0 / 1 / 2. */
int ff_pixels_clamped_batch_cuda(int mode, const int16_t *blocks, uint8_t *frame, const uint32_t *dst_off,
                                 ptrdiff_t stride, size_t n, int tiles_per_row, void *stream);

/* ff_simple_idct_put_10 / _add_10 / _10 (libavcodec/simple_idct_template.c, BIT_DEPTH 10; idctdsp.c:151-155) over n blocks: mode 0 put, 1 add,
 * 2 in place; block i -> the 8 x 8 16-bit samples at frame + dst_off[i] (bytes), rows `stride` bytes apart.  The blocks are left as the C
 * functions leave them.  Functional path for 10-bit content (thread per block), not tuned like the 8-bit kernel. */
int ff_simple_idct10_batch_cuda(int mode, int16_t *blocks, uint8_t *frame, const uint32_t *dst_off, ptrdiff_t stride, size_t n, void *stream);

/* BlockDSPContext.clear_block / clear_blocks over n_blocks * 64 coefficients (blockdsp.c:29-37) and
 * fill_block_tab[w16 ? 0 : 1] over n records (blockdsp.c:39-58). */
int ff_clear_blocks_batch_cuda(int16_t *blocks, size_t n_blocks, void *stream);
int ff_fill_blocks_batch_cuda(uint8_t *frame, const uint32_t *dst_off, const uint8_t *value, ptrdiff_t stride,
                              int h, int w16, size_t n, void *stream);

/* Same op with HOST buffers: uploads `blocks` (and `frame` for mode 1), runs the batch in chunks that
 * overlap H2D / kernel / D2H on three streams, downloads frame (or blocks for mode 2) and synchronises.
 * This is the end-to-end call a caller without device-resident data makes.  frame_bytes = size of the
 * host frame buffer that the offsets address. */
int ff_simple_idct_batch_host_cuda(int mode, int16_t *blocks, uint8_t *frame, size_t frame_bytes,
                                   const uint32_t *dst_off, ptrdiff_t stride, size_t n, int tiles_per_row);

/* ---- H.264 (8 bit, 4:2:0, frame macroblocks) ------------------------------------------------------------
 * The reference calls these slots once per block / partition / edge from its macroblock loop
 * (libavcodec/h264_mb_template.c:40-, h264_mb.c:204-320, h264_loopfilter.c:238-418).  The batched entry points take
 * the same information as arrays of per-macroblock records plus the picture planes, all DEVICE memory. */

/* Inverse quantisation, alone and fused in front of the simple IDCT: MpegEncContext.dct_unquantize_{mpeg1,mpeg2,h263}_
 * {intra,inter} (libavcodec/mpegvideo.c:51-270) as used by put_dct() / add_dequant_dct() (mpegvideo.c:1401-1427).
 *   kind 0 mpeg1_intra  1 mpeg1_inter  2 mpeg2_intra  3 mpeg2_intra_bitexact (mismatch control, what
 *        AV_CODEC_FLAG_BITEXACT installs, mpegvideo.c:281-282)  4 mpeg2_inter  5 h263_intra  6 h263_inter
 * One record per block carries what the C functions read from the MpegEncContext for that block.
 * ff_mpeg_dequant_batch_cuda rewrites the blocks in place (inter kinds leave a block with last_index < 0 untouched,
 * like add_dequant_dct); ff_mpeg_dequant_idct_batch_cuda = dequantise + ff_simple_idct_put (intra kinds) or
 * ff_simple_idct_add (inter kinds; blocks with last_index < 0 are skipped) in one kernel, destinations addressed like
 * ff_simple_idct_batch_cuda.  `tables` is a host struct, everything else device memory.
 * Domain: any int16 level, qscale <= 112, matrix entries <= 255 (so |level| * qscale * matrix stays below 2^31, as it
 * does for every conforming stream); outside it the C code's 32-bit wrap-around is not reproduced. */
typedef struct FFMpegDequantBlock {
    uint8_t qscale;
    int8_t  last_index;    /* s->block_last_index[n] */
    uint8_t dc_scale;      /* n < 4 ? s->y_dc_scale : s->c_dc_scale (intra kinds) */
    uint8_t flags;         /* bit 0: s->ac_pred (h263_intra) */
} FFMpegDequantBlock;
typedef struct FFMpegDequantTables {
    uint16_t intra_matrix[64], inter_matrix[64];   /* s->intra_matrix / s->inter_matrix, raster order */
    uint8_t  permutated[64];                       /* s->intra_scantable.permutated (idctdsp.c:28-47) */
    uint8_t  raster_end[64];                       /* s->inter_scantable.raster_end */
    int      alternate_scan, h263_aic;
} FFMpegDequantTables;
int ff_mpeg_dequant_batch_cuda(int kind, const FFMpegDequantTables *tables, const FFMpegDequantBlock *recs, int16_t *blocks,
                               size_t n, void *stream);
int ff_mpeg_dequant_idct_batch_cuda(int kind, const FFMpegDequantTables *tables, const FFMpegDequantBlock *recs, int16_t *blocks,
                                    uint8_t *frame, const uint32_t *dst_off, ptrdiff_t stride, size_t n, int tiles_per_row,
                                    int clear, void *stream);

/* Residual: H264DSPContext.h264_idct_add16 / _add16intra / idct8_add4 (+ h264_idct_add8 for chroma)
 * (libavcodec/h264idct_template.c:174-214).  Record i consumes coeffs + i * coeff_stride (int16; block k of the MB at
 * + 16 * k, luma k = 0..15, cb 16..19, cr 32..35 -- the reference's sl->mb layout) and nnzc + i * 120
 * (non_zero_count_cache, scan8 addressing, libavcodec/h264dec.h:631-645).  Block offsets are the frame-MB defaults
 * (libavcodec/h264_slice.c:486-493).  Consumed coefficients are zeroed exactly as the C functions zero them. */
typedef struct FFH264ResidualMB {
    uint32_t luma_off;     /* byte offset of the MB origin in the luma plane */
    uint32_t chroma_off;   /* byte offset of the MB origin in the cb / cr planes */
    uint8_t  luma_mode;    /* 0 idct_add16, 1 idct_add16intra, 2 idct8_add4, 3 no luma residual */
    uint8_t  chroma;       /* != 0: run h264_idct_add8 on cb and cr */
    uint8_t  pad[2];
} FFH264ResidualMB;
int ff_h264_idct_add_mb_batch_cuda(const FFH264ResidualMB *mbs, size_t n, int16_t *coeffs, size_t coeff_stride,
                                   const uint8_t *nnzc, uint8_t *luma, uint8_t *cb, uint8_t *cr, int linesize,
                                   int uvlinesize, void *stream);

/* DC transforms + dequantisation, the step hl_decode_mb() runs before the residual of a macroblock:
 * H264DSPContext.h264_luma_dc_dequant_idct (intra 16x16; libavcodec/h264_mb.c:714-719, h264idct_template.c:242-275) from
 * the macroblock's 16 luma DC levels (sl->mb_luma_dc, luma_dc + 16 * i) into the DC positions of its coefficient arena, and
 * h264_chroma_dc_dequant_idct (h264_mb_template.c:182-189, h264idct_template.c:304-324) in place on cb (coeffs + 256) and
 * cr (+ 512).  A qmul of 0 skips that transform (the C code's non_zero_count_cache[...DC_BLOCK_INDEX] test).  Run it on the
 * stream ahead of ff_h264_idct_add_mb_batch_cuda / ff_h264_intra_mb_batch_cuda. */
typedef struct FFH264DCRecord {
    uint32_t luma_qmul;        /* pps->dequant4_coeff[0][qscale][0], 0 = no luma DC block */
    uint32_t chroma_qmul[2];   /* pps->dequant4_coeff[1 + (intra ? 0 : 3)][chroma_qp[0]][0], [2 + ...][chroma_qp[1]][0]; 0 = none */
} FFH264DCRecord;
int ff_h264_dc_dequant_batch_cuda(const FFH264DCRecord *recs, size_t n, int16_t *coeffs, size_t coeff_stride, const int16_t *luma_dc,
                                  void *stream);
/* 8-bit 4:2:2 pictures: the chroma transform is h264_chroma422_dc_dequant_idct (the 2 x 4 DC block, h264idct_template.c:271-306, installed for
 * chroma_format_idc 2 at h264dsp.c:87-90); luma as above */
int ff_h264_dc_dequant_batch_422_cuda(const FFH264DCRecord *recs, size_t n, int16_t *coeffs, size_t coeff_stride, const int16_t *luma_dc,
                                      void *stream);

/* Intra reconstruction: for every intra macroblock of a picture, H264PredContext prediction interleaved with the
 * residual exactly as hl_decode_mb() orders them (libavcodec/h264_mb.c:607-731 hl_decode_mb_predict_luma,
 * h264_mb_template.c:158-197 chroma): intra 4x4 -> per block pred4x4 then idct_add / idct_dc_add; intra 8x8 -> pred8x8l
 * then idct8_add / idct8_dc_add; intra 16x16 -> pred16x16 then h264_idct_add16intra; chroma pred8x8 on cb and cr then
 * h264_idct_add8.  One record per macroblock of the picture in raster order (kind 0 = not intra: its pixels, written by
 * the motion-compensation / residual passes, are only read as neighbours).  Coefficients / nnzc as for
 * ff_h264_idct_add_mb_batch_cuda (DC transforms already applied); consumed coefficients are zeroed like the C code.
 * Rows run as a wavefront two macroblocks behind the row above; progress = mb_h * n_pictures uint32 of scratch.
 * Pictures of a batch are stacked vertically (no prediction across a picture boundary; the records say what is
 * available).  Not covered: transform bypass (the lossless *_add predictors exist as table slots only), MBAFF. */
typedef struct FFH264IntraMB {
    uint8_t  kind;             /* 0 not intra, 1 intra 4x4, 2 intra 4x4 with the 8x8 transform (pred8x8l), 3 intra 16x16 */
    uint8_t  mode16;           /* sl->intra16x16_pred_mode (kind 3) */
    uint8_t  chroma_mode;      /* sl->chroma_pred_mode */
    uint8_t  chroma_residual;  /* != 0: h264_idct_add8 after the chroma prediction (cbp & 0x30) */
    uint8_t  mode4[16];        /* sl->intra4x4_pred_mode_cache[scan8[i]]; kind 2 reads entries 0, 4, 8, 12 */
    uint16_t topleft_samples_available, topright_samples_available;   /* h264_mvpred.h:468-508 */
} FFH264IntraMB;
int ff_h264_intra_mb_batch_cuda(const FFH264IntraMB *mbs, int mb_w, int mb_h, int n_pictures, int16_t *coeffs,
                                size_t coeff_stride, const uint8_t *nnzc, uint8_t *luma, uint8_t *cb, uint8_t *cr,
                                int linesize, int uvlinesize, uint32_t *progress, void *stream);

/* Motion compensation: H264QpelContext put/avg tabs + H264ChromaContext put/avg tabs driven like mc_dir_part()
 * (libavcodec/h264_mb.c:204-320): one record = one luma partition and its two chroma partitions.  (x, y) is the
 * partition's luma position, (mvx, mvy) the quarter-pel vector; luma_xy = (mvx & 3) + 4 * (mvy & 3), chroma phase
 * (mvx & 7, mvy & 7).  Reference samples outside the picture are edge-replicated by clamped addressing, which is what
 * emulated_edge_mc (libavcodec/videodsp_template.c:27-94) / the decoder's padded edges provide.
 * All `put` records run before all `avg` records (list 0 then list 1, h264_mb.c:322-366); records of the same kind
 * must address disjoint destination pixels.  A batch of independent pictures is expressed by stacking them vertically
 * (destination and reference planes alike): pic_h is the height of ONE picture and a record with y in
 * [k * pic_h, (k + 1) * pic_h) is clamped to picture k of its reference. */
typedef struct FFH264MCRecord {
    int16_t x, y;          /* luma position of the partition */
    int16_t mvx, mvy;      /* quarter-pel motion vector */
    uint8_t w, h;          /* luma partition size: 16, 8 or 4 each */
    uint8_t avg;           /* 0 put, 1 avg (second prediction direction) */
    uint8_t ref;           /* index into refs[] */
} FFH264MCRecord;
typedef struct FFH264RefPlanes { const uint8_t *y, *cb, *cr; } FFH264RefPlanes;
int ff_h264_mc_batch_cuda(const FFH264MCRecord *recs, size_t n, const FFH264RefPlanes *refs, uint8_t *dst_y, uint8_t *dst_cb,
                          uint8_t *dst_cr, int linesize, int uvlinesize, int pic_w, int pic_h, void *stream);

/* Weighted prediction: H264DSPContext.weight_h264_pixels_tab / biweight_h264_pixels_tab
 * (libavcodec/h264dsp_template.c:30-98) over n rectangles of one plane. */
typedef struct FFH264WeightRecord {
    uint32_t off;          /* byte offset of the block in the plane (dst; src uses the same offset in `src`) */
    uint8_t  w, h;         /* width 16 / 8 / 4 / 2, height */
    uint8_t  log2_denom, pad;
    int16_t  weight, weight_src;   /* weight (uni) or weightd / weights (bi) */
    int16_t  offset, pad2;
} FFH264WeightRecord;
int ff_h264_weight_batch_cuda(const FFH264WeightRecord *recs, size_t n, uint8_t *plane, const uint8_t *src /* NULL = uni */,
                              int stride, void *stream);

/* Deblocking: the 12 loop-filter slots of H264DSPContext (libavcodec/h264dsp_template.c:104-328) applied in the
 * reference's order (macroblocks in raster order; inside one: vertical edges 0..3 then horizontal edges 0..3,
 * libavcodec/h264_loopfilter.c:397-415) to a whole picture.  One record per macroblock, raster order, carrying what
 * filter_mb_edge{v,h,cv,ch} (h264_loopfilter.c:104-236) pass to the slots.  Macroblock rows run as a wavefront that
 * stays two macroblocks behind the row above, so the result is bit-identical to the serial order.
 * dir 0 = vertical edges (h264_h_loop_filter_*), dir 1 = horizontal edges (h264_v_loop_filter_*).
 * An edge with alpha == 0 or beta == 0 is skipped (h264_loopfilter.c:112). */
typedef struct FFH264DeblockMB {
    uint8_t alpha[2][4], beta[2][4];
    int8_t  tc0[2][4][4];          /* per 4-line group, < 0 = group not filtered; ignored on intra edges */
    uint8_t intra[2];              /* bit e set: edge e uses the *_intra (bS = 4) filter */
    uint8_t calpha[2][2][2], cbeta[2][2][2];   /* [plane cb/cr][dir][chroma edge 0/1 = luma edge 0/2] */
    int8_t  ctc0[2][2][2][4];      /* as passed to the chroma slots (already +1), <= 0 = group not filtered */
    uint8_t cintra[2][2];          /* [plane][dir], bit e */
    uint8_t pad[2];
} FFH264DeblockMB;
int ff_h264_deblock_picture_cuda(const FFH264DeblockMB *mbs, int mb_w, int mb_h, uint8_t *luma, uint8_t *cb, uint8_t *cr,
                                 int linesize, int uvlinesize, uint32_t *progress /* 2 * mb_h uint32, scratch */, void *stream);
/* n_pictures independent pictures in one launch: their planes are stacked vertically (picture k starts at luma row
 * k * 16 * mb_h / chroma row k * 8 * mb_h of the same allocations), records likewise (k * mb_w * mb_h);
 * progress holds 2 * mb_h * n_pictures words. */
int ff_h264_deblock_batch_cuda(const FFH264DeblockMB *mbs, int mb_w, int mb_h, int n_pictures, uint8_t *luma, uint8_t *cb,
                               uint8_t *cr, int linesize, int uvlinesize, uint32_t *progress, void *stream);

/* The same three stages for 9 / 10-bit pictures (the BIT_DEPTH > 8 instances, libavcodec/bit_depth_template.c:49-67: uint16 samples,
 * int32 coefficients -- coeffs / coeff_stride count int32 elements, 16 per block like the 8-bit layout; every pitch and every record
 * offset stays in BYTES) and, for the residual and the motion compensation, chroma_format_idc 2 (4:2:2: 8 x 16 chroma macroblocks,
 * h264_idct_add8_422 h264idct_template.c:216-236, chroma vectors at full vertical resolution h264_mb.c:287-316).  Records, order rules
 * and picture stacking are those of the 8-bit calls above; alpha / beta / tc0 in FFH264DeblockMB are the 8-bit-scale table values the
 * decisions produce (the filters scale them by 2^(bit_depth - 8) like h264dsp_template.c:110-113,240).  Field pictures of a frame are
 * pictures in their own right here: pass the field's first row and twice the frame's pitch.  MBAFF frames are not covered.
 * The residual and the motion compensation also take bit_depth 8 with chroma_format_idc 2 (8-bit 4:2:2 pictures: uint8 samples, and `coeffs`
 * then points at int16 coefficients, coeff_stride counting int16); 8-bit 4:2:0 stays with the calls without _hbd, which refuse nothing here. */
int ff_h264_idct_add_mb_batch_hbd_cuda(int bit_depth, int chroma_format_idc, const FFH264ResidualMB *mbs, size_t n, int32_t *coeffs,
                                       size_t coeff_stride, const uint8_t *nnzc, uint8_t *luma, uint8_t *cb, uint8_t *cr, int linesize,
                                       int uvlinesize, void *stream);
int ff_h264_mc_batch_hbd_cuda(int bit_depth, int chroma_format_idc, const FFH264MCRecord *recs, size_t n, const FFH264RefPlanes *refs,
                              uint8_t *dst_y, uint8_t *dst_cb, uint8_t *dst_cr, int linesize, int uvlinesize, int pic_w, int pic_h,
                              void *stream);
int ff_h264_deblock_batch_hbd_cuda(int bit_depth, const FFH264DeblockMB *mbs, int mb_w, int mb_h, int n_pictures, uint8_t *luma,
                                   uint8_t *cb, uint8_t *cr, int linesize, int uvlinesize, void *stream);
/* weighted prediction and the DC transforms for the same pictures: weight_h264_pixels / biweight_h264_pixels at 9 / 10 bit (the record's offset is
 * in 8-bit units and scaled by 2^(bit_depth - 8) like h264dsp_template.c:39,70; `off` and `stride` in bytes) and h264_luma_dc_dequant_idct /
 * h264_chroma_dc_dequant_idct / h264_chroma422_dc_dequant_idct on int32 coefficients (h264idct_template.c:242-324; luma_dc = 16 int32 per macroblock) */
int ff_h264_weight_batch_hbd_cuda(int bit_depth, const FFH264WeightRecord *recs, size_t n, uint8_t *plane, const uint8_t *src /* NULL = uni */,
                                  int stride, void *stream);
int ff_h264_dc_dequant_batch_hbd_cuda(int chroma_format_idc, const FFH264DCRecord *recs, size_t n, int32_t *coeffs, size_t coeff_stride,
                                      const int32_t *luma_dc, void *stream);
/* intra reconstruction at 9 / 10 bit (4:2:0): ff_h264_intra_mb_batch_cuda's records and order, uint16 samples, int32 coefficients; no scratch
 * argument (the wavefront's progress lives in shared memory, one CTA per picture) */
int ff_h264_intra_mb_batch_hbd_cuda(int bit_depth, const FFH264IntraMB *mbs, int mb_w, int mb_h, int n_pictures, int32_t *coeffs,
                                    size_t coeff_stride, const uint8_t *nnzc, uint8_t *luma, uint8_t *cb, uint8_t *cr, int linesize,
                                    int uvlinesize, void *stream);
/* intra reconstruction of 4:2:2 pictures, bit_depth 8 / 9 / 10: the chroma macroblock is 8 x 16 -- pred8x8[] holds the 8 x 16 predictors
 * (h264pred.c:477-563, h264pred_template.c:502-838) and the chroma residual is h264_idct_add8_422 (h264idct_template.c:216-236: blocks 20..23 /
 * 36..39 of the arena, counted at scan8[i + 4]).  coeffs: int16 at 8 bit, int32 at 9 / 10 (coeff_stride in elements of that type). */
int ff_h264_intra_mb_batch_422_cuda(int bit_depth, const FFH264IntraMB *mbs, int mb_w, int mb_h, int n_pictures, void *coeffs,
                                    size_t coeff_stride, const uint8_t *nnzc, uint8_t *luma, uint8_t *cb, uint8_t *cr, int linesize,
                                    int uvlinesize, void *stream);
/* chroma_format_idc 2, bit depth 8 / 9 / 10: the 8 x 16 chroma macroblock has two vertical edges of sixteen lines (the 104-byte record's
 * calpha / cbeta / ctc0 / cintra [plane][0][edge]: h264_h_loop_filter_chroma422, a tc0 entry per four lines) and FOUR horizontal edges, one per
 * luma edge at chroma rows 0, 4, 8, 12 (h264_loopfilter.c:633,693-700: also inside 8x8-transform macroblocks), carried by a second record per
 * macroblock; luma is filtered from the 104-byte record as always ([plane][1][..] of its chroma fields is not read). */
typedef struct FFH264DeblockChroma422 {
    uint8_t alpha[2][4], beta[2][4];   /* [plane cb / cr][edge] */
    int8_t  tc0[2][4][4];              /* as passed to h264_v_loop_filter_chroma (already + 1), <= 0 = group not filtered */
    uint8_t intra[2];                  /* bit e: edge e uses the intra (bS 4) filter */
    uint8_t pad[2];
} FFH264DeblockChroma422;
int ff_h264_deblock_batch_422_cuda(int bit_depth, const FFH264DeblockMB *mbs, const FFH264DeblockChroma422 *chroma422, int mb_w, int mb_h,
                                   int n_pictures, uint8_t *luma, uint8_t *cb, uint8_t *cr, int linesize, int uvlinesize, void *stream);

/* Deblocking DECISIONS: what loop_filter() -> fill_filter_caches() -> ff_h264_filter_mb()
 * (libavcodec/h264_slice.c:1972-2262, libavcodec/h264_loopfilter.c:438-846) decide for every macroblock of a
 * progressive 4:2:0 8-bit picture -- boundary strengths, averaged qp, alpha / beta / tc0 per edge -- computed on the
 * device from the decoder's side-information arrays, written as the FFH264DeblockMB records the deblocking kernels
 * above consume.  All pointers are DEVICE pointers to copies of the decoder's own arrays, in the decoder's layouts:
 *   mb_type, qscale_table, cbp_table, slice_table: [n_pictures * mb_h][mb_w + 1] (mb_stride = mb_w + 1, the extra
 *       column is the decoder's padding and is never read)          H264Picture.mb_type / qscale_table, H264Context
 *   non_zero_count: [..][48] per macroblock                          H264Context.non_zero_count
 *   motion_val[list]: [4 * n_pictures * mb_h][4 * mb_w][2] int16     H264Picture.motion_val, b_stride = 4 * mb_w
 *   ref_index[list]: [..][4] per macroblock (one per 8x8)            H264Picture.ref_index
 *   slices[slice_num]: the slice-header values the filter uses and the slice's ref2frm table (h264dec.h:186,538;
 *       ref2frm[list][2 + ref_index] = identity of the referenced frame, entries 0/1 = -1); at most 32 slice numbers
 *   chroma_qp_table: PPS.chroma_qp_table[2][64] (h264_ps.h:124)
 * Pictures of a batch are stacked row-wise; no edge is filtered across a picture boundary.
 * Field pictures (PAFF) are covered through field_picture, 4:2:2 through chroma422; not covered: MBAFF frames, 4:4:4 (the records
 * themselves are bit-depth agnostic: ff_h264_deblock_batch_hbd_cuda takes them for 9 / 10-bit pictures). */
typedef struct FFH264DeblockSlice {
    int32_t alpha_c0_offset, beta_offset;   /* sl->slice_alpha_c0_offset, sl->slice_beta_offset */
    int32_t deblocking_filter;              /* 0 off, 1 every edge, 2 not across slice boundaries */
    int32_t list_count;                     /* 1 (P) or 2 (B) */
    int32_t qp_thresh;                      /* sl->qp_thresh, h264_slice.c:1815-1818 */
    int32_t ref2frm[2][64];
} FFH264DeblockSlice;
typedef struct FFH264DeblockInfo {
    int mb_w, mb_h, n_pictures;
    const uint32_t *mb_type;
    const int8_t   *qscale_table;
    const uint8_t  *non_zero_count;
    const uint16_t *cbp_table, *slice_table;
    const int16_t  *motion_val[2];
    const int8_t   *ref_index[2];
    const FFH264DeblockSlice *slices;
    int n_slices;
    const uint8_t  *chroma_qp_table;
    int cabac, transform_8x8_mode;          /* PPS.cabac, PPS.transform_8x8_mode */
    int field_picture;                      /* h->picture_structure != PICT_FRAME: the pictures are fields (every mb_type carries
                                               MB_TYPE_INTERLACED): vertical vector limit 2, bS 3 on horizontal intra macroblock edges
                                               (h264_loopfilter.c:551-557,723) */
    FFH264DeblockChroma422 *chroma422;      /* device pointer, NULL for 4:2:0; else chroma_format_idc 2: the horizontal chroma edges of
                                               macroblock m are written to chroma422[m] instead of the 104-byte record's [plane][1][..] */
} FFH264DeblockInfo;
int ff_h264_deblock_params_cuda(const FFH264DeblockInfo *info /* host struct */, FFH264DeblockMB *out, void *stream);

/* The flush point of a decoder back-end: the CPU parses and entropy-decodes, RECORDS the work of a batch of
 * pictures (or of the slices decoded so far) in the arrays below, and flushes it with one call.  The stages run on `stream` in
 * the order hl_decode_mb() (libavcodec/h264_mb_template.c:40-260: inter prediction, weighted prediction, DC transforms,
 * residual; intra prediction interleaved with its residual) and loop_filter() (libavcodec/h264_slice.c:1972-2066) impose; the
 * call returns without synchronising.  Every array is a DEVICE pointer (deblock_info is a host struct of device pointers), a
 * NULL array / zero count skips its stage.  Per-macroblock arrays (dc, residual, intra, deblock_records, the coefficient arena
 * coeffs + i * coeff_stride, nnzc + i * 120) have one entry per macroblock of the batch in raster order, pictures stacked:
 * a macroblock that takes no part in a stage says so in its record (residual luma_mode 3 / chroma 0 for intra and skipped
 * macroblocks, intra kind 0 for inter ones, dc qmul 0).  Planes: n_pictures pictures stacked vertically, 16 * mb_h luma rows
 * each.  progress: 2 * mb_h * n_pictures uint32 of scratch. */
typedef struct FFH264PictureWork {
    int mb_w, mb_h, n_pictures;
    uint8_t *luma, *cb, *cr;                 /* the pictures being reconstructed */
    int linesize, uvlinesize;
    const FFH264MCRecord *mc; size_t n_mc;   /* ff_h264_mc_batch_cuda; pic_w / pic_h = 16 * mb_w / 16 * mb_h */
    const FFH264RefPlanes *refs;
    const FFH264WeightRecord *weight[3]; size_t n_weight[3];   /* per plane: ff_h264_weight_batch_cuda */
    const uint8_t *weight_src[3];            /* bi-weighting: the plane holding the second prediction, else NULL */
    int16_t *coeffs; size_t coeff_stride;    /* coefficient arena, consumed (zeroed) like the C functions do */
    const uint8_t *nnzc;
    const FFH264DCRecord *dc; const int16_t *luma_dc;          /* ff_h264_dc_dequant_batch_cuda */
    const FFH264ResidualMB *residual;        /* ff_h264_idct_add_mb_batch_cuda */
    const FFH264IntraMB *intra;              /* ff_h264_intra_mb_batch_cuda */
    const FFH264DeblockInfo *deblock_info;   /* host struct: decisions are derived on the device into deblock_records */
    FFH264DeblockMB *deblock_records;        /* with deblock_info NULL: records the caller already filled; NULL = no loop filter */
    uint32_t *progress;
    int bit_depth, chroma_format_idc;        /* 0 = 8 / 1.  9 / 10-bit pictures (int32 coeffs / luma_dc behind the same pointers) and 4:2:2 chroma at
                                                8 / 9 / 10 bit run the *_hbd_cuda / *_422_cuda stages in the same order (progress is not used then;
                                                4:2:2 intra macroblocks: ff_h264_intra_mb_batch_422_cuda) */
    const FFH264DeblockChroma422 *deblock_chroma422;   /* 4:2:2 with caller-filled deblock_records: their second record array */
} FFH264PictureWork;
int ff_h264_flush_pictures_cuda(const FFH264PictureWork *work /* host struct */, void *stream);

/* ---- MECmpContext, motion search, HpelDSPContext, FDCTDSPContext -----------------------------------------
 * me_cmp: n block pairs (cur + cur_off vs ref + ref_off, common stride, height h) through one metric; out[i] is
 * what the C slot returns.  kind / sidx / dxy select the slot like the reference's tables (libavcodec/me_cmp.h:39-63):
 *   kind 0 pix_abs[sidx][dxy] (sidx 0 = 16 wide, 1 = 8 wide; dxy = x2 + 2*y2 half-pel flag)   me_cmp.c:109-307
 *        1 sad[sidx]   2 sse[sidx] (sidx 2 = 4 wide)   3 hadamard8_diff[sidx]   4 vsad[0]   5 vsse[0]
 *        6 nsse[sidx] (weight 8: the NULL-context default, me_cmp.c:331)   7 hadamard8_diff[4 + sidx] (intra)
 *        8 vsad[4 + sidx] (intra)   9 vsse[4 + sidx] (intra)   10 sum_abs_dctelem (cur_off in bytes into int16 blocks)
 *       11 dct_sad[sidx]   12 dct_max[sidx]   13 dct264_sad[sidx]   (me_cmp.c:538-621; for 11 / 12 `dxy` names the
 *          FDCTDSPContext.fdct the encoder context holds: 0 = ff_jpeg_fdct_islow_8, 2 = ff_fdct_ifast)
 * These three only dereference the MpegEncContext's DSP tables, so the batch call covers them; their table slots are
 * left to the C code (a slot cannot know which fdct the caller's context selected).  quant_psnr, bit and rd
 * (me_cmp.c:621-782) need the quantiser and VLC length tables of a live encoder: see ff_me_cmp_enc_batch_cuda below. */
typedef struct FFMECmpRecord { uint32_t cur_off, ref_off; } FFMECmpRecord;
int ff_me_cmp_batch_cuda(int kind, int sidx, int dxy, const uint8_t *cur, const uint8_t *ref, ptrdiff_t stride, int h,
                         const FFMECmpRecord *recs, size_t n, int32_t *out, void *stream);

/* quant_psnr / bit / rd (libavcodec/me_cmp.c:621-782: quant_psnr8x8_c, bit8x8_c, rd8x8_c and their 16-wide wrappers :883-885): the three
 * metrics that run the encoder's own quantiser on the difference block -- s->fast_dct_quantize = ff_dct_quantize_c
 * (libavcodec/mpegvideo_enc.c:773-777, :4371-4450: forward DCT, DC division, biased threshold quantiser in scan order) -- then count
 * VLC bits from the codec's run / level length tables (bit, rd) and / or reconstruct through s->dct_unquantize_intra / _inter
 * (libavcodec/mpegvideo.c:51-270) and the simple IDCT (quant_psnr, rd).  FFMECmpEncState = the MpegEncContext fields they read, by value
 * (q_*_matrix = the row of s->q_intra_matrix / s->q_inter_matrix for `qscale`, as ff_convert_matrix() built it, mpegvideo_enc.c:84-160);
 * FFMECmpVlcTables = the codec's static length tables (HOST pointers; [64 * 128] indexed UNI_AC_ENC_INDEX(run, level + 64), the DC table
 * [512] indexed level + 256; may be NULL for quant_psnr, which reads none).  The IDCT is the table's own: FF_IDCT_PERM_NONE
 * (what ff_idctdsp_init_cuda installs), so scantable serves the quantiser and the bit counter alike.
 *   kind 14 quant_psnr (always quantises as an inter block, like the C code)   15 bit   16 rd;   sidx 0 = 16 wide (h 16 or 8), 1 = 8x8
 *   out[i] = the slot's return value; last_index[i] (optional) = s->block_last_index[0] as the C function leaves it.
 * A DC level outside the DC table (|level| > 255: impossible for 8-bit differences with dc_scale >= 1) is clamped to its ends. */
typedef struct FFMECmpEncState {
    int32_t fdct;                  /* s->fdsp.fdct: 0 ff_jpeg_fdct_islow_8, 2 ff_fdct_ifast */
    int32_t dequant;               /* s->dct_unquantize_intra / _inter (mpegvideo_enc.c:1613-1622): 0 mpeg1, 1 mpeg2, 2 mpeg2 under
                                      AV_CODEC_FLAG_BITEXACT (mismatch control on intra blocks too), 3 h263 */
    int32_t qscale, mb_intra, y_dc_scale, h263_aic;
    int32_t intra_quant_bias, inter_quant_bias, ac_esc_length;
    int32_t q_intra_matrix[64], q_inter_matrix[64];
    uint16_t intra_matrix[64], inter_matrix[64];   /* s->intra_matrix / s->inter_matrix (the inverse quantisers' side) */
    uint8_t scantable[64];                         /* s->intra_scantable.scantable */
} FFMECmpEncState;
typedef struct FFMECmpVlcTables {
    const uint8_t *intra_ac_vlc_length, *intra_ac_vlc_last_length, *inter_ac_vlc_length, *inter_ac_vlc_last_length, *luma_dc_vlc_length;
} FFMECmpVlcTables;
/* device-resident copy of one encoder state (one per (qscale, mb_intra) a batch uses); NULL + avb200_last_error() on failure */
void *ff_me_cmp_enc_state_cuda(const FFMECmpEncState *state, const FFMECmpVlcTables *vlc);
void ff_me_cmp_enc_state_free_cuda(void *enc_state);
int ff_me_cmp_enc_batch_cuda(int kind, int sidx, const void *enc_state, const uint8_t *cur, const uint8_t *ref, ptrdiff_t stride, int h,
                             const FFMECmpRecord *recs, size_t n, int32_t *out, int32_t *last_index, void *stream);

/* Exhaustive search (libavcodec/motion_est_template.c:620-655) for every 16x16 macroblock of rows [mb_y0, mb_y1):
 * candidates within +-range (16) clipped so the block stays inside the picture (get_limits, motion_est.c:517-548),
 * score = pix_abs[0][0] (h = 16), penalty_factor 0, strict-< minimum in raster order (mathops.h:133-140).
 * out[3 * (mb_y * (w / 16) + mb_x)] = { mx, my, sad }.  Sharding over GPUs = disjoint row ranges. */
int ff_full_search_cuda(const uint8_t *cur, const uint8_t *ref, int stride, int w, int h, int range, int mb_y0, int mb_y1,
                        int32_t *out, void *stream);

/* HpelDSPContext (libavcodec/hpeldsp.h:45-93): tab 0 put, 1 avg, 2 put_no_rnd, 3 avg_no_rnd; sidx 0..3 = width
 * 16 / 8 / 4 / 2; dxy = x_halfpel + 2 * y_halfpel; h rows.  dst and src share `stride` like the C slots. */
typedef struct FFHpelRecord { uint32_t dst_off, src_off; uint8_t tab, sidx, dxy, h; } FFHpelRecord;
int ff_hpel_batch_cuda(const FFHpelRecord *recs, size_t n, uint8_t *dst, const uint8_t *src, ptrdiff_t stride, void *stream);

/* QpelDSPContext (libavcodec/qpeldsp.h:69-73): MPEG-4 quarter-pel motion compensation, n blocks per launch.
 * kind 0 put_qpel_pixels_tab, 1 put_no_rnd_qpel_pixels_tab, 2 avg_qpel_pixels_tab; sidx 0 = 16x16, 1 = 8x8;
 * mc = x + 4 * y quarter-pel phase (the table index, qpeldsp.c:734-752).  dst and src share `stride` like the C slots;
 * records of one launch must address disjoint destination blocks. */
typedef struct FFQpelRecord { uint32_t dst_off, src_off; uint8_t kind, sidx, mc, pad; } FFQpelRecord;
int ff_mpeg4_qpel_batch_cuda(const FFQpelRecord *recs, size_t n, uint8_t *dst, const uint8_t *src, ptrdiff_t stride, void *stream);

/* FDCTDSPContext (libavcodec/fdctdsp.h:26-29), in place over n blocks: which 0 = ff_jpeg_fdct_islow_8, 1 =
 * ff_fdct248_islow_8 (jfdctint_template.c:260-398), 2 = ff_fdct_ifast, 3 = ff_fdct_ifast248 (jfdctfst.c:207-332), 4 = ff_jpeg_fdct_islow_10,
 * 5 = ff_fdct248_islow_10 (the BIT_DEPTH 10 instances: PASS1_BITS 1, OUT_SHIFT 2; thread per block, untuned). */
int ff_fdct_batch_cuda(int which, int16_t *blocks, size_t n, void *stream);
/* PixblockDSPContext.get_pixels / diff_pixels (libavcodec/pixblockdsp_template.c:24-66) over n 8x8 blocks, optionally
 * followed by the forward DCT in the same kernel -- the encoder's "fetch (difference) -> fdct" front end
 * (mpegvideo_enc.c:1923-2019 dct loop).  Block i reads s1 + off1[i] and, if s2 != NULL, subtracts s2 + off2[i] (off2
 * NULL = the same offsets); which_fdct < 0 stores the samples / differences, 0..3 select the transform as in
 * ff_fdct_batch_cuda.  Output: int16 blocks[n][64], row-major. */
int ff_pixblock_fdct_batch_cuda(int which_fdct, const uint8_t *s1, const uint8_t *s2, const uint32_t *off1, const uint32_t *off2,
                                ptrdiff_t stride, int16_t *blocks, size_t n, void *stream);

/* ---- float FFT / MDCT filterbank (FFTContext, libavcodec/fft.h:73-99); parity contract: 1e-6 relative ------------
 * ff_fft_batch_cuda: n_transforms independent complex FFTs of 2^nbits points (nbits 1..12), in place, natural order in
 * and out == the reference's fft_permute followed by fft_calc (forward kernel exp(-2 pi i jk / n), no 1/n).
 * ff_mdct_batch_cuda: op 0 imdct_half (n/2 in -> n/2 out), 1 imdct_calc (n/2 -> n), 2 mdct_calc (n -> n/2) of size
 * n = 2^nbits (4..14) with the rotation tables of ff_mdct_init(nbits, inverse, scale) (libavcodec/mdct_template.c:86-92). */
int ff_fft_batch_cuda(int nbits, int inverse, float *z, size_t n_transforms, void *stream);
int ff_mdct_batch_cuda(int op, int nbits, double scale, float *out, const float *in, size_t n_transforms, void *stream);

/* ---- libswscale boundary (libswscale/swscale.h:159-207) ------------------------------------------------
 * Same argument lists as sws_getContext / sws_scale / sws_freeContext; pixel formats are the reference's
 * AVPixelFormat values (AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_RGB24 = 2, AV_PIX_FMT_BGR24 = 3), flags the
 * reference's SWS_* bits.  Taken over: planar 8-bit yuv (420p, 422p = 4, 444p = 5, 410p = 6, 411p = 7, 440p = 31; yuva420p = 33 like 420p, src[3] is not read), gray8 = 8 (src[0] only), pal8 = 11 (src[0] indices, src[1] the 256-entry 0xAARRGGBB palette: 1024 bytes, per frame in the batch call), nv12 = 23,
 * nv21 = 24, planar 9 / 10 / 16-bit yuv 420p 422p 444p (LE and BE; scaled, or to packed destinations / another sub-sampling), and the packed yuyv422 = 1, uyvy422 = 15, rgb24, bgr24, argb / rgba / abgr / bgra (to planar yuv, or scaled to rgb24 / bgr24) sources (src[0] only) -> gray8 = 8 (the luma plane alone) / rgb24 / bgr24 / argb = 25 / rgba = 26 / abgr = 27 / bgra = 28 / rgb565 = 37 / bgr565 = 41 / rgb555 = 39 / bgr555 = 43 / rgb444 = 54 / bgr444 = 56 (and their big-endian twins 36 / 40 / 38 / 42 / 55 / 57; from planar or semi-planar yuv sources through the scaler's dithered 16-bpp output stage, output.c:869-902) / yuyv422 / uyvy422 / nv12 / nv21 (dst[0] luma, dst[1] interleaved chroma) / planar yuv (8-bit 420p
 * 422p 444p 410p 411p 440p and their full-range yuvj twins 12 / 13 / 14 / 32 on either side -- a yuv destination of the other range gets the
 * reference's range conversion, planar 8 / 9 / 10-bit only; 9 / 10 / 16-bit 420p = 62 / 64 / 47, 422p = 72 / 66 / 49, 444p = 68 / 70 / 51 little-endian and their big-endian twins), any size, SWS_FULL_CHR_H_INT or not, every scaler
 * algorithm of initFilter (libswscale/utils.c:249-632), results identical to the reference's C path under
 * SWS_ACCURATE_RND | SWS_BITEXACT.  srcFilter / dstFilter: pointers to the reference's SwsFilter (four pointers to { double *coeff; int length; }),
 * taken over for planar 8-bit yuv sources to packed rgb / planar 8-bit yuv destinations, vertical vectors symmetric.  48-bit destinations
 * rgb48be = 34 / rgb48le = 35 / bgr48be = 59 / bgr48le = 60 (yuv2rgb48_X / _2 / _1 on hScale8To19_c lines, libswscale/output.c:593-760; same size without
 * SWS_ACCURATE_RND: yuv2rgb_c_48, yuv2rgb.c:106-236) from planar 8-bit yuv / yuvj sources.  Anything else returns NULL with an error (no fallback).
 *   sws_scale_cuda         HOST pointers; whole frames (srcSliceY = 0, srcSliceH = srcH) with strides of either sign (bottom-up pictures), or
 *                          slices (not into a gray8 destination) like sws_scale() takes them (swscale_unscaled.c:1212-1340): srcSlice[] at source row srcSliceY, dst[] at the
 *                          top of the picture, top-down and contiguous, boundaries on whole chroma rows, positive strides; each call returns
 *                          (and writes) the rows the reference's loop completes with the same slices (swscale.c:483-485), at the cost of a
 *                          whole-frame pass per slice.  Uploads, runs, downloads, synchronises; returns output lines like sws_scale(),
 *                          0 on bad arguments.
 *   sws_scale_frames_cuda  DEVICE pointers, asynchronous on `stream`: nframes frames whose planes lie
 *                          *_frame_stride[] bytes apart (NULL = a single frame); returns lines written or -1.
 *                          For odd dstW the reference writes whole pixel pairs (libswscale/output.c:947); that
 *                          extra pixel is written only when dst_stride >= 3 * (dstW + 1). */
#define AVB_PIX_FMT_YUV420P 0
#define AVB_PIX_FMT_RGB24   2
#define AVB_PIX_FMT_BGR24   3
typedef struct SwsContextCUDA SwsContextCUDA;
SwsContextCUDA *sws_getContext_cuda(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat, int flags,
                                    void *srcFilter, void *dstFilter, const double *param);
void sws_freeContext_cuda(SwsContextCUDA *ctx);
/* sws_setColorspaceDetails (libswscale/swscale.h:269-271, utils.c:807-835) for packed rgb destinations: inv_table = the four
 * yuv -> rgb coefficients (ff_yuv2rgb_coeffs[], e.g. SWS_CS_ITU709 = { 117504, 138453, 13954, 34903 }), srcRange 1 = full-range
 * (JPEG) yuv, brightness / contrast / saturation in 16.16 fixed point (0, 1 << 16, 1 << 16 = neutral).  Returns 0, or -1 for a yuv
 * destination (like the reference) and for settings that would index outside the reference's 1024-entry colour table.  The
 * full-range source formats yuvj420p = 12, yuvj422p = 13, yuvj444p = 14 are taken over as sources of rgb destinations. */
int sws_setColorspaceDetails_cuda(SwsContextCUDA *ctx, const int inv_table[4], int srcRange, const int table[4], int dstRange,
                                  int brightness, int contrast, int saturation);
int  sws_scale_cuda(SwsContextCUDA *ctx, const uint8_t *const srcSlice[], const int srcStride[], int srcSliceY,
                    int srcSliceH, uint8_t *const dst[], const int dstStride[]);
int  sws_scale_frames_cuda(SwsContextCUDA *ctx, const uint8_t *const src[3], const int srcStride[3],
                           const size_t srcFrameStride[3], uint8_t *const dst[3], const int dstStride[3],
                           const size_t dstFrameStride[3], int nframes, void *stream);
int  sws_is_fused_cuda(SwsContextCUDA *ctx);   /* 1 when the single-kernel same-size path is selected */
/* host-only introspection of the set-up stage (filter banks / colour constants), used to pin it against the
 * reference without a GPU: which = 0 hLum, 1 hChr, 2 vLum, 3 vChr; returns taps per output sample */
int  sws_debug_filter_cuda(int which, int srcW, int srcH, int dstW, int dstH, int dstFormat, int flags,
                           int16_t *filter, int32_t *pos, int cap, int *n_out);
/* the same with a SwsFilter pair (the reference's struct: four pointers to { double *coeff; int length; }), NULL = none */
int  sws_debug_filter2_cuda(int which, int srcW, int srcH, int dstW, int dstH, int dstFormat, int flags, const void *srcFilter,
                            const void *dstFilter, int16_t *filter, int32_t *pos, int cap, int *n_out);
void sws_debug_rgb_constants_cuda(int32_t out[10]);
/* what sws_getContext_cuda would decide, computed on the host only: 1 = taken over (out[0] path: 1 plane copy, 2 unscaled table
 * converter, 3 fused same-size kernel, 4 general scaler, 5 packed-source special converter, 6 planar -> packed 4:2:2 converter,
 * 7 yuv420p -> nv12 interleave; out[1..4] chrSrcW chrSrcH chrDstW chrDstH; out[5] source pre-pass 0 / 1 nv split / 2 reader;
 * out[6] destination bits per sample (planar) or bytes per pixel (packed); out[7] full-range source), 0 = refused */
int  sws_debug_plan_cuda(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat, int flags, int32_t out[8]);
/* what the per-line slots (section 3, ff_sws_init_swscale_cuda) would know about that context: 26 int32 (19 colour constants, flags,
 * planar, destination bits, big endian, packed target, nv12 / nv21, range conversion); returns the count, 0 = refused.  Host only. */
int  sws_debug_slot_view_cuda(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat, int flags, int32_t out[32]);

/* ------------------------------------------------------------------ 3. table hooks --------------- */
/* One more arch behind ff_idctdsp_init()'s dispatch (libavcodec/idctdsp.c:183-188; same shape as
 * ff_idctdsp_init_x86, libavcodec/idctdsp.h:109-110).  AVCodecContext is opaque to this library, so the
 * two fields the hook needs are passed by value (INTEGRATION.md shows the one-line caller).
 * Taken over: idct_algo FF_IDCT_SIMPLE / FF_IDCT_AUTO at 8 bit, and bits_per_raw_sample 10 (ff_simple_idct_*_10 whatever idct_algo says,
 * idctdsp.c:151-155); anything else leaves `c` untouched. */
void ff_idctdsp_init_cuda(IDCTDSPContext *c, int idct_algo, int bits_per_raw_sample, unsigned high_bit_depth);
void ff_blockdsp_init_cuda(BlockDSPContext *c);   /* libavcodec/blockdsp.c:60-74 */
/* libavcodec/fdctdsp.c:27-50 (same shape as ff_fdctdsp_init_x86): dct_algo FF_DCT_AUTO / FF_DCT_INT -> islow,
 * FF_DCT_FASTINT -> ifast; bits_per_raw_sample 10 -> jpeg_fdct_islow_10 / fdct248_islow_10; FF_DCT_FAAN is left to the C path */
void ff_fdctdsp_init_cuda(FDCTDSPContext *c, int dct_algo, int bits_per_raw_sample, unsigned high_bit_depth);
/* libavcodec/me_cmp.c:895-944: every slot ff_me_cmp_init fills except the encoder-state metrics */
void ff_me_cmp_init_cuda(MECmpContext *c);
/* MECmpContext.quant_psnr[0..1] / bit[0..1] / rd[0..1] as slots.  They read a LIVE encoder context through `view`: host pointers to the
 * fields of the MpegEncContext `s` the C functions read (and the two they write: mb_intra, block_last_index[0]), filled once by the glue
 * with the reference's own headers (examples/reference_binding/dsp_init_cuda.c); every call snapshots them, so qscale / mb_intra may change
 * between macroblocks like in the encoder loop.  `s` is only the key the slots find their view by (struct MpegEncContext stays opaque).
 * Returns 0, or -1 (table untouched) for a state the kernels do not cover: a permuting IDCT, denoise_dct (s->dct_error_sum), trellis. */
typedef struct FFMECmpEncView {
    const int *qscale, *y_dc_scale, *h263_aic, *intra_quant_bias, *inter_quant_bias, *ac_esc_length;
    int *mb_intra, *block_last_index;
    int (*const *q_intra_matrix)[64], (*const *q_inter_matrix)[64];     /* &s->q_intra_matrix, &s->q_inter_matrix */
    const uint16_t *intra_matrix, *inter_matrix;
    const uint8_t *scantable;                                            /* s->intra_scantable.scantable */
    uint8_t *const *intra_ac_vlc_length, *const *intra_ac_vlc_last_length, *const *inter_ac_vlc_length, *const *inter_ac_vlc_last_length;
    const uint8_t *const *luma_dc_vlc_length;                            /* &s->intra_ac_vlc_length ... (codec init may set them later) */
    int fdct, dequant;                                                   /* as in FFMECmpEncState (the glue compares function pointers) */
    int idct_perm_none, plain_quantiser;                                 /* s->idsp.perm_type == FF_IDCT_PERM_NONE; fast_dct_quantize == ff_dct_quantize_c && !dct_error_sum */
} FFMECmpEncView;
int ff_me_cmp_enc_init_cuda(MECmpContext *c, struct MpegEncContext *s, const FFMECmpEncView *view);
void ff_me_cmp_enc_uninit_cuda(struct MpegEncContext *s);
/* libavcodec/h264dsp.c:57-143, h264qpel.c:36-89, h264chroma.c:32-55, hpeldsp.c:338-366.  bit_depth 8, and 9 / 10 (the reference's BIT_DEPTH > 8
 * instances: uint8_t * arguments point at uint16 samples, int16_t * at int32 coefficients, strides stay in bytes).  ff_h264dsp_init_cuda fills
 * every entry ff_h264dsp_init fills (mbaff loop filters and startcode_find_candidate included); chroma_format_idc > 1 selects the
 * 4:2:2 entries (idct_add8_422, the 2x4 chroma DC transform, the 16-line h_ chroma filters) exactly like h264dsp.c:81-122 */
void ff_h264dsp_init_cuda(H264DSPContext *c, const int bit_depth, const int chroma_format_idc);
void ff_h264qpel_init_cuda(H264QpelContext *c, int bit_depth);
void ff_h264chroma_init_cuda(H264ChromaContext *c, int bit_depth);
void ff_hpeldsp_init_cuda(HpelDSPContext *c, int flags);
/* libavcodec/h264pred.h:112-123; takes over codec_id AV_CODEC_ID_H264, bit_depth 8 / 9 / 10 (16-bit samples, int32 residual above 8);
 * chroma_format_idc > 1 installs the 8 x 16 chroma functions in pred8x8[] / pred8x8_add[] like h264pred.c:477-563 */
void ff_h264_pred_init_cuda(H264PredContext *h, int codec_id, const int bit_depth, const int chroma_format_idc);
/* libavcodec/pixblockdsp.h:37-43; high_bit_depth != 0 leaves the table untouched */
void ff_pixblockdsp_init_cuda(PixblockDSPContext *c, unsigned high_bit_depth);
/* libavcodec/qpeldsp.h:75-77 */
void ff_qpeldsp_init_cuda(QpelDSPContext *c);
/* libavcodec/fft_template.c:152-159, mdct_template.c:58-66 (same shape as ff_fft_init_x86 / ff_mdct_init_x86): called after
 * ff_fft_init / ff_mdct_init filled the context.  fft_permute stays the reference's; the CUDA fft_calc accepts its revtab
 * order, the MDCT slots use the context's own tcos / tsin tables. */
void ff_fft_init_cuda(FFTContext *s);
void ff_mdct_init_cuda(FFTContext *s);


/* libswscale's per-line slots: what sws_init_swscale() and its arch hooks (ff_sws_init_swscale_x86 / _ppc, libswscale/swscale.c:723-783)
 * install in a SwsContext (the function-pointer fields of libswscale/swscale_internal.h:312-330,478-535), with the reference's
 * signatures (swscale_internal.h:62-110) and HOST pointers.  `struct SwsContext` stays opaque to this library: it is only the key under
 * which ff_sws_init_swscale_cuda() remembers which SwsContextCUDA (destination format, colour constants) a slot call belongs to.
 * Filled like the reference fills them for an 8-bit source: hyScale / hcScale = hScale8To15_c (hScale8To19_c for a 16-bit destination),
 * hyscale_fast / hcscale_fast only with SWS_FAST_BILINEAR (and not for 16-bit destinations), yuv2plane1 / yuv2planeX by destination depth
 * (8, 9, 10, 16; LE / BE), yuv2nv12cX for nv12 / nv21, yuv2packed1 / 2 / X for rgb24 bgr24 argb rgba abgr bgra yuyv422 uyvy422 (only
 * yuv2packedX with SWS_FULL_CHR_H_INT, output.c:1392-1472); entries the reference leaves NULL are set to NULL.  Each call is a batch of
 * one line on the GPU (stage, copy, kernel, copy back, synchronise): plumbing and parity, like the DSP table slots -- whole frames go
 * through sws_scale_cuda / sws_scale_frames_cuda. */
struct SwsContext;
typedef struct SwsLineSlotsCUDA {
    void (*hyScale)(struct SwsContext *c, int16_t *dst, int dstW, const uint8_t *src, const int16_t *filter, const int32_t *filterPos, int filterSize);
    void (*hcScale)(struct SwsContext *c, int16_t *dst, int dstW, const uint8_t *src, const int16_t *filter, const int32_t *filterPos, int filterSize);
    void (*hyscale_fast)(struct SwsContext *c, int16_t *dst, int dstWidth, const uint8_t *src, int srcW, int xInc);
    void (*hcscale_fast)(struct SwsContext *c, int16_t *dst1, int16_t *dst2, int dstWidth, const uint8_t *src1, const uint8_t *src2, int srcW, int xInc);
    void (*yuv2plane1)(const int16_t *src, uint8_t *dest, int dstW, const uint8_t *dither, int offset);
    void (*yuv2planeX)(const int16_t *filter, int filterSize, const int16_t **src, uint8_t *dest, int dstW, const uint8_t *dither, int offset);
    void (*yuv2nv12cX)(struct SwsContext *c, const int16_t *chrFilter, int chrFilterSize, const int16_t **chrUSrc, const int16_t **chrVSrc,
                       uint8_t *dest, int dstW);
    void (*yuv2packed1)(struct SwsContext *c, const int16_t *lumSrc, const int16_t *chrUSrc[2], const int16_t *chrVSrc[2], const int16_t *alpSrc,
                        uint8_t *dest, int dstW, int uvalpha, int y);
    void (*yuv2packed2)(struct SwsContext *c, const int16_t *lumSrc[2], const int16_t *chrUSrc[2], const int16_t *chrVSrc[2],
                        const int16_t *alpSrc[2], uint8_t *dest, int dstW, int yalpha, int uvalpha, int y);
    void (*yuv2packedX)(struct SwsContext *c, const int16_t *lumFilter, const int16_t **lumSrc, int lumFilterSize, const int16_t *chrFilter,
                        const int16_t **chrUSrc, const int16_t **chrVSrc, int chrFilterSize, const int16_t **alpSrc, uint8_t *dest, int dstW, int y);
    /* set when the context converts between full-range (yuvj) and limited-range yuv (swscale.c:748-757): lum / chrRangeFromJpeg_c or
     * lum / chrRangeToJpeg_c on 15-bit lines, in place; NULL otherwise */
    void (*lumConvertRange)(int16_t *dst, int width);
    void (*chrConvertRange)(int16_t *dst1, int16_t *dst2, int width);
} SwsLineSlotsCUDA;
/* 0, or -1 (sticky error) for a NULL argument.  The registration lasts until sws_freeContext_cuda(cuda). */
int ff_sws_init_swscale_cuda(struct SwsContext *c, SwsContextCUDA *cuda, SwsLineSlotsCUDA *slots);

#ifdef __cplusplus
}
#endif
#endif /* AVDSP_B200_H */
