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
 * returns 0 / -1 and records the first failure in a sticky string (avb200_last_error()); slot functions
 * record it too.  There is NO CPU fallback: without a B200 every call fails loudly.
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

/* ---- libswscale boundary (libswscale/swscale.h:159-207) ------------------------------------------------
 * Same argument lists as sws_getContext / sws_scale / sws_freeContext; pixel formats are the reference's
 * AVPixelFormat values (AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_RGB24 = 2, AV_PIX_FMT_BGR24 = 3), flags the
 * reference's SWS_* bits.  Taken over: yuv420p -> rgb24 / bgr24 / yuv420p, any size, every scaler
 * algorithm of initFilter (libswscale/utils.c:249-632), results identical to the reference's C path under
 * SWS_ACCURATE_RND | SWS_BITEXACT.  Anything else returns NULL with an error (no fallback).
 *   sws_scale_cuda         HOST pointers, whole frames (srcSliceY = 0, srcSliceH = srcH); uploads, runs,
 *                          downloads, synchronises; returns output lines like sws_scale(), 0 on bad arguments.
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
void sws_debug_rgb_constants_cuda(int32_t out[10]);

/* ------------------------------------------------------------------ 3. table hooks --------------- */
/* One more arch behind ff_idctdsp_init()'s dispatch (libavcodec/idctdsp.c:183-188; same shape as
 * ff_idctdsp_init_x86, libavcodec/idctdsp.h:109-110).  AVCodecContext is opaque to this library, so the
 * two fields the hook needs are passed by value (INTEGRATION.md shows the one-line caller).
 * Only idct_algo FF_IDCT_SIMPLE/FF_IDCT_AUTO at 8 bit is taken over; anything else leaves `c` untouched. */
void ff_idctdsp_init_cuda(IDCTDSPContext *c, int idct_algo, int bits_per_raw_sample, unsigned high_bit_depth);
void ff_blockdsp_init_cuda(BlockDSPContext *c);   /* libavcodec/blockdsp.c:60-74 */

#ifdef __cplusplus
}
#endif
#endif /* AVDSP_B200_H */
