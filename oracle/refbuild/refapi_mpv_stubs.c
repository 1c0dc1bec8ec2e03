/*
 * oracle/refbuild/refapi_mpv_stubs.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 * Link-time stand-ins for the decoder entry points libavcodec/mpegvideo.c references but the inverse-quantiser path
 * (refapi_mpv.c) never calls.  Kept free of reference headers so the untyped definitions do not clash.
 */
#include <stdio.h>
#include <stdlib.h>

#define STUB(name) void name(void) { fprintf(stderr, "oracle/_ref: %s is outside the DSP path\n", #name); abort(); }
STUB(ff_alloc_picture) STUB(ff_draw_horiz_band) STUB(ff_find_unused_picture) STUB(ff_free_picture_tables)
STUB(ff_mpeg_er_init) STUB(ff_mpeg_framesize_alloc) STUB(ff_mpeg_ref_picture) STUB(ff_mpeg_unref_picture)
STUB(ff_mpv_motion) STUB(ff_thread_await_progress) STUB(ff_thread_report_progress) STUB(ff_update_picture_tables)
unsigned int avpriv_toupper4(unsigned int x) { return x; }      /* libavcodec/utils.c; only touches codec_tag */

