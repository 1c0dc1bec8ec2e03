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
/* libavcodec/mpegvideo_enc.c (compiled for ff_dct_quantize_c / ff_convert_matrix, refapi_enc.c): rate control, motion estimation, packets */
STUB(av_cpb_properties_alloc) STUB(av_init_packet) STUB(av_packet_add_side_data) STUB(av_packet_new_side_data)
STUB(av_packet_shrink_side_data) STUB(av_packet_unref) STUB(avcodec_alloc_context3) STUB(avcodec_find_encoder) STUB(avcodec_free_context)
STUB(avcodec_open2) STUB(avcodec_receive_packet) STUB(avcodec_send_frame) STUB(avpriv_align_put_bits) STUB(avpriv_copy_bits)
STUB(ff_add_cpb_side_data) STUB(ff_alloc_packet) STUB(ff_estimate_b_frame_motion) STUB(ff_estimate_p_frame_motion) STUB(ff_fix_long_mvs)
STUB(ff_fix_long_p_mvs) STUB(ff_get_2pass_fcode) STUB(ff_get_best_fcode) STUB(ff_init_me) STUB(ff_pre_estimate_p_frame_motion)
STUB(ff_rate_control_init) STUB(ff_rate_control_uninit) STUB(ff_rate_estimate_qscale) STUB(ff_vbv_update)
unsigned int avpriv_toupper4(unsigned int x) { return x; }      /* libavcodec/utils.c; only touches codec_tag */

