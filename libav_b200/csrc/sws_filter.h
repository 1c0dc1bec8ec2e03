// libav_b200/csrc/sws_filter.h -- host-side scaler set-up for the CUDA swscale path: integer filter banks
// and yuv->rgb constants.  This is the part of libswscale that stays on the host (once per context); it
// has to reproduce the reference's numbers exactly because they are the kernels' constant inputs:
//   filter design      libswscale/utils.c:249-632 (initFilter)
//   geometry           libswscale/utils.c:887-1198 (sws_init_context)
//   colour constants   libswscale/yuv2rgb.c:671-863 (ff_yuv2rgb_c_init_tables, 24 bpp case)
#pragma once
#include <stdint.h>
#include <vector>

namespace avb {

enum {   // libswscale/swscale.h:57-83
    SWS_FAST_BILINEAR = 1, SWS_BILINEAR = 2, SWS_BICUBIC = 4, SWS_X = 8, SWS_POINT = 0x10, SWS_AREA = 0x20,
    SWS_BICUBLIN = 0x40, SWS_GAUSS = 0x80, SWS_SINC = 0x100, SWS_LANCZOS = 0x200, SWS_SPLINE = 0x400,
    SWS_FULL_CHR_H_INT = 0x2000, SWS_FULL_CHR_H_INP = 0x4000, SWS_ACCURATE_RND = 0x40000, SWS_BITEXACT = 0x80000,
};
constexpr double SWS_PARAM_DEFAULT = 123456;

struct FilterBank {
    int size = 0;                 // taps per output sample
    int n = 0;                    // output samples
    std::vector<int16_t> coef;    // n * size
    std::vector<int32_t> pos;     // n
};

// returns 0, or -1 with *err set
// SwsVector (libswscale/swscale.h:106-109): a caller-made filter that initFilter() convolves into every row of a bank (utils.c:444-474)
struct SwsVec { const double *coeff; int length; };
int design_filter(FilterBank &fb, int x_inc, int src_len, int dst_len, int one, int flags,
                  const double param[2], bool horizontal, const char **err, const SwsVec *srcv = nullptr, const SwsVec *dstv = nullptr);

struct SwsGeometry {
    int srcW, srcH, dstW, dstH;
    int chrSrcW, chrSrcH, chrDstW, chrDstH;
    int chrSrcHSub, chrSrcVSub, chrDstHSub, chrDstVSub;
    int lumXInc, lumYInc, chrXInc, chrYInc;
    int flags;
};
// planar 8-bit yuv source; dst_is_rgb selects packed 24-bit RGB (chroma shared by pixel pairs) or planar yuv of the given sub-sampling
int derive_geometry(SwsGeometry &g, int srcW, int srcH, int dstW, int dstH, bool dst_is_rgb, int flags, const char **err,
                    int chrSrcHSub = 1, int chrSrcVSub = 1, int chrDstHSub = 1, int chrDstVSub = 1);     // source chroma sub-sampling (log2), 4:2:0 by default

struct RgbConstants {      // r = clip_u8((cy * (Y + ar + ((V * crv) >> 16)) + k1) >> 16), etc.
    int cy, k1;
    int crv, cgu, cgv, cbu;
    int ar, agu, agv, ab;  // yoffs - (crv >> 9), yoffs - (cgu >> 9), -(cgv >> 9), yoffs - (cbu >> 9)
    int kr, kg, kb;        // cy * ar + k1, cy * (agu + agv) + k1, cy * ab + k1: the additive terms with the offsets folded in
    // SWS_FULL_CHR_H_INT output stage (output.c:1165-1240): c->yuv2rgb_{y_coeff,y_offset,v2r,v2g,u2g,u2b}_coeff, yuv2rgb.c:735-740
    int fy_coeff, fy_offset, fv2r, fv2g, fu2g, fu2b;
};
void rgb_constants(RgbConstants &k, const int inv_table[4], int full_range, int brightness, int contrast, int saturation);


// what the per-line SwsContext slots (sws_slots.cu) need to know about a context made by sws_getContext_cuda (swscale.cu)
struct SwsSlotView {
    RgbConstants k;
    int flags;
    int planar;        // planar / semi-planar yuv destination
    int dstBits, dstBE;
    int target;        // packed destinations: 0 rgb24, 1 bgr24, 2 yuyv422, 3 uyvy422, 4 argb, 5 rgba, 6 abgr, 7 bgra; -1 for planar
    int dstNV;         // 1 nv12, 2 nv21
    int rangeConv;     // 1 full -> limited, 2 limited -> full range on the hscaled lines (yuv destinations), else 0
    int srcBits;       // 8, or 9 / 10 / 16 for the planar high-bit-depth sources (their line functions are hScale16To15_c / ...To19_c)
};
// lum / chrRange{From,To}Jpeg_c (swscale.c:166-197) in place on `rows` lines of `w` 15-bit samples, `stridePx` samples apart (device memory):
// kind 0 lumFromJpeg, 1 chrFromJpeg, 2 lumToJpeg, 3 chrToJpeg.  Defined in sws_slots.cu; 0 / -1
int sws_launch_range(int16_t *plane, int stridePx, int w, int rows, int kind, cudaStream_t stream);
bool sws_slot_view(const void *ctx, SwsSlotView &v);      // false for NULL
void sws_slots_forget(const void *ctx);                   // sws_freeContext_cuda: drop the registrations of this context

}  // namespace avb
