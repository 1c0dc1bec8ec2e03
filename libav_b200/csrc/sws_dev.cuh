// libav_b200/csrc/sws_dev.cuh -- per-sample device arithmetic of libswscale's output stage, shared by the frame kernels
// (swscale.cu) and the per-line SwsContext slots (sws_slots.cu).
#pragma once
#include "common.cuh"
#include "sws_filter.h"

namespace avb {

// yuv2plane1_8_c / yuv2planeX_8_c (output.c:242-265, dither 64 everywhere) and yuv2plane1_10_c / yuv2planeX_10_c (:183-213) are one
// recipe in the output depth: plane1 (v + (1 << (14 - bits))) >> (15 - bits), planeX ((1 << (26 - bits)) + sum) >> (27 - bits),
// clipped to `bits` bits
__device__ __forceinline__ int plane_clip(int v, int bits) { return min(max(v, 0), (1 << bits) - 1); }
__device__ __forceinline__ int swap16_if(int v, int be) { return be ? ((v >> 8) | (v << 8)) & 0xFFFF : v; }
__device__ __forceinline__ void plane_store(uint8_t *row, int x, int v, int bits, int be)
{
    if (bits == 8) row[x] = (uint8_t)v; else reinterpret_cast<uint16_t *>(row)[x] = (uint16_t)swap16_if(v, be);
}

struct ChromaTerms { int tr, tg, tb; };

// per (U,V): the additive term of each channel, so a pixel costs one IMAD + shift per channel
__device__ __forceinline__ ChromaTerms chroma_terms(int U, int V, const RgbConstants &k)
{
    int ro = k.ar + ((V * k.crv) >> 16);
    int go = k.agu + ((U * k.cgu) >> 16) + k.agv + ((V * k.cgv) >> 16);
    int bo = k.ab + ((U * k.cbu) >> 16);
    ChromaTerms t;
    t.tr = k.cy * ro + k.k1;
    t.tg = k.cy * go + k.k1;
    t.tb = k.cy * bo + k.k1;
    return t;
}

// the "clip only when bit 8 is set somewhere" rule of yuv2rgb_X_c_template (output.c:966-971)
__device__ __forceinline__ void clip_if_flagged(int &y1, int &y2, int &u, int &v)
{
    if ((y1 | y2 | u | v) & 0x100) { y1 = clip_u8(y1); y2 = clip_u8(y2); u = clip_u8(u); v = clip_u8(v); }
}

// yuv2rgb24_full_X_c's pixel (output.c:1193-1225): 30-bit fixed-point matrix on Y, U - 128, V - 128 (all << 9 here)
__device__ __forceinline__ void full_pixel(int Y, int U, int V, const RgbConstants &k, int bgr, uint8_t *d)
{
    Y = (Y - k.fy_offset) * k.fy_coeff + (1 << 21);
    int R = Y + V * k.fv2r, G = Y + V * k.fv2g + U * k.fu2g, B = Y + U * k.fu2b;
    if ((R | G | B) & 0xC0000000) { R = min(max(R, 0), 0x3FFFFFFF); G = min(max(G, 0), 0x3FFFFFFF); B = min(max(B, 0), 0x3FFFFFFF); }
    d[bgr ? 2 : 0] = (uint8_t)(R >> 22); d[1] = (uint8_t)(G >> 22); d[bgr ? 0 : 2] = (uint8_t)(B >> 22);
}

}  // namespace avb
