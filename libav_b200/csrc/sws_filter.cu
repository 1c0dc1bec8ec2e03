// libav_b200/csrc/sws_filter.cu -- host code (no kernels): see sws_filter.h.
// Every constant produced here is checked against the reference's SwsContext in tests/test_sws_cpu.py
// through sws_debug_filter_cuda()/sws_debug_rgb_constants_cuda() (no GPU needed for those).
#include "sws_filter.h"
#include <math.h>
#include <stdlib.h>

namespace avb {

namespace {

const int64_t kOne54 = 1LL << 54;           // working precision of the design stage (utils.c:262)

inline int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }

double spline_weight(double a, double b, double c, double d, double dist)      // utils.c:236-247
{
    while (dist > 1.0) {
        double nb = b + 2.0 * c + 3.0 * d, nc = c + 3.0 * d, nd = -b - 3.0 * c - 6.0 * d;
        a = 0.0; b = nb; c = nc; d = nd;
        dist -= 1.0;
    }
    return ((d * dist + c) * dist + b) * dist + a;
}

// weight of one tap at (scaled) distance d for the windowed kernels, utils.c:345-437
int64_t kernel_weight(int flags, const double param[2], int64_t d, int x_inc)
{
    const double fd = d * (1.0 / (1 << 30));
    if (flags & SWS_BICUBIC) {
        int64_t B = (int64_t)((param[0] != SWS_PARAM_DEFAULT ? param[0] : 0) * (1 << 24));
        int64_t C = (int64_t)((param[1] != SWS_PARAM_DEFAULT ? param[1] : 0.6) * (1 << 24));
        int64_t w;
        if (d >= 1LL << 31) {
            w = 0;
        } else {
            int64_t dd = (d * d) >> 30, ddd = (dd * d) >> 30;
            if (d < 1LL << 30)
                w = (12 * (1 << 24) - 9 * B - 6 * C) * ddd + (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                    (6 * (1 << 24) - 2 * B) * (1LL << 30);
            else
                w = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd + (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1LL << 30);
        }
        return w * (kOne54 >> (30 + 24));
    }
    if (flags & SWS_X) {
        double A = param[0] != SWS_PARAM_DEFAULT ? param[0] : 1.0;
        double c = fd < 1.0 ? cos(fd * M_PI) : -1.0;
        c = c < 0.0 ? -pow(-c, A) : pow(c, A);
        return (int64_t)((c * 0.5 + 0.5) * kOne54);
    }
    if (flags & SWS_AREA) {
        int64_t d2 = d - (1 << 29), w;
        if (d2 * x_inc < -(1LL << (29 + 16)))      w = (int64_t)(1.0 * (1LL << (30 + 16)));
        else if (d2 * x_inc < (1LL << (29 + 16))) w = -d2 * x_inc + (1LL << (29 + 16));
        else                                       w = 0;
        return w * (kOne54 >> (30 + 16));
    }
    if (flags & SWS_GAUSS) {
        double p = param[0] != SWS_PARAM_DEFAULT ? param[0] : 3.0;
        return (int64_t)(pow(2.0, -p * fd * fd) * kOne54);
    }
    if (flags & SWS_SINC)
        return (int64_t)((d ? sin(fd * M_PI) / (fd * M_PI) : 1.0) * kOne54);
    if (flags & SWS_LANCZOS) {
        double p = param[0] != SWS_PARAM_DEFAULT ? param[0] : 3.0;
        int64_t w = (int64_t)((d ? sin(fd * M_PI) * sin(fd * M_PI / p) / (fd * fd * M_PI * M_PI / p) : 1.0) * kOne54);
        return fd > p ? 0 : w;
    }
    if (flags & SWS_BILINEAR) {
        int64_t w = (1 << 30) - d;
        if (w < 0) w = 0;
        return w * (kOne54 >> 30);
    }
    if (flags & SWS_SPLINE) {
        const double p = -2.196152422706632;
        return (int64_t)(spline_weight(1.0, 0.0, p, -p - 1.0, fd) * kOne54);
    }
    return 0;
}

}  // namespace

int design_filter(FilterBank &fb, int x_inc, int src_len, int dst_len, int one, int flags,
                  const double param[2], bool horizontal, const char **err, const SwsVec *srcv, const SwsVec *dstv)
{
    const int n = dst_len;
    int taps;
    std::vector<int64_t> w;          // n * taps working weights
    std::vector<int32_t> pos(n);

    if (abs(x_inc - 0x10000) < 10) {                         // same size: identity
        taps = 1;
        w.assign(n, kOne54);
        for (int i = 0; i < n; i++) pos[i] = i;
    } else if (flags & SWS_POINT) {
        taps = 1;
        w.assign(n, kOne54);
        int x = x_inc / 2 - 0x8000;
        for (int i = 0; i < n; i++, x += x_inc) pos[i] = (x + (1 << 15)) >> 16;
    } else if ((x_inc <= (1 << 16) && (flags & SWS_AREA)) || (flags & SWS_FAST_BILINEAR)) {
        taps = 2;
        w.resize((size_t)n * 2);
        int x = x_inc / 2 - 0x8000;
        for (int i = 0; i < n; i++, x += x_inc) {
            int xx = x >> 16;
            pos[i] = xx;
            for (int j = 0; j < 2; j++, xx++) {
                int64_t c = kOne54 - (int64_t)abs((int)(((unsigned)xx << 16) - (unsigned)x)) * (kOne54 >> 16);
                w[(size_t)i * 2 + j] = c < 0 ? 0 : c;
            }
        }
    } else {
        int support;
        if (flags & SWS_BICUBIC)       support = 4;
        else if (flags & SWS_X)        support = 8;
        else if (flags & SWS_AREA)     support = 1;
        else if (flags & SWS_GAUSS)    support = 8;
        else if (flags & SWS_LANCZOS)  support = param[0] != SWS_PARAM_DEFAULT ? (int)ceil(2 * param[0]) : 6;
        else if (flags & SWS_SINC)     support = 20;
        else if (flags & SWS_SPLINE)   support = 20;
        else if (flags & SWS_BILINEAR) support = 2;
        else { *err = "no scaler algorithm selected"; return -1; }

        taps = x_inc <= (1 << 16) ? 1 + support : 1 + (support * src_len + dst_len - 1) / dst_len;
        if (taps > src_len - 2) taps = src_len - 2;
        if (taps < 1) taps = 1;
        w.resize((size_t)n * taps);
        int64_t x = x_inc - 0x10000;
        for (int i = 0; i < n; i++, x += 2 * (int64_t)x_inc) {
            int xx = (int)((x - ((int64_t)(taps - 2) << 16)) / (1 << 17));     // C division: toward zero
            pos[i] = xx;
            for (int j = 0; j < taps; j++, xx++) {
                int64_t d = iabs64(((int64_t)xx << 17) - x) << 13;
                if (x_inc > (1 << 16)) d = d * dst_len / src_len;
                w[(size_t)i * taps + j] = kernel_weight(flags, param, d, x_inc);
            }
        }
    }

    // the caller's SwsFilter vectors (utils.c:444-474): the source-side vector is convolved into every row -- int64 weights times double
    // coefficients, accumulated through the reference's own int64 += double conversions -- the destination-side vector only widens the
    // rows ("FIXME dstFilter" in the reference); the positions move by the difference of the half widths
    if ((srcv && srcv->length > 0) || (dstv && dstv->length > 0)) {
        int taps2 = taps;
        if (srcv && srcv->length > 0) taps2 += srcv->length - 1;
        if (dstv && dstv->length > 0) taps2 += dstv->length - 1;
        std::vector<int64_t> w2((size_t)n * taps2, 0);
        for (int i = 0; i < n; i++) {
            if (srcv && srcv->length > 0) {
                for (int k = 0; k < srcv->length; k++)
                    for (int j = 0; j < taps; j++) {
                        int64_t &d = w2[(size_t)i * taps2 + k + j];
                        d = (int64_t)((double)d + srcv->coeff[k] * (double)w[(size_t)i * taps + j]);
                    }
            } else {
                for (int j = 0; j < taps; j++) w2[(size_t)i * taps2 + j] = w[(size_t)i * taps + j];
            }
            pos[i] += (taps - 1) / 2 - (taps2 - 1) / 2;
        }
        w.swap(w2);
        taps = taps2;
    }

    // trim near-zero taps: shift rows left while the leading mass is below the cut-off, then find the
    // longest remaining row (utils.c:476-520); rows are processed last to first because the monotonicity
    // guard looks at the already trimmed successor.
    const double cutoff = 0.002 * (double)kOne54;            // SWS_MAX_REDUCE_CUTOFF * fone
    int min_taps = 0;
    for (int i = n - 1; i >= 0; i--) {
        int64_t *row = &w[(size_t)i * taps];
        int64_t acc = 0;
        for (int j = 0; j < taps; j++) {
            acc += iabs64(row[0]);
            if ((double)acc > cutoff) break;
            if (i < n - 1 && pos[i] >= pos[i + 1]) break;
            for (int k = 1; k < taps; k++) row[k - 1] = row[k];
            row[taps - 1] = 0;
            pos[i]++;
        }
        acc = 0;
        int keep = taps;
        for (int j = taps - 1; j > 0; j--) {
            acc += iabs64(row[j]);
            if ((double)acc > cutoff) break;
            keep--;
        }
        if (keep > min_taps) min_taps = keep;
    }
    if (min_taps < 1) { *err = "degenerate filter"; return -1; }
    const int out_taps = min_taps;                            // filterAlign == 1 on this back-end
    // admission rule of utils.c:541-543: MAX_FILTER_SIZE * 16 / APCK_SIZE.  APCK_SIZE is 16 on every target but
    // x86-64 (24, which lowers the limit to 170 under SWS_ACCURATE_RND); this back-end follows the portable C
    // configuration, the one the parity oracle is built with.
    if (out_taps >= 256) { *err = "filter too large"; return -1; }

    std::vector<int64_t> f((size_t)n * out_taps);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < out_taps; j++) f[(size_t)i * out_taps + j] = j < taps ? w[(size_t)i * taps + j] : 0;

    if (horizontal) {                                         // fold taps that fall outside the line, :565-590
        for (int i = 0; i < n; i++) {
            int64_t *row = &f[(size_t)i * out_taps];
            if (pos[i] < 0) {
                for (int j = 1; j < out_taps; j++) {
                    int left = j + pos[i] > 0 ? j + pos[i] : 0;
                    row[left] += row[j];
                    row[j] = 0;
                }
                pos[i] = 0;
            }
            if (pos[i] + out_taps > src_len) {
                int shift = pos[i] + out_taps - src_len;
                for (int j = out_taps - 2; j >= 0; j--) {
                    int right = j + shift < out_taps - 1 ? j + shift : out_taps - 1;
                    row[right] += row[j];
                    row[j] = 0;
                }
                pos[i] = src_len - out_taps;
            }
        }
    }

    fb.size = out_taps;
    fb.n = n;
    fb.pos = pos;
    fb.coef.assign((size_t)n * out_taps, 0);
    for (int i = 0; i < n; i++) {                             // normalise with error diffusion, :597-613
        const int64_t *row = &f[(size_t)i * out_taps];
        int64_t sum = 0, carry = 0;
        for (int j = 0; j < out_taps; j++) sum += row[j];
        sum = (sum + one / 2) / one;
        if (sum == 0) { *err = "zero filter row"; return -1; }
        for (int j = 0; j < out_taps; j++) {
            int64_t v = row[j] + carry;
            int q = (int)((v > 0 ? v + (sum >> 1) : v - (sum >> 1)) / sum);
            fb.coef[(size_t)i * out_taps + j] = (int16_t)q;
            carry = v - q * sum;
        }
    }
    return 0;
}

int derive_geometry(SwsGeometry &g, int srcW, int srcH, int dstW, int dstH, bool dst_is_rgb, int flags, const char **err,
                    int chrSrcHSub, int chrSrcVSub, int chrDstHSub, int chrDstVSub)
{
    int algos = flags & (SWS_POINT | SWS_AREA | SWS_BILINEAR | SWS_FAST_BILINEAR | SWS_BICUBIC | SWS_X | SWS_GAUSS |
                         SWS_LANCZOS | SWS_SINC | SWS_SPLINE | SWS_BICUBLIN);
    if (!algos) {                                             // utils.c:937-944
        if (dstW < srcW && dstH < srcH)      flags |= SWS_GAUSS;
        else if (dstW > srcW && dstH > srcH) flags |= SWS_SINC;
        else                                 flags |= SWS_LANCZOS;
    } else if (algos & (algos - 1)) { *err = "exactly one scaler algorithm must be chosen"; return -1; }
    if (srcW < 4 || srcH < 1 || dstW < 8 || dstH < 1) { *err = "invalid scaling dimension"; return -1; }
    g.srcW = srcW; g.srcH = srcH; g.dstW = dstW; g.dstH = dstH; g.flags = flags;
    g.lumXInc = (int)((((int64_t)srcW << 16) + (dstW >> 1)) / dstW);
    g.lumYInc = (int)((((int64_t)srcH << 16) + (dstH >> 1)) / dstH);
    g.chrSrcHSub = chrSrcHSub; g.chrSrcVSub = chrSrcVSub;     // getSubSampleFactors(), utils.c:983
    if (dst_is_rgb) { g.chrDstHSub = (flags & SWS_FULL_CHR_H_INT) ? 0 : 1; g.chrDstVSub = 0; }   // :1013-1014
    else            { g.chrDstHSub = chrDstHSub; g.chrDstVSub = chrDstVSub; }
    g.chrSrcVSub += (flags & 0x30000) >> 16;                  // SWS_SRC_V_CHR_DROP
    g.chrSrcW = -((-srcW) >> g.chrSrcHSub);
    g.chrSrcH = -((-srcH) >> g.chrSrcVSub);
    g.chrDstW = -((-dstW) >> g.chrDstHSub);
    g.chrDstH = -((-dstH) >> g.chrDstVSub);
    g.chrXInc = (int)((((int64_t)g.chrSrcW << 16) + (g.chrDstW >> 1)) / g.chrDstW);
    g.chrYInc = (int)((((int64_t)g.chrSrcH << 16) + (g.chrDstH >> 1)) / g.chrDstH);
    return 0;
}

void rgb_constants(RgbConstants &k, const int inv_table[4], int full_range, int brightness, int contrast, int saturation)
{
    const int yoffs = full_range ? 384 : 326;
    int64_t crv = inv_table[0], cbu = inv_table[1], cgu = -inv_table[2], cgv = -inv_table[3];
    int64_t cy = 1 << 16, oy = 0;
    if (!full_range) { cy = (cy * 255) / 219; oy = 16 << 16; }
    else { crv = (crv * 224) / 255; cbu = (cbu * 224) / 255; cgu = (cgu * 224) / 255; cgv = (cgv * 224) / 255; }
    cy  = (cy * contrast) >> 16;
    crv = (crv * contrast * saturation) >> 32;
    cbu = (cbu * contrast * saturation) >> 32;
    cgu = (cgu * contrast * saturation) >> 32;
    cgv = (cgv * contrast * saturation) >> 32;
    oy -= 256 * (int64_t)brightness;
    {   // roundToInt16(), yuv2rgb.c:659-669, then the (int16_t) casts of :735-740
        auto r16 = [](int64_t f) { int r = (int)((f + (1 << 15)) >> 16); return (int)(int16_t)(r < -0x7FFF ? 0x8000 : r > 0x7FFF ? 0x7FFF : r); };
        k.fy_coeff = r16(cy << 13); k.fy_offset = r16(oy << 9);
        k.fv2r = r16(crv << 13); k.fv2g = r16(cgv << 13); k.fu2g = r16(cgu << 13); k.fu2b = r16(cbu << 13);
    }
    crv = ((crv << 16) + 0x8000) / cy;
    cbu = ((cbu << 16) + 0x8000) / cy;
    cgu = ((cgu << 16) + 0x8000) / cy;
    cgv = ((cgv << 16) + 0x8000) / cy;
    k.cy = (int)cy;
    k.k1 = (int)(-(384LL << 16) - oy + 0x8000);
    k.crv = (int)crv; k.cgu = (int)cgu; k.cgv = (int)cgv; k.cbu = (int)cbu;
    k.ar  = yoffs - (int)(crv >> 9);
    k.agu = yoffs - (int)(cgu >> 9);
    k.agv = -(int)(cgv >> 9);
    k.ab  = yoffs - (int)(cbu >> 9);
    k.kr = k.cy * k.ar + k.k1;
    k.kg = k.cy * (k.agu + k.agv) + k.k1;
    k.kb = k.cy * k.ab + k.k1;
}

}  // namespace avb
