/*
 * oracle/port/orc_mpv.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of MpegEncContext.dct_unquantize_* (libavcodec/mpegvideo.c:51-270) for the simple IDCT (no
 * coefficient permutation): one rule per coefficient -- "is raster position j inside the coded part of the block,
 * and what does a non-zero level become" -- instead of the reference's seven scan-order loops.  Scan tables follow
 * ff_init_scantable (libavcodec/idctdsp.c:28-47) over ff_zigzag_direct / ff_alternate_vertical_scan
 * (libavcodec/mathtables.c, mpegvideodata.c).  Pinned against oracle/_ref in tests/test_oracle_mpv_cpu.py.
 */
#include <stdint.h>
#include <string.h>
#include "../oracle_api.h"

static const uint8_t zigzag[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
static const uint8_t alt_vertical[64] = {
    0, 8, 16, 24, 1, 9, 2, 10, 17, 25, 32, 40, 48, 56, 57, 49, 41, 33, 26, 18, 3, 11, 4, 12, 19, 27, 34, 42, 50, 58, 35, 43,
    51, 59, 20, 28, 5, 13, 6, 14, 21, 29, 36, 44, 52, 60, 37, 45, 53, 61, 22, 30, 7, 15, 23, 31, 38, 46, 54, 62, 39, 47, 55, 63 };

void orc_mpeg_scantables(int alternate_scan, uint8_t *permutated, uint8_t *raster_end)
{
    const uint8_t *scan = alternate_scan ? alt_vertical : zigzag;
    int end = -1;
    for (int i = 0; i < 64; i++) {
        permutated[i] = scan[i];
        if (scan[i] > end) end = scan[i];
        raster_end[i] = end;
    }
}

void orc_mpeg_dequant(int kind, int16_t *block, int n, int qscale, int last_index, int y_dc_scale, int c_dc_scale,
                      const uint16_t *intra_matrix, const uint16_t *inter_matrix, int alternate_scan, int h263_aic, int ac_pred)
{
    uint8_t scan[64], rend[64], rank[64];
    const int intra = kind == 0 || kind == 2 || kind == 3 || kind == 5;
    const int h263 = kind >= 5;
    orc_mpeg_scantables(alternate_scan, scan, rend);
    for (int i = 0; i < 64; i++) rank[scan[i]] = i;
    /* how far the coded part reaches: in scan order for the MPEG quantisers, in raster order for H.263 */
    int limit;
    if (h263) limit = (kind == 5 && ac_pred) ? 63 : (last_index >= 0 ? rend[last_index] : -1);
    else      limit = ((kind == 2 || kind == 3 || kind == 4) && alternate_scan) ? 63 : last_index;
    const int qmul = qscale << 1, qadd = (kind == 5 && h263_aic) ? 0 : (qscale - 1) | 1;
    int sum = -1;
    for (int j = 0; j < 64; j++) {
        int level = block[j];
        if (intra && j == 0) {
            if (!(kind == 5 && h263_aic)) block[0] = level * (n < 4 ? y_dc_scale : c_dc_scale);
            continue;
        }
        if ((h263 ? j : rank[j]) > limit || !level) continue;
        const int neg = level < 0, a = neg ? -level : level;
        int v;
        switch (kind) {
        case 0:          v = ((a * qscale * intra_matrix[j]) >> 3); v = (v - 1) | 1; break;
        case 1:          v = (((a << 1) + 1) * qscale * inter_matrix[j]) >> 4; v = (v - 1) | 1; break;
        case 2: case 3:  v = (a * qscale * intra_matrix[j]) >> 3; break;
        case 4:          v = (((a << 1) + 1) * qscale * inter_matrix[j]) >> 4; break;
        default:         v = a * qmul + qadd; break;
        }
        v = neg ? -v : v;
        block[j] = v;
        sum += v;
    }
    if (kind == 3 || kind == 4) block[63] ^= sum & 1;          /* MPEG-2 mismatch control */
}
