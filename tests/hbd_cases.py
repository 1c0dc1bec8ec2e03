"""Cases for the 9 / 10-bit H.264 DSP entries (uint16 samples, int32 coefficients), shared by the oracle pinning, the host simulation
and the GPU tests.  A `callee` maps the oracle's entry names (oracle/oracle_api.h h264_hbd_*) to callables with the oracle's argument
order; tests/test_oracle_h264_hbd_cpu.py passes the two oracles, the slot tests an adapter over the product's tables.  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from h264_util import block_offsets, block_offsets_422, scan8


def vp(a, off=0):
    return C.c_void_p(a.ctypes.data + int(off))


def residual(rng, which, bits):
    """48 x 16 int32 coefficients + nnz cache for one macroblock: per block nothing / DC only / full, as the dispatchers distinguish them"""
    coeffs = np.zeros((48, 16), np.int32)
    nnzc = np.zeros(120, np.uint8)
    amp = 600 << (bits - 8)
    if which == 2:
        for i in range(0, 16, 4):
            kind = int(rng.integers(0, 3))
            if kind == 1:
                coeffs[i, 0] = rng.integers(-amp, amp); nnzc[scan8(i)] = 1
            elif kind == 2:
                coeffs[i:i + 4] = rng.integers(-amp // 4, amp // 4, size=(4, 16)); nnzc[scan8(i)] = rng.integers(2, 16)
        return coeffs, nnzc
    blocks = range(16) if which < 2 else [16 + 16 * p + k for p in (0, 1) for k in range(8 if which == 4 else 4)]
    for i in blocks:
        e = i + 4 if (which == 4 and (i & 15) >= 4) else i
        kind = int(rng.integers(0, 4))
        if kind == 1:                           # DC only; add16 takes the dc path for nnz == 1, the intra / chroma dispatchers for nnz == 0
            coeffs[i, 0] = rng.integers(-amp, amp)
            nnzc[scan8(e)] = 1 if which == 0 else 0
        elif kind == 2:
            coeffs[i] = rng.integers(-amp // 4, amp // 4, size=16); nnzc[scan8(e)] = rng.integers(1, 16)
        elif kind == 3 and which == 0:
            nnzc[scan8(e)] = 1                  # nnz == 1 with a zero DC: the full transform of a zero block
    return coeffs, nnzc


def compare(a, b, bits, seed=0):
    """every entry through callee a and callee b on identical inputs; outputs AND clobbered inputs must agree"""
    rng = np.random.default_rng(seed + bits)
    top = (1 << bits) - 1
    n = 0
    pixels = lambda shape: rng.integers(0, top + 1, size=shape).astype(np.uint16)
    # single-block transforms
    for which in range(4):
        for it in range(12):
            nco = 64 if which in (1, 3) else 16
            blk = rng.integers(-(2000 << (bits - 8)), 2000 << (bits - 8), size=nco).astype(np.int32)
            if it == 0:
                blk[:] = rng.integers(-(1 << 20), 1 << 20, size=nco)
            pix = pixels((8, 32))
            res = []
            for c in (a, b):
                p, k = pix.copy(), blk.copy()
                c.h264_hbd_idct(bits, which, vp(p, 16), vp(k), 64)
                res.append((p, k))
            assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]), ("idct", bits, which, it); n += 1
    # per-macroblock dispatchers (which 4 = add8_422)
    for which in range(5):
        bo = 2 * (block_offsets_422(24) if which == 4 else block_offsets(48, 24))        # byte offsets of 16-bit samples
        for it in range(8):
            coeffs, nnzc = residual(rng, which, bits)
            yp, cb, cr = pixels((16, 48)), pixels((16, 24)), pixels((16, 24))
            res = []
            for c in (a, b):
                y2, b2, r2, k = yp.copy(), cb.copy(), cr.copy(), coeffs.copy()
                d2 = (C.c_void_p * 2)(b2.ctypes.data + 16, r2.ctypes.data + 16)
                c.h264_hbd_idct_mb(bits, which, vp(y2, 32), d2, vp(bo), vp(k), 96 if which < 3 else 48, vp(nnzc))
                res.append((y2, b2, r2, k))
            for u, v in zip(*res):
                assert np.array_equal(u, v), ("idct_mb", bits, which, it)
            n += 1
    # DC transforms
    for kind in range(3):
        for q in [777, 1, 4000] + [int(v) for v in rng.integers(1, 4000, size=10)]:
            inp = rng.integers(-3000 << (bits - 8), 3000 << (bits - 8), size=16).astype(np.int32)
            blk = rng.integers(-3000 << (bits - 8), 3000 << (bits - 8), size=256).astype(np.int32)
            res = []
            for c in (a, b):
                o, i2 = blk.copy(), inp.copy()
                c.h264_hbd_dc_dequant(bits, kind, vp(o), vp(i2), q)
                res.append((o, i2))
            assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and not np.array_equal(res[0][0], blk), ("dc", bits, kind, q); n += 1
    # bypass add
    for w8 in (0, 1):
        nn = 8 if w8 else 4
        blk = rng.integers(-300 << (bits - 8), 300 << (bits - 8), size=nn * nn).astype(np.int32)
        pix = pixels((8, 16))
        res = []
        for c in (a, b):
            p, k = pix.copy(), blk.copy()
            c.h264_hbd_add_pixels_clear(bits, w8, vp(p), vp(k), 32)
            res.append((p, k))
        assert np.array_equal(res[0][0], res[1][0]) and not res[0][1].any() and not res[1][1].any(); n += 1
    # weighted prediction
    for widx in range(4):
        for it in range(6):
            ld, w, off, ws = int(rng.integers(0, 8)), int(rng.integers(-128, 128)), int(rng.integers(-128, 128)), int(rng.integers(-128, 128))
            hgt = int(rng.choice([2, 4, 8, 16]))
            pix, src = pixels((16, 32)), pixels((16, 32))
            res = []
            for c in (a, b):
                p, q = pix.copy(), pix.copy()
                c.h264_hbd_weight(bits, widx, vp(p), 64, hgt, ld, w, off)
                c.h264_hbd_biweight(bits, widx, vp(q), vp(src), 64, hgt, ld, w, ws, off)
                res.append((p, q))
            assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]), ("weight", bits, widx, it); n += 1
    # loop filters
    for which in range(16):
        for it in range(10):
            base = rng.integers(0, top + 1)
            pix = np.clip(base + rng.integers(-10 << (bits - 8), (10 << (bits - 8)) + 1, size=(24, 32)), 0, top).astype(np.uint16)
            alpha, beta = int(rng.integers(1, 200)), int(rng.integers(1, 19))
            tc0 = rng.integers(-1, 12, size=4).astype(np.int8)
            res = []
            for c in (a, b):
                p = pix.copy()
                c.h264_hbd_loop_filter(bits, which, vp(p, 2 * (4 * 32 + 8)), 64, alpha, beta, vp(tc0))
                res.append(p)
            assert np.array_equal(res[0], res[1]), ("loop", bits, which, it); n += 1
    # motion compensation
    for avg in (0, 1):
        for sidx in range(3 if avg else 4):
            nn = 16 >> sidx
            for mc in range(16):
                src, dst = pixels((nn + 8, 48)), pixels((nn, 48))
                if mc == 10:
                    src[:] = top                 # the centre position at full scale: the largest intermediate sums
                res = []
                for c in (a, b):
                    d = dst.copy()
                    c.h264_hbd_qpel(bits, avg, sidx, mc, vp(d, 8), vp(src, 2 * (3 * 48 + 8)), 96)
                    res.append(d)
                assert np.array_equal(res[0], res[1]), ("qpel", bits, avg, sidx, mc); n += 1
        for widx in range(3):
            for (fx, fy) in ((0, 0), (3, 0), (0, 5), (7, 7), (1, 6)):
                h = 8 >> widx
                src, dst = pixels((h + 2, 32)), pixels((h, 32))
                res = []
                for c in (a, b):
                    d = dst.copy()
                    c.h264_hbd_chroma(bits, avg, widx, vp(d), vp(src, 4), 64, h, fx, fy)
                    res.append(d)
                assert np.array_equal(res[0], res[1]), ("chroma", bits, avg, widx, fx, fy); n += 1
    n += pred_compare(a, b, bits, rng)
    return n


PRED_TABS = {0: (4, 12), 1: (8, 12), 2: (8, 11), 3: (16, 7)}


def pred422_compare(a, b, bits, seed=0):
    """chroma_format_idc 2: the 8 x 16 pred8x8[] entries and the two pred8x8_add[] entries, bits 8 / 9 / 10"""
    rng = np.random.default_rng(seed + 100 * bits)
    top, sb, n = (1 << bits) - 1, 2 if bits > 8 else 1, 0
    dt, ct = (np.uint16, np.int32) if bits > 8 else (np.uint8, np.int16)
    stride = 64 * sb
    for it in range(14):
        img = rng.integers(0, top + 1, size=(48, 64)).astype(dt)
        if it % 4 == 0:
            img[:] = rng.choice([0, top, top // 2])
            img[::3, ::5] = rng.integers(0, top + 1)
        x, y = 16 + 4 * int(rng.integers(0, 3)), 16 + 4 * int(rng.integers(0, 3))
        for mode in range(11):
            res = []
            for c in (a, b):
                p = img.copy()
                c.h264_pred422(bits, mode, vp(p, sb * (y * 64 + x)), stride)
                res.append(p)
            assert np.array_equal(res[0], res[1]), ("pred422", bits, mode, np.argwhere(res[0] != res[1])[:4].tolist()); n += 1
        bo = np.zeros(16, np.int32)
        for k in range(4):
            bo[k] = sb * (4 * (k & 1) + 4 * (k >> 1) * 64)
            bo[8 + k] = sb * (4 * (k & 1) + (8 + 4 * (k >> 1)) * 64)
        blk = rng.integers(-300 << (bits - 8), (300 << (bits - 8)) + 1, size=128).astype(ct)
        for add_mode in (0, 1):
            res = []
            for c in (a, b):
                p, k = img.copy(), blk.copy()
                c.h264_pred422_add(bits, add_mode, vp(p, sb * (16 * 64 + 16)), vp(bo), vp(k), stride)
                res.append((p, k))
            assert np.array_equal(res[0][0], res[1][0]) and not res[0][1].any() and not res[1][1].any(), ("pred422_add", bits, add_mode); n += 1
    return n


def pred_compare(a, b, bits, rng):
    """H264PredContext: every mode of pred4x4 / pred8x8l (all has_topleft / has_topright pairs) / pred8x8 / pred16x16 and the ten lossless
    *_add predictors on 64-sample-wide pictures of 16-bit samples; whole pictures are compared (a stray write shows)"""
    top, stride, n = (1 << bits) - 1, 128, 0
    for tab, (size, modes) in PRED_TABS.items():
        for it in range(10):
            img = rng.integers(0, top + 1, size=(48, 64)).astype(np.uint16)
            if it % 4 == 0:
                img[:] = rng.choice([0, top, top // 2])        # flat pictures: plane prediction clipping
                img[::3, ::5] = rng.integers(0, top + 1)
            x, y = 16 + 4 * int(rng.integers(0, 3)), 16 + 4 * int(rng.integers(0, 3))
            trbuf = img[y - 1, x + size:x + size + 4].copy() if it & 1 else np.full(4, img[y - 1, x + size - 1], np.uint16)
            for mode in range(modes):
                for tl, tr in ((0, 0), (1, 1), (1, 0), (0, 1)) if tab == 1 else ((0, 0),):
                    res = []
                    for c in (a, b):
                        p = img.copy()
                        c.h264_hbd_pred(bits, tab, mode, vp(p, 2 * (y * 64 + x)), vp(trbuf), tl, tr, stride)
                        res.append(p)
                    assert np.array_equal(res[0], res[1]), ("pred", bits, tab, mode, tl, tr, np.argwhere(res[0] != res[1])[:4].tolist()); n += 1
    for tab in range(5):
        nco = {0: 16, 1: 64, 2: 64, 3: 64, 4: 256}[tab]
        chroma = tab == 3
        bo = np.array([2 * (4 * ((k & 1) if chroma else (k & 1) + 2 * ((k >> 2) & 1)) + 4 * ((k >> 1) if chroma else ((k >> 1) & 1) + 2 * (k >> 3)) * 64)
                       for k in range(4 if chroma else 16)], np.int32)              # frame-macroblock offsets in bytes
        for it in range(12):
            img = rng.integers(0, top + 1, size=(48, 64)).astype(np.uint16)
            blk = rng.integers(-300 << (bits - 8), (300 << (bits - 8)) + 1, size=nco).astype(np.int32)
            for mode in (0, 1):
                for tl, tr in ((0, 0), (1, 1), (1, 0), (0, 1)) if tab == 2 else ((0, 0),):
                    res = []
                    for c in (a, b):
                        p, k = img.copy(), blk.copy()
                        c.h264_hbd_pred_add(bits, tab, mode, vp(p, 2 * (16 * 64 + 16)), vp(bo), vp(k), tl, tr, stride)
                        res.append((p, k))
                    assert np.array_equal(res[0][0], res[1][0]) and not res[0][1].any() and not res[1][1].any(), ("pred_add", bits, tab, mode, tl, tr); n += 1
    return n


class TableCallee:
    """the oracle's h264_hbd_* entry names served by the product's tables: ff_h264dsp_init_cuda(c, bits, idc), ff_h264qpel_init_cuda(c, bits),
    ff_h264chroma_init_cuda(c, bits) of `lib` (the product library on a GPU, or the host simulation)"""

    def __init__(self, lib, pred_init=None):
        """pred_init(table, bits): fills an H264PredContext; default = the product's ff_h264_pred_init_cuda(h, AV_CODEC_ID_H264, bits, 1)"""
        from libav_b200 import tables
        self.t, self.p = {}, {}
        for bits in (9, 10):
            h = tables.H264PredContext()
            if pred_init:
                pred_init(C.byref(h), bits)
            else:
                lib.ff_h264_pred_init_cuda(C.byref(h), 27, bits, 1)
            assert h.pred4x4[0] and h.pred16x16[3] and h.pred8x8_add[2], bits
            self.p[bits] = h
        for bits in (9, 10):
            d1, d2, q, ch = tables.H264DSPContext(), tables.H264DSPContext(), tables.H264QpelContext(), tables.H264ChromaContext()
            lib.ff_h264dsp_init_cuda(C.byref(d1), bits, 1); lib.ff_h264dsp_init_cuda(C.byref(d2), bits, 2)
            lib.ff_h264qpel_init_cuda(C.byref(q), bits); lib.ff_h264chroma_init_cuda(C.byref(ch), bits)
            assert d1.h264_idct_add and d2.h264_idct_add8 and q.put_h264_qpel_pixels_tab[0][0] and ch.put_h264_chroma_pixels_tab[0], bits
            self.t[bits] = (d1, d2, q, ch)

    @staticmethod
    def u8(v):
        return C.cast(v, C.POINTER(C.c_uint8))

    @staticmethod
    def i16(v):
        return C.cast(v, C.POINTER(C.c_int16))

    def h264_hbd_idct(self, bits, which, dst, block, stride):
        d = self.t[bits][0]
        (d.h264_idct_add, d.h264_idct8_add, d.h264_idct_dc_add, d.h264_idct8_dc_add)[which](self.u8(dst), self.i16(block), stride)

    def h264_hbd_idct_mb(self, bits, which, dst, dst2, bo, block, stride, nnzc):
        d = self.t[bits][1 if which == 4 else 0]
        bo, nn = C.cast(bo, C.POINTER(C.c_int)), self.u8(nnzc)
        if which < 3:
            (d.h264_idct_add16, d.h264_idct_add16intra, d.h264_idct8_add4)[which](self.u8(dst), bo, self.i16(block), stride, nn)
        else:
            u8p = C.POINTER(C.c_uint8)
            d.h264_idct_add8((u8p * 2)(C.cast(dst2[0], u8p), C.cast(dst2[1], u8p)), bo, self.i16(block), stride, nn)

    def h264_hbd_dc_dequant(self, bits, kind, out, inp, qmul):
        if kind == 0:
            self.t[bits][0].h264_luma_dc_dequant_idct(self.i16(out), self.i16(inp), qmul)
        else:
            self.t[bits][kind - 1].h264_chroma_dc_dequant_idct(self.i16(out), qmul)

    def h264_hbd_add_pixels_clear(self, bits, w8, dst, block, stride):
        d = self.t[bits][0]
        (d.h264_add_pixels8_clear if w8 else d.h264_add_pixels4_clear)(self.u8(dst), self.i16(block), stride)

    def h264_hbd_weight(self, bits, widx, block, stride, height, ld, w, off):
        self.t[bits][0].weight_h264_pixels_tab[widx](self.u8(block), stride, height, ld, w, off)

    def h264_hbd_biweight(self, bits, widx, dst, src, stride, height, ld, wd, ws, off):
        self.t[bits][0].biweight_h264_pixels_tab[widx](self.u8(dst), self.u8(src), stride, height, ld, wd, ws, off)

    def h264_hbd_loop_filter(self, bits, which, pix, stride, alpha, beta, tc0):
        d = self.t[bits][1 if which >= 12 else 0]
        i8 = C.cast(tc0, C.POINTER(C.c_int8))
        lf = {0: d.h264_v_loop_filter_luma, 1: d.h264_h_loop_filter_luma, 4: d.h264_v_loop_filter_chroma, 5: d.h264_h_loop_filter_chroma, 12: d.h264_h_loop_filter_chroma,
              8: d.h264_h_loop_filter_luma_mbaff, 10: d.h264_h_loop_filter_chroma_mbaff, 14: d.h264_h_loop_filter_chroma_mbaff}
        lfi = {2: d.h264_v_loop_filter_luma_intra, 3: d.h264_h_loop_filter_luma_intra, 6: d.h264_v_loop_filter_chroma_intra, 7: d.h264_h_loop_filter_chroma_intra,
               13: d.h264_h_loop_filter_chroma_intra, 9: d.h264_h_loop_filter_luma_mbaff_intra, 11: d.h264_h_loop_filter_chroma_mbaff_intra,
               15: d.h264_h_loop_filter_chroma_mbaff_intra}
        if which in lf:
            lf[which](self.u8(pix), stride, alpha, beta, i8)
        else:
            lfi[which](self.u8(pix), stride, alpha, beta)

    def h264_hbd_qpel(self, bits, avg, sidx, mc, dst, src, stride):
        q = self.t[bits][2]
        (q.avg_h264_qpel_pixels_tab if avg else q.put_h264_qpel_pixels_tab)[sidx][mc](self.u8(dst), self.u8(src), stride)

    def h264_hbd_chroma(self, bits, avg, widx, dst, src, stride, h, x, y):
        ch = self.t[bits][3]
        (ch.avg_h264_chroma_pixels_tab if avg else ch.put_h264_chroma_pixels_tab)[widx](self.u8(dst), self.u8(src), stride, h, x, y)

    def h264_hbd_pred(self, bits, tab, mode, src, topright, tl, tr, stride):
        h = self.p[bits]
        if tab == 0:
            h.pred4x4[mode](src, topright, stride)
        elif tab == 1:
            h.pred8x8l[mode](src, tl, tr, stride)
        elif tab == 2:
            h.pred8x8[mode](src, stride)
        else:
            h.pred16x16[mode](src, stride)

    def h264_hbd_pred_add(self, bits, tab, mode, pix, bo, block, tl, tr, stride):
        h = self.p[bits]
        if tab == 0:
            h.pred4x4_add[mode](pix, block, stride)
        elif tab == 1:
            h.pred8x8l_add[mode](pix, block, stride)
        elif tab == 2:
            h.pred8x8l_filter_add[mode](pix, block, tl, tr, stride)
        elif tab == 3:
            h.pred8x8_add[1 if mode else 2](pix, bo, block, stride)          # [HOR_PRED8x8 = 1], [VERT_PRED8x8 = 2]
        else:
            h.pred16x16_add[1 if mode else 2](pix, bo, block, stride)


class Pred422Callee:
    """h264_pred422 / h264_pred422_add served by an H264PredContext filled for chroma_format_idc 2: fill(table_ref, bits)"""

    def __init__(self, fill):
        from libav_b200 import tables
        self.h = {}
        for bits in (8, 9, 10):
            h = tables.H264PredContext()
            fill(C.byref(h), bits)
            assert all(C.cast(h.pred8x8[m], C.c_void_p).value for m in range(11)) and h.pred8x8_add[1] and h.pred8x8_add[2], bits
            self.h[bits] = h

    def h264_pred422(self, bits, mode, src, stride):
        self.h[bits].pred8x8[mode](src, stride)

    def h264_pred422_add(self, bits, add_mode, pix, bo, block, stride):
        self.h[bits].pred8x8_add[1 if add_mode else 2](pix, bo, block, stride)
