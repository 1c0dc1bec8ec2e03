"""GPU: the reference's own tables, filled by ff_*_init_cuda(), called slot by slot with HOST pointers exactly like a codec
(or tests/checkasm) calls them, and byte-compared with the CPU oracle's slot -- outputs AND clobbered inputs (checkasm's
rule, tests/checkasm/h264dsp.c:213-215)."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr
from h264_util import at, block_offsets

pytestmark = pytest.mark.gpu
u8p, i16p, i8p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int8)


def P(a, off=0, t=u8p):
    return C.cast(a.ctypes.data + int(off), t)


def test_fdct_slots(gpu, checker):
    from libav_b200 import tables
    rng = np.random.default_rng(0)
    for algo, base in ((0, 0), (2, 0), (1, 2)):                 # FF_DCT_AUTO, FF_DCT_INT -> islow ; FF_DCT_FASTINT -> ifast
        c = tables.FDCTDSPContext()
        gpu.lib.ff_fdctdsp_init_cuda(C.byref(c), algo, 8, 0)
        for _ in range(20):
            blk = rng.integers(-256, 256, size=64).astype(np.int16)
            for slot, which in ((c.fdct, base), (c.fdct248, base + 1)):
                a, b = blk.copy(), blk.copy()
                slot(P(a, 0, i16p))
                checker.fdct(which, ptr(b))
                assert np.array_equal(a, b)
    assert gpu.last_error() == ""


def test_me_cmp_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.MECmpContext()
    gpu.lib.ff_me_cmp_init_cuda(C.byref(c))
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=(24, 48), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    cases = [(c.pix_abs[s][d], 0, s, d) for s in (0, 1) for d in range(4)] + \
            [(c.sad[0], 1, 0, 0), (c.sad[1], 1, 1, 0), (c.sse[0], 2, 0, 0), (c.sse[1], 2, 1, 0), (c.sse[2], 2, 2, 0),
             (c.hadamard8_diff[0], 3, 0, 0), (c.hadamard8_diff[1], 3, 1, 0), (c.hadamard8_diff[4], 7, 0, 0), (c.hadamard8_diff[5], 7, 1, 0),
             (c.vsad[0], 4, 0, 0), (c.vsad[4], 8, 0, 0), (c.vsad[5], 8, 1, 0), (c.vsse[0], 5, 0, 0), (c.vsse[4], 9, 0, 0), (c.vsse[5], 9, 1, 0),
             (c.nsse[0], 6, 0, 0), (c.nsse[1], 6, 1, 0)]
    for slot, kind, sidx, dxy in cases:
        w = 16 >> sidx
        h = 8 if w < 16 else 16
        got = slot(None, P(a, 48 * 2 + 16), P(b, 48 * 3 + 5), 48, h)
        assert got == checker.me_cmp(kind, sidx, dxy, at(a, 48 * 2 + 16), at(b, 48 * 3 + 5), 48, h), (kind, sidx, dxy)
    blk = rng.integers(-500, 500, size=64).astype(np.int16)
    assert c.sum_abs_dctelem(P(blk, 0, i16p)) == checker.me_cmp(10, 0, 0, ptr(blk), None, 0, 0)
    assert not c.dct_sad[0] and not c.rd[0] and not c.bit[0]      # encoder-state metrics are not taken over
    assert gpu.last_error() == ""


def test_hpel_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.HpelDSPContext()
    gpu.lib.ff_hpeldsp_init_cuda(C.byref(c), 0)
    rng = np.random.default_rng(2)
    tabs = {0: c.put_pixels_tab, 1: c.avg_pixels_tab, 2: c.put_no_rnd_pixels_tab}
    for tab in range(4):
        for sidx in range(4):
            for dxy in range(4):
                if tab == 3:
                    slot = c.avg_no_rnd_pixels_tab[dxy] if sidx == 0 else None
                else:
                    slot = tabs[tab][sidx][dxy]
                if (tab == 2 and sidx > 1) or (tab == 3 and sidx):
                    assert not slot
                    continue
                w = 16 >> sidx
                h = max(w // 2, 2)
                pix = rng.integers(0, 256, size=(h + 2, 40), dtype=np.uint8)
                blk = rng.integers(0, 256, size=(h, 40), dtype=np.uint8)
                x, y = blk.copy(), blk.copy()
                slot(P(x, 8), P(pix, 3), 40, h)
                assert checker.hpel(tab, sidx, dxy, at(y, 8), at(pix, 3), 40, h) == 0
                assert np.array_equal(x, y), (tab, sidx, dxy)
    assert gpu.last_error() == ""


def test_h264_qpel_and_chroma_slots(gpu, checker):
    from libav_b200 import tables
    q, ch = tables.H264QpelContext(), tables.H264ChromaContext()
    gpu.lib.ff_h264qpel_init_cuda(C.byref(q), 8)
    gpu.lib.ff_h264chroma_init_cuda(C.byref(ch), 8)
    rng = np.random.default_rng(3)
    for avg in (0, 1):
        tab = q.avg_h264_qpel_pixels_tab if avg else q.put_h264_qpel_pixels_tab
        for sidx in range(4):
            if avg and sidx == 3:
                assert not tab[3][0]
                continue
            n = 16 >> sidx
            for mc in range(16):
                src = rng.integers(0, 256, size=(n + 8, 48), dtype=np.uint8)
                dst = rng.integers(0, 256, size=(n, 48), dtype=np.uint8)
                x, y = dst.copy(), dst.copy()
                tab[sidx][mc](P(x, 4), P(src, 3 * 48 + 8), 48)
                checker.h264_qpel(avg, sidx, mc, at(y, 4), at(src, 3 * 48 + 8), 48)
                assert np.array_equal(x, y), (avg, sidx, mc)
        ctab = ch.avg_h264_chroma_pixels_tab if avg else ch.put_h264_chroma_pixels_tab
        for widx in range(3):
            for (fx, fy) in ((0, 0), (3, 0), (0, 5), (7, 7), (1, 6)):
                h = 8 >> widx
                src = rng.integers(0, 256, size=(h + 2, 32), dtype=np.uint8)
                dst = rng.integers(0, 256, size=(h, 32), dtype=np.uint8)
                x, y = dst.copy(), dst.copy()
                ctab[widx](P(x), P(src, 2), 32, h, fx, fy)
                checker.h264_chroma(avg, widx, ptr(y), at(src, 2), 32, h, fx, fy)
                assert np.array_equal(x, y), (avg, widx, fx, fy)
    assert gpu.last_error() == ""


def test_h264dsp_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.H264DSPContext()
    gpu.lib.ff_h264dsp_init_cuda(C.byref(c), 8, 1)
    rng = np.random.default_rng(4)
    # single-block transforms: dst and the (zeroed) block must both match
    for which, slot in enumerate((c.h264_idct_add, c.h264_idct8_add, c.h264_idct_dc_add, c.h264_idct8_dc_add)):
        for _ in range(10):
            n = 64 if which in (1, 3) else 16
            blk = rng.integers(-2000, 2000, size=n).astype(np.int16)
            pix = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
            x, bx, y, by = pix.copy(), blk.copy(), pix.copy(), blk.copy()
            slot(P(x, 8), P(bx, 0, i16p), 32)
            checker.h264_idct(which, at(y, 8), ptr(by), 32)
            assert np.array_equal(x, y) and np.array_equal(bx, by), which
    # per-MB dispatchers with the frame-MB block offsets
    bo = block_offsets(48, 24)
    for which, slot in ((0, c.h264_idct_add16), (1, c.h264_idct_add16intra), (2, c.h264_idct8_add4), (3, c.h264_idct_add8)):
        for seed in range(8):
            rec, coeffs, nnzc = synth.h264_residual_work(1, 1, seed=seed, modes=(which if which < 3 else 0,))
            yp = rng.integers(0, 256, size=(16, 48), dtype=np.uint8)
            cb = rng.integers(0, 256, size=(8, 24), dtype=np.uint8)
            cr = rng.integers(0, 256, size=(8, 24), dtype=np.uint8)
            got = [yp.copy(), cb.copy(), cr.copy(), coeffs[0].copy()]
            want = [yp.copy(), cb.copy(), cr.copy(), coeffs[0].copy()]
            if which < 3:
                slot(P(got[0], 16), P(bo, 0, C.POINTER(C.c_int)), P(got[3], 0, i16p), 48, P(nnzc[0]))
            else:
                d2 = (u8p * 2)(P(got[1], 8), P(got[2], 8))
                slot(d2, P(bo, 0, C.POINTER(C.c_int)), P(got[3], 0, i16p), 24, P(nnzc[0]))
            d2 = (C.c_void_p * 2)(want[1].ctypes.data + 8, want[2].ctypes.data + 8)
            checker.h264_idct_mb(which, at(want[0], 16), d2, ptr(bo), ptr(want[3]), 48 if which < 3 else 24, ptr(nnzc[0]))
            for u, v in zip(got, want):
                assert np.array_equal(u, v), (which, seed)
    # DC dequant
    inp = rng.integers(-3000, 3000, size=16).astype(np.int16)
    a, b = np.full(256, 7, np.int16), np.full(256, 7, np.int16)
    c.h264_luma_dc_dequant_idct(P(a, 0, i16p), P(inp.copy(), 0, i16p), 1234)
    checker.h264_luma_dc_dequant_idct(ptr(b), ptr(inp.copy()), 1234)
    assert np.array_equal(a, b)
    blk = rng.integers(-3000, 3000, size=64).astype(np.int16)
    a, b = blk.copy(), blk.copy()
    c.h264_chroma_dc_dequant_idct(P(a, 0, i16p), 777)
    checker.h264_chroma_dc_dequant_idct(ptr(b), 777)
    assert np.array_equal(a, b)
    # bypass add
    for w8, slot in ((0, c.h264_add_pixels4_clear), (1, c.h264_add_pixels8_clear)):
        n = 8 if w8 else 4
        blk = rng.integers(-300, 300, size=n * n).astype(np.int16)
        pix = rng.integers(0, 256, size=(8, 16), dtype=np.uint8)
        x, bx, y, by = pix.copy(), blk.copy(), pix.copy(), blk.copy()
        slot(P(x), P(bx, 0, i16p), 16)
        checker.h264_add_pixels_clear(w8, ptr(y), ptr(by), 16)
        assert np.array_equal(x, y) and not bx.any()
    # weighted prediction
    for widx in range(4):
        pix = rng.integers(0, 256, size=(16, 32), dtype=np.uint8)
        src = rng.integers(0, 256, size=(16, 32), dtype=np.uint8)
        x, y = pix.copy(), pix.copy()
        c.weight_h264_pixels_tab[widx](P(x), 32, 8, 5, 37, -3)
        checker.h264_weight(widx, ptr(y), 32, 8, 5, 37, -3)
        assert np.array_equal(x, y)
        x, y = pix.copy(), pix.copy()
        c.biweight_h264_pixels_tab[widx](P(x), P(src), 32, 8, 5, 40, 24, 2)
        checker.h264_biweight(widx, ptr(y), ptr(src), 32, 8, 5, 40, 24, 2)
        assert np.array_equal(x, y)
    # loop filters
    slots = [c.h264_v_loop_filter_luma, c.h264_h_loop_filter_luma, c.h264_v_loop_filter_luma_intra, c.h264_h_loop_filter_luma_intra,
             c.h264_v_loop_filter_chroma, c.h264_h_loop_filter_chroma, c.h264_v_loop_filter_chroma_intra, c.h264_h_loop_filter_chroma_intra]
    for which, slot in enumerate(slots):
        for it in range(12):
            base = rng.integers(0, 256)
            pix = np.clip(base + rng.integers(-10, 11, size=(24, 32)), 0, 255).astype(np.uint8)
            alpha, beta = int(rng.integers(1, 200)), int(rng.integers(1, 19))
            tc0 = rng.integers(-1, 12, size=4).astype(np.int8)
            x, y = pix.copy(), pix.copy()
            if which & 2:
                slot(P(x, 8 * 32 + 8), 32, alpha, beta)
            else:
                slot(P(x, 8 * 32 + 8), 32, alpha, beta, P(tc0, 0, i8p))
            checker.h264_loop_filter(which, at(y, 8 * 32 + 8), 32, alpha, beta, ptr(tc0))
            assert np.array_equal(x, y), (which, it)
    assert not c.h264_loop_filter_strength and not c.startcode_find_candidate      # left to the caller, as documented
    assert gpu.last_error() == ""
