"""The H.264 slot cases shared by the GPU tests (tests/test_gpu_slots.py, the product library on a B200) and the host simulation
(tests/test_hostsim_slots_cpu.py, the same slots.cu compiled for the host): the reference's tables, filled by ff_*_init_cuda(),
called slot by slot with HOST pointers like a codec (or tests/checkasm) calls them, byte-compared with the CPU oracle -- outputs
AND clobbered inputs (checkasm's rule, tests/checkasm/h264dsp.c:213-215).  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from libav_b200 import synth
from oracle.loader import ptr
from h264_util import at, block_offsets, block_offsets_422, residual_422

u8p, i16p, i8p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int8)


def P(a, off=0, t=u8p):
    return C.cast(a.ctypes.data + int(off), t)


def qpel_and_chroma_cases(lib, last_error, checker):
    from libav_b200 import tables
    q, ch = tables.H264QpelContext(), tables.H264ChromaContext()
    lib.ff_h264qpel_init_cuda(C.byref(q), 8)
    lib.ff_h264chroma_init_cuda(C.byref(ch), 8)
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
    assert last_error() == ""


def h264dsp_cases(lib, last_error, checker, weights=True):
    from libav_b200 import tables
    c = tables.H264DSPContext()
    lib.ff_h264dsp_init_cuda(C.byref(c), 8, 1)
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
    # weighted prediction (a batch of one of the batched weight kernel)
    for widx in range(4 if weights else 0):
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
    # loop filters: the eight frame filters, then the mbaff variants (oracle numbering 8..11)
    slots = [c.h264_v_loop_filter_luma, c.h264_h_loop_filter_luma, c.h264_v_loop_filter_luma_intra, c.h264_h_loop_filter_luma_intra,
             c.h264_v_loop_filter_chroma, c.h264_h_loop_filter_chroma, c.h264_v_loop_filter_chroma_intra, c.h264_h_loop_filter_chroma_intra,
             c.h264_h_loop_filter_luma_mbaff, c.h264_h_loop_filter_luma_mbaff_intra, c.h264_h_loop_filter_chroma_mbaff, c.h264_h_loop_filter_chroma_mbaff_intra]
    loop_filter_cases(rng, checker, [(which, slot, bool(which & 2) if which < 8 else bool(which & 1)) for which, slot in enumerate(slots)])
    assert not c.h264_loop_filter_strength                                          # NULL in C too (h264dsp.c:124)
    assert c.startcode_find_candidate                                               # (exercised by startcode_cases from the late / hostsim tests)
    assert last_error() == ""


def startcode_cases(rng, slot):
    """ff_startcode_find_candidate_c (libavcodec/startcode.c:31-59): the index of the first zero byte, `size` when there is none before it"""
    for size in (1, 7, 8, 9, 255, 256, 257, 5000, 100001):
        buf = rng.integers(1, 256, size=size + 64, dtype=np.uint8)
        assert slot(P(buf), size) == size                          # no zero at all
        buf[size:] = 0
        assert slot(P(buf), size) == size                          # zeros only in the padding
        for pos in sorted({0, size // 2, size - 1}):
            b2 = buf.copy()
            b2[pos] = 0
            b2[min(pos + 3, size - 1)] = 0
            assert slot(P(b2), size) == pos, (size, pos)
    assert slot(P(np.zeros(16, np.uint8)), 0) == 0


def loop_filter_cases(rng, checker, cases, iters=12):
    """the bytes outside the lines / samples a filter may touch must stay as they were, too: the whole 24 x 32 array is compared"""
    for which, slot, intra in cases:
        for it in range(iters):
            base = rng.integers(0, 256)
            pix = np.clip(base + rng.integers(-10, 11, size=(24, 32)), 0, 255).astype(np.uint8)
            alpha, beta = int(rng.integers(1, 200)), int(rng.integers(1, 19))
            tc0 = rng.integers(-1, 12, size=4).astype(np.int8)
            x, y = pix.copy(), pix.copy()
            if intra:
                slot(P(x, 4 * 32 + 8), 32, alpha, beta)
            else:
                slot(P(x, 4 * 32 + 8), 32, alpha, beta, P(tc0, 0, i8p))
            checker.h264_loop_filter(which, at(y, 4 * 32 + 8), 32, alpha, beta, ptr(tc0))
            assert np.array_equal(x, y), (which, it)


def h264dsp_422_cases(lib, last_error, checker):
    """ff_h264dsp_init_cuda(c, 8, 2): the entries that differ for chroma_format_idc 2 (h264dsp.c:81-122): idct_add8 -> add8_422,
    chroma_dc_dequant_idct -> the 2x4 transform, the four h_ chroma loop filters -> 16 (mbaff: 8) lines"""
    from libav_b200 import tables
    c = tables.H264DSPContext()
    lib.ff_h264dsp_init_cuda(C.byref(c), 8, 2)
    rng = np.random.default_rng(14)
    bo = block_offsets_422(24)
    for seed in range(12):
        coeffs, nnzc = residual_422(seed)
        cb = rng.integers(0, 256, size=(16, 24), dtype=np.uint8)
        cr = rng.integers(0, 256, size=(16, 24), dtype=np.uint8)
        got, want = [cb.copy(), cr.copy(), coeffs.copy()], [cb.copy(), cr.copy(), coeffs.copy()]
        d2 = (u8p * 2)(P(got[0], 8), P(got[1], 8))
        c.h264_idct_add8(d2, P(bo, 0, C.POINTER(C.c_int)), P(got[2], 0, i16p), 24, P(nnzc))
        d2 = (C.c_void_p * 2)(want[0].ctypes.data + 8, want[1].ctypes.data + 8)
        checker.h264_idct_mb(4, None, d2, ptr(bo), ptr(want[2]), 24, ptr(nnzc))
        for u, v in zip(got, want):
            assert np.array_equal(u, v), seed
    for q in [777, 1, 4000] + [int(v) for v in rng.integers(1, 4000, size=40)]:
        blk = rng.integers(-3000, 3000, size=128).astype(np.int16)
        a, b = blk.copy(), blk.copy()
        c.h264_chroma_dc_dequant_idct(P(a, 0, i16p), q)
        checker.h264_chroma422_dc_dequant_idct(ptr(b), q)
        assert np.array_equal(a, b) and not np.array_equal(a, blk)
    loop_filter_cases(rng, checker, [(4, c.h264_v_loop_filter_chroma, False), (6, c.h264_v_loop_filter_chroma_intra, True),
                                     (12, c.h264_h_loop_filter_chroma, False), (13, c.h264_h_loop_filter_chroma_intra, True),
                                     (14, c.h264_h_loop_filter_chroma_mbaff, False), (15, c.h264_h_loop_filter_chroma_mbaff_intra, True),
                                     (0, c.h264_v_loop_filter_luma, False), (9, c.h264_h_loop_filter_luma_mbaff_intra, True)])
    assert last_error() == ""
