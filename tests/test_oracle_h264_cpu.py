"""CPU: pin the H.264 restatement (oracle/port/orc_h264.c) byte-for-byte against the unmodified reference in
oracle/_ref, slot by slot, on checkasm-style random inputs (tests/checkasm/h264dsp.c, h264qpel.c recipes: random
pixels, coefficient blocks with the nnz patterns the dispatchers distinguish, alpha/beta/tc0 ladders)."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr
from h264_util import at, block_offsets, block_offsets_422, residual_422


def both(orc, refo, fn, make_args):
    """Run fn on both oracles with independently built (identical) arguments; return the two argument sets."""
    a, b = make_args(), make_args()
    getattr(orc, fn)(*a[0])
    getattr(refo, fn)(*b[0])
    return a[1], b[1]


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_idct_single(orc, refo, which):
    rng = np.random.default_rng(which)
    for it in range(200):
        n = 64 if which in (1, 3) else 16
        blk = rng.integers(-2000, 2000, size=n).astype(np.int16)
        if it % 7 == 0:
            blk[:] = rng.integers(-32768, 32767, size=n)           # int16 wrap territory
        pix = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)

        def mk():
            b, p = blk.copy(), pix.copy()
            return (which, at(p, 8), ptr(b), 32), (b, p)
        x, y = both(orc, refo, "h264_idct", mk)
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_idct_mb_dispatch(orc, refo, which):
    for seed in range(40):
        rec, coeffs, nnzc = synth.h264_residual_work(1, 1, seed=seed, modes=(which if which < 3 else 0,))
        rng = np.random.default_rng(seed)
        y = rng.integers(0, 256, size=(16, 48), dtype=np.uint8)
        cb = rng.integers(0, 256, size=(8, 24), dtype=np.uint8)
        cr = rng.integers(0, 256, size=(8, 24), dtype=np.uint8)
        bo = block_offsets(48, 24)
        res = []
        for o in (orc, refo):
            yy, b, r, c = y.copy(), cb.copy(), cr.copy(), coeffs[0].copy()
            d2 = (C.c_void_p * 2)(b.ctypes.data + 8, r.ctypes.data + 8)
            o.h264_idct_mb(which, at(yy, 16), d2, ptr(bo), ptr(c), 48 if which < 3 else 24, ptr(nnzc[0]))
            res.append((yy, b, r, c))
        for u, v in zip(*res):
            assert np.array_equal(u, v)


def test_dc_dequant(orc, refo):
    rng = np.random.default_rng(1)
    for _ in range(100):
        inp = rng.integers(-3000, 3000, size=16).astype(np.int16)
        q = int(rng.integers(1, 4000))
        a, b = np.full(256, 7, np.int16), np.full(256, 7, np.int16)
        orc.h264_luma_dc_dequant_idct(ptr(a), ptr(inp.copy()), q)
        refo.h264_luma_dc_dequant_idct(ptr(b), ptr(inp.copy()), q)
        assert np.array_equal(a, b)
        blk = rng.integers(-3000, 3000, size=64).astype(np.int16)
        a, b = blk.copy(), blk.copy()
        orc.h264_chroma_dc_dequant_idct(ptr(a), q)
        refo.h264_chroma_dc_dequant_idct(ptr(b), q)
        assert np.array_equal(a, b)
        blk = rng.integers(-3000, 3000, size=128).astype(np.int16)        # 4:2:2: eight DC values, 16 coefficients apart
        a, b = blk.copy(), blk.copy()
        orc.h264_chroma422_dc_dequant_idct(ptr(a), q)
        refo.h264_chroma422_dc_dequant_idct(ptr(b), q)
        assert np.array_equal(a, b) and not np.array_equal(a, blk)


def test_add_pixels_clear_and_weight(orc, refo):
    rng = np.random.default_rng(2)
    for w8 in (0, 1):
        n = 8 if w8 else 4
        blk = rng.integers(-300, 300, size=n * n).astype(np.int16)
        pix = rng.integers(0, 256, size=(8, 16), dtype=np.uint8)
        res = []
        for o in (orc, refo):
            p, b = pix.copy(), blk.copy()
            o.h264_add_pixels_clear(w8, ptr(p), ptr(b), 16)
            res.append((p, b))
        assert np.array_equal(res[0][0], res[1][0]) and not res[0][1].any() and not res[1][1].any()
    for widx in range(4):
        for _ in range(50):
            ld, w, off = int(rng.integers(0, 8)), int(rng.integers(-128, 128)), int(rng.integers(-128, 128))
            hgt = int(rng.choice([2, 4, 8, 16]))
            pix = rng.integers(0, 256, size=(16, 32), dtype=np.uint8)
            src = rng.integers(0, 256, size=(16, 32), dtype=np.uint8)
            a, b = pix.copy(), pix.copy()
            orc.h264_weight(widx, ptr(a), 32, hgt, ld, w, off)
            refo.h264_weight(widx, ptr(b), 32, hgt, ld, w, off)
            assert np.array_equal(a, b)
            a, b = pix.copy(), pix.copy()
            ws = int(rng.integers(-128, 128))
            orc.h264_biweight(widx, ptr(a), ptr(src), 32, hgt, ld, w, ws, off)
            refo.h264_biweight(widx, ptr(b), ptr(src), 32, hgt, ld, w, ws, off)
            assert np.array_equal(a, b)


@pytest.mark.parametrize("which", range(16))        # 8..15: mbaff and chroma_format_idc 2 entries
def test_loop_filters(orc, refo, which):
    rng = np.random.default_rng(which)
    for it in range(300):
        # smooth-ish content so that the alpha/beta conditions trigger, plus noise
        base = rng.integers(0, 256)
        pix = np.clip(base + rng.integers(-12, 13, size=(24, 32)), 0, 255).astype(np.uint8)
        if it % 5 == 0:
            pix = rng.integers(0, 256, size=(24, 32), dtype=np.uint8)
        alpha, beta = int(rng.integers(0, 256)), int(rng.integers(0, 19))
        tc0 = rng.integers(-1, 26, size=4).astype(np.int8)
        a, b = pix.copy(), pix.copy()
        orc.h264_loop_filter(which, at(a, 8 * 32 + 8), 32, alpha, beta, ptr(tc0))
        refo.h264_loop_filter(which, at(b, 8 * 32 + 8), 32, alpha, beta, ptr(tc0))
        assert np.array_equal(a, b), (which, it)


@pytest.mark.parametrize("avg", [0, 1])
@pytest.mark.parametrize("sidx", [0, 1, 2, 3])
def test_qpel_all_positions(orc, refo, avg, sidx):
    if avg and sidx == 3:
        pytest.skip("the reference has no avg 2x2 functions (h264qpel.c:60-67)")
    rng = np.random.default_rng(10 * avg + sidx)
    n = 16 >> sidx
    for mc in range(16):
        for _ in range(6):
            src = rng.integers(0, 256, size=(n + 8, 48), dtype=np.uint8)
            dst = rng.integers(0, 256, size=(n, 48), dtype=np.uint8)
            a, b = dst.copy(), dst.copy()
            orc.h264_qpel(avg, sidx, mc, ptr(a), at(src, 3 * 48 + 8), 48)
            refo.h264_qpel(avg, sidx, mc, ptr(b), at(src, 3 * 48 + 8), 48)
            assert np.array_equal(a, b), (avg, sidx, mc)


@pytest.mark.parametrize("avg", [0, 1])
def test_chroma_mc(orc, refo, avg):
    rng = np.random.default_rng(avg)
    for widx in range(3):
        for x in range(8):
            for y in range(8):
                h = int(rng.choice([2, 4, 8] if widx else [4, 8, 16]))
                src = rng.integers(0, 256, size=(h + 2, 32), dtype=np.uint8)
                dst = rng.integers(0, 256, size=(h, 32), dtype=np.uint8)
                a, b = dst.copy(), dst.copy()
                orc.h264_chroma(avg, widx, ptr(a), ptr(src), 32, h, x, y)
                refo.h264_chroma(avg, widx, ptr(b), ptr(src), 32, h, x, y)
                assert np.array_equal(a, b), (avg, widx, x, y)
