"""CPU: pin the swscale restatement (oracle/port/orc_sws.c) and the product's host-side set-up stage
(libav_b200/csrc/sws_filter.cu, reached through sws_debug_*_cuda -- no GPU involved) against
 (a) known answers measured on the reference by the survey: av_crc(AV_CRC_32_IEEE)
     of the rgb24 output for the LFG-seed-1 frame = f046d710 (640x480) / e1558c0a (3840x2160), and
 (b) the unmodified reference in oracle/_ref: filter banks, colour tables and whole frames, byte for byte."""
import ctypes as C

import numpy as np
import pytest

import libav_b200._lib as L
from libav_b200 import synth
from oracle.loader import ptr

BICUBIC, ACC = 4, 0x40000 | 0x80000
ALGOS = {"fast_bilinear": 1, "bilinear": 2, "bicubic": 4, "x": 8, "point": 0x10, "area": 0x20, "bicublin": 0x40,
         "gauss": 0x80, "sinc": 0x100, "lanczos": 0x200, "spline": 0x400}
GEOMS = [(640, 480, 640, 480), (352, 288, 640, 480), (640, 480, 352, 288), (96, 96, 64, 128), (100, 37, 333, 211),
         (1000, 700, 123, 77), (17, 9, 24, 31)]


def _planes(arrs):
    return (C.c_void_p * 3)(*[a.ctypes.data for a in arrs]), (C.c_int * 3)(*[a.strides[0] for a in arrs])


def to_rgb(o, yuv, dw, dh, flags, pad=0):
    y, u, v = yuv
    dst = np.zeros((dh, dw * 3 + pad), dtype=np.uint8)
    p, s = _planes(yuv)
    r = o.sws_yuv420p_to_rgb24(p, s, y.shape[1], y.shape[0], ptr(dst), dst.strides[0], dw, dh, flags)
    return r, dst


def to_yuv(o, yuv, dw, dh, flags):
    y = yuv[0]
    out = [np.zeros((dh, dw), np.uint8), np.zeros(((dh + 1) // 2, (dw + 1) // 2), np.uint8), np.zeros(((dh + 1) // 2, (dw + 1) // 2), np.uint8)]
    p, s = _planes(yuv)
    dp, ds = _planes(out)
    r = o.sws_yuv420p_to_yuv420p(p, s, y.shape[1], y.shape[0], dp, ds, dw, dh, flags)
    return r, out


@pytest.mark.parametrize("w,h,crc", [(640, 480, 0xF046D710), (3840, 2160, 0xE1558C0A)])
def test_known_answer_crc(checker, orc, w, h, crc):
    yuv = synth.yuv420p_frame(w, h, 1)
    for o in {checker, orc}:
        r, rgb = to_rgb(o, yuv, w, h, BICUBIC | ACC)
        assert r == h
        if w <= 640 or o is checker:           # the pure-python CRC is slow; one 4K pass is enough
            assert synth.crc32_ieee_be(rgb.tobytes()) == crc
    if checker is not orc:
        assert np.array_equal(to_rgb(checker, yuv, w, h, BICUBIC | ACC)[1], to_rgb(orc, yuv, w, h, BICUBIC | ACC)[1])


def _bank(fn, *args):
    cap = 1 << 21
    f, p, n = np.zeros(cap, np.int16), np.zeros(cap, np.int32), C.c_int(0)
    taps = fn(*args, ptr(f), ptr(p), cap, C.byref(n))
    if taps < 0:
        return taps, None, None
    return taps, f[:taps * n.value].copy(), p[:n.value].copy()


@pytest.mark.parametrize("algo", list(ALGOS))
def test_filter_banks_match_reference(refo, orc, algo):
    flags = ALGOS[algo] | ACC
    for rgb in (1, 0):
        for (sw, sh, dw, dh) in GEOMS:
            for which in range(4):
                a = _bank(refo.sws_get_filter, which, rgb, sw, sh, dw, dh, flags)
                b = _bank(orc.sws_get_filter, which, rgb, sw, sh, dw, dh, flags)
                c = _bank(L.lib.sws_debug_filter_cuda, which, sw, sh, dw, dh, 2 if rgb else 0, flags)
                L.lib.avb200_clear_error()
                if a[0] == -3:                   # unscaled yuv->yuv: the reference has no banks (plane copy)
                    continue
                for other, who in ((b, "port"), (c, "product")):
                    assert a[0] == other[0], (who, algo, rgb, sw, sh, dw, dh, which)
                    if a[0] > 0:
                        assert np.array_equal(a[1], other[1]) and np.array_equal(a[2], other[2]), (who, algo, rgb, sw, sh, dw, dh, which)


def _tables(o):
    yt = np.zeros(1024, np.uint8)
    t = [np.zeros(256, np.int32) for _ in range(4)]
    o.sws_rgb24_tables(ptr(yt), *[ptr(x) for x in t])
    return [yt] + t


def test_colour_tables(checker, orc):
    a, b = _tables(checker), _tables(orc)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # the product evaluates the table arithmetically: its constants must regenerate every entry
    k = np.zeros(10, np.int32)
    L.lib.sws_debug_rgb_constants_cuda(ptr(k))
    cy, k1, crv, cgu, cgv, cbu, ar, agu, agv, ab = [int(v) for v in k]
    i = np.arange(1024, dtype=np.int64)
    assert np.array_equal(np.clip((cy * i + k1) >> 16, 0, 255).astype(np.uint8), a[0])
    j = np.arange(256, dtype=np.int64)
    assert np.array_equal(ar + ((j * crv) >> 16), a[1])
    assert np.array_equal(agu + ((j * cgu) >> 16), a[2])
    assert np.array_equal(agv + ((j * cgv) >> 16), a[3])
    assert np.array_equal(ab + ((j * cbu) >> 16), a[4])
    # every index the kernels can form stays inside the 1024-entry table
    lo = min(a[1].min(), a[4].min(), (a[2][:, None] + a[3][None, :]).min())
    hi = max(a[1].max(), a[4].max(), (a[2][:, None] + a[3][None, :]).max()) + 255
    assert lo >= 0 and hi < 1024


@pytest.mark.parametrize("algo", ["bicubic", "bilinear", "fast_bilinear", "point", "area", "lanczos", "bicublin"])
def test_frames_match_reference(refo, orc, algo):
    flags = ALGOS[algo] | ACC
    for (sw, sh, dw, dh) in GEOMS[:6] + [(64, 48, 64, 48), (66, 50, 33, 25)]:
        yuv = tuple(synth.pad_rows(pl) for pl in synth.yuv420p_frame(sw, sh, 3))
        ra, a = to_rgb(refo, yuv, dw, dh, flags, pad=6)
        rb, b = to_rgb(orc, yuv, dw, dh, flags, pad=6)
        assert ra == rb == dh
        assert np.array_equal(a, b), ("rgb", algo, sw, sh, dw, dh)
        ra, a = to_yuv(refo, yuv, dw, dh, flags)
        rb, b = to_yuv(orc, yuv, dw, dh, flags)
        assert ra == rb == dh
        for pa, pb in zip(a, b):
            assert np.array_equal(pa, pb), ("yuv", algo, sw, sh, dw, dh)


def test_unscaled_table_converter(refo, orc):
    """same size, no SWS_ACCURATE_RND, even height: the reference installs yuv2rgb_c_24_rgb (nearest chroma)"""
    for (w, h) in ((64, 48), (66, 50), (70, 36), (641, 480)):
        yuv = synth.yuv420p_frame(w, h, 5)
        ra, a = to_rgb(refo, yuv, w, h, BICUBIC, pad=6)
        rb, b = to_rgb(orc, yuv, w, h, BICUBIC, pad=6)
        assert ra == rb == h
        assert np.array_equal(a, b), (w, h)
        assert not np.array_equal(a, to_rgb(refo, yuv, w, h, BICUBIC | ACC, pad=6)[1])      # it really is a different path


def test_full_chroma_interpolation_port_matches_reference(orc, refo):
    """SWS_FULL_CHR_H_INT (yuv2rgb24_full_X_c): one chroma sample per pixel and the 30-bit colour matrix"""
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    import ctypes as C
    from libav_b200 import synth
    from oracle.loader import ptr
    FULL = 0x2000
    for (w, h, dw, dh) in [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (100, 37, 333, 211), (66, 50, 33, 25), (641, 479, 641, 479)]:
        yuv = tuple(synth.pad_rows(pl) for pl in synth.yuv420p_frame(w, h, 3))
        p, s = (C.c_void_p * 3)(*[a.ctypes.data for a in yuv]), (C.c_int * 3)(*[a.strides[0] for a in yuv])
        for fl in (4 | 0xC0000, 2 | 0xC0000, 4, 0x10 | 0xC0000, 1 | 0xC0000, 0x200 | 0xC0000):
            outs = []
            for o in (refo, orc):
                dst = np.full((dh, dw * 3 + 6), 9, np.uint8)
                assert o.sws_yuv420p_to_rgb24(p, s, w, h, ptr(dst), dst.strides[0], dw, dh, fl | FULL) == dh
                outs.append(dst)
            assert np.array_equal(outs[0], outs[1]), (w, h, dw, dh, hex(fl))
