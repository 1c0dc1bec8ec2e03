"""CPU: the format / path decisions of sws_getContext_cuda (host logic, no device) -- which requests are taken over, which of
the reference's special converters they map to, the chroma geometry -- and that the product and the oracle port (itself pinned
against the compiled reference) refuse exactly the same requests."""
import ctypes as C

import numpy as np
import pytest

from libav_b200.device import PLANAR_FORMATS

ACC = 0x40000 | 0x80000
PLANAR_SRC = {0: (1, 1), 4: (1, 0), 5: (0, 0), 6: (2, 2), 7: (2, 0), 31: (0, 1), 12: (1, 1), 13: (1, 0), 14: (0, 0), 32: (0, 1), 33: (1, 1), 8: (0, 0)}     # 33: yuva420p (alpha never read), 8: gray8 (one plane; the two others are never read)
PACKED_SRC = {1: 2, 15: 2, 2: 3, 3: 3, 25: 4, 26: 4, 27: 4, 28: 4}
HBD_SRC = [62, 63, 64, 48, 66, 70]                       # 9 / 10 / 16-bit planar sources (a sample; every one in tests/test_sws_hbd_sources_cpu.py)
SRCS = list(PLANAR_SRC) + list(PACKED_SRC) + [23, 24, 11] + HBD_SRC
DSTS = [0, 4, 5, 6, 31, 62, 64, 47, 48, 2, 3, 25, 26, 27, 28, 1, 15, 23, 24, 12, 14, 32, 37, 40, 43, 55, 8, 35, 59]        # 12 14 32: full-range (yuvj) planar
GEOMS = [(64, 48, 64, 48), (66, 50, 66, 50), (64, 48, 96, 80), (96, 80, 64, 48)]
FLAGS = (4 | ACC, 4, 0x10, 1 | ACC, 2 | ACC | 0x2000, 2)


def plan(L, sw, sh, sf, dw, dh, df, flags):
    out = (C.c_int32 * 8)()
    ok = L.lib.sws_debug_plan_cuda(sw, sh, sf, dw, dh, df, flags, out)
    if not ok:
        L.lib.avb200_clear_error()
    return (ok, list(out))


def source(fmt, w, h):
    if fmt in HBD_SRC:
        import test_sws_hbd_sources_cpu as H
        return H.planes(fmt, w, h, 3)
    r = np.random.RandomState(fmt * 7 + w)
    if fmt == 11:                                      # pal8: indices + the palette as plane 1
        return [r.randint(0, 256, (h + 1, w + 16)).astype(np.uint8), r.randint(0, 256, (1, 1024)).astype(np.uint8)]
    if fmt in PACKED_SRC:
        return [r.randint(0, 256, (h + 1, PACKED_SRC[fmt] * w + 16)).astype(np.uint8)]
    if fmt in (23, 24):
        return [r.randint(0, 256, (h, w + 16)).astype(np.uint8), r.randint(0, 256, ((h + 1) // 2, 2 * ((w + 1) // 2) + 16)).astype(np.uint8)]
    hs, vs = PLANAR_SRC[fmt]
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    return [r.randint(0, 256, (h, w + 16)).astype(np.uint8), r.randint(0, 256, (ch, cw + 16)).astype(np.uint8), r.randint(0, 256, (ch, cw + 16)).astype(np.uint8)]


def dest(fmt, w, h):
    if fmt in (23, 24):
        return [np.zeros((h, w + 8), np.uint8), np.zeros(((h + 1) // 2, 2 * ((w + 1) // 2) + 8), np.uint8)]
    if fmt in PLANAR_FORMATS:
        hs, vs, bits = PLANAR_FORMATS[fmt]
        dt = np.uint8 if bits == 8 else np.uint16
        cw, ch = -((-w) >> hs), -((-h) >> vs)
        return [np.zeros((h, w + 8), dt), np.zeros((ch, cw + 8), dt), np.zeros((ch, cw + 8), dt)]
    return [np.zeros((h + 1, w * 6 + 16), np.uint8)]


def port_accepts(orc, sf, pl, sw, sh, df, dw, dh, flags):
    out = dest(df, dw, dh)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return orc.sws_planar(sf, sp, ss, sw, sh, df, dp, ds, dw, dh, flags) == dh


def test_product_and_port_take_over_the_same_requests(built, orc):
    import libav_b200._lib as L
    n = refused = 0
    for sf in SRCS:
        for (sw, sh, dw, dh) in GEOMS:
            pl = source(sf, sw, sh)
            for df in DSTS:
                for flags in FLAGS:
                    packed_rgb = df in (2, 3, 25, 26, 27, 28)
                    if flags & 0x2000 and not packed_rgb:
                        continue             # the reference drops the flag silently for these destinations, the port refuses it
                    ok, _ = plan(L, sw, sh, sf, dw, dh, df, flags)
                    want = port_accepts(orc, sf, pl, sw, sh, df, dw, dh, flags)
                    assert bool(ok) == want, (sf, df, sw, sh, dw, dh, hex(flags), ok, want)
                    n += 1; refused += not ok
    assert n > 5000 and 50 < refused < n // 4


def test_paths_and_geometry(built):
    import libav_b200._lib as L
    P = lambda *a: plan(L, *a)
    assert P(64, 48, 0, 64, 48, 2, 4)[1][0] == 2                    # unscaled table converter (no SWS_ACCURATE_RND, even height)
    assert P(64, 49, 0, 64, 49, 2, 4)[1][0] == 3                    # odd height: swscale(), here the fused same-size kernel
    assert P(64, 48, 0, 64, 48, 2, 4 | ACC)[1][0] == 3
    assert P(64, 48, 4, 64, 48, 2, 4)[1][0] == 2 and P(64, 48, 5, 64, 48, 2, 4)[1][0] == 4
    assert P(64, 48, 0, 64, 48, 0, 4)[1][0] == 1 and P(64, 48, 23, 64, 48, 0, 4)[1][0] == 1 and P(64, 48, 4, 64, 48, 66, 4)[1][0] == 1
    assert P(64, 48, 23, 64, 48, 64, 4)[1][0] == 4                  # nv12 -> 10 bit: no plane copy for semi-planar sources
    assert P(64, 48, 2, 64, 48, 3, 4)[1][0] == 5 and P(64, 48, 3, 64, 48, 0, 4)[1][0] == 5 and P(64, 48, 3, 64, 48, 0, 4 | ACC)[1][0] == 4
    assert P(64, 48, 1, 64, 48, 0, 4 | ACC)[1][0] == 5 and P(64, 48, 15, 64, 48, 4, 4)[1][0] == 5 and P(64, 48, 1, 64, 48, 5, 4)[1][0] == 4
    assert P(64, 48, 4, 64, 48, 1, 4 | ACC)[1][0] == 6 and P(64, 48, 0, 64, 48, 15, 0x10)[1][0] == 6 and P(64, 48, 0, 64, 48, 15, 4)[1][0] == 4
    assert P(64, 48, 1, 64, 48, 1, 4)[1][0] == 6 and P(64, 48, 1, 64, 48, 15, 4)[1][0] == 4
    assert P(64, 48, 0, 64, 48, 23, 4)[1][0] == 7 and P(64, 48, 4, 64, 48, 23, 4)[1][0] == 4
    for f in (4, 4 | ACC, 1):
        assert P(64, 48, 0, 96, 80, 2, f)[1][0] == 4
    # chroma geometry: source sub-sampling by format, rgb sources at half width unless asked / forced otherwise (utils.c:1021-1034)
    assert P(65, 49, 6, 100, 70, 2, 4)[1][1:5] == [17, 13, 50, 70]
    assert P(65, 49, 31, 100, 70, 5, 4)[1][1:5] == [65, 25, 100, 70]
    assert P(64, 48, 2, 32, 24, 0, 4)[1][1:6] == [32, 48, 16, 12, 2]
    assert P(64, 48, 2, 256, 96, 0, 4)[1][1:3] == [64, 48] and P(64, 48, 2, 256, 96, 0, 4 | 0x4000)[1][1:3] == [64, 48]
    assert P(64, 48, 2, 256, 96, 0, 1)[1][1:3] == [32, 48]          # SWS_FAST_BILINEAR keeps the half-width reader
    assert P(64, 48, 23, 32, 24, 2, 4 | 0x2000)[1][1:6] == [32, 24, 32, 24, 1]
    assert P(64, 48, 12, 96, 80, 2, 4)[1][7] == 1 and P(64, 48, 0, 96, 80, 2, 4)[1][7] == 0
    assert [P(64, 48, 0, 96, 80, d, 4)[1][6] for d in (0, 62, 64, 47, 2, 26, 1)] == [8, 9, 10, 16, 3, 4, 2]
    # a yuv destination of the other range: the scaler even at the same size (no plane copy, utils.c:1043-1044); same range: still the copy
    assert P(64, 48, 12, 64, 48, 0, 4)[1][0] == 4 and P(64, 48, 0, 64, 48, 12, 4)[1][0] == 4 and P(64, 48, 12, 64, 48, 12, 4)[1][0] == 1 and P(64, 48, 13, 64, 48, 66, 4)[1][0] == 4


def test_refusals_carry_a_reason(built):
    import libav_b200._lib as L
    out = (C.c_int32 * 8)()
    cases = [((64, 48, 9, 64, 48, 2, 4), "sources taken over"), ((64, 48, 0, 64, 48, 33, 4), "destinations taken over"),
             ((64, 48, 0, 96, 80, 2, 4 | 0x10000), "CHR_DROP"), ((64, 48, 0, 96, 80, 27, 4 | 0x2000), "abgr"),
             ((64, 48, 26, 96, 80, 28, 4), "32-bit rgb source"), ((64, 48, 2, 64, 48, 25, 4), "writes past the row"), ((64, 49, 3, 64, 49, 0, 4), "even height"),
             ((64, 48, 12, 96, 80, 47, 4), "range conversion"), ((64, 48, 23, 64, 48, 24, 4), "nv12"), ((64, 48, 6, 64, 48, 0, 4), "yvu9ToYv12Wrapper"),
             ((2, 2, 0, 64, 48, 2, 4), ""), ((64, 48, 0, 96, 80, 2, 4 | 2), "")]
    for args, reason in cases:
        assert L.lib.sws_debug_plan_cuda(*args, out) == 0, args
        assert reason in L.last_error(), (args, L.last_error())
        L.lib.avb200_clear_error()
