"""Range conversion between full-range (yuvj) and limited-range planar yuv (lumRangeFromJpeg_c / chrRangeFromJpeg_c / lumRangeToJpeg_c /
chrRangeToJpeg_c between the horizontal and the vertical pass, swscale.c:166-197,748-765; handle_jpeg on both formats, utils.c:855-873;
no unscaled special converter when the ranges differ, utils.c:1043-1044).  CPU: port vs the compiled reference."""
import numpy as np
import pytest

from test_sws_planar_dst import ACC, run, source

J = {12: 0, 13: 4, 14: 5, 32: 31}
SRC_SUB = {0: (1, 1), 4: (1, 0), 5: (0, 0), 6: (2, 2), 7: (2, 0), 31: (0, 1)}
GEOMS = [(64, 48, 64, 48), (66, 50, 66, 50), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211)]
FLAGS = (4 | ACC, 2 | 0x80000, 0x10, 1 | ACC, 0x200 | ACC)
# (source, destination): full -> limited, limited -> full, same range, 9 / 10-bit destinations
PAIRS = [(12, 0), (13, 4), (14, 5), (32, 31), (12, 5), (14, 0), (13, 62), (12, 64), (14, 63), (0, 12), (4, 13), (5, 14), (31, 32), (5, 12), (6, 12), (7, 13),
         (12, 12), (13, 14), (14, 13), (32, 12)]
# ... and with a semi-planar / packed side (two-pass path: the range kernel sits between the passes whatever reads or writes the planes)
MIXED = [(12, 23), (13, 24), (14, 1), (12, 15), (13, 1), (23, 12), (24, 14), (1, 12), (15, 13), (2, 12), (3, 14), (26, 13)]


def planes(fmt, w, h, seed):
    """planar 8-bit source of any sub-sampling, rows padded like tests/test_sws_planar_dst.source"""
    from libav_b200 import synth
    hs, vs = SRC_SUB[J.get(fmt, fmt)]
    r = np.random.RandomState(seed)
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    return [synth.pad_rows(r.randint(0, 256, s).astype(np.uint8)) for s in ((h, w), (ch, cw), (ch, cw))]


def cases():
    for (sf, df) in PAIRS:
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS:
                yield sf, df, w, h, dw, dh, flags


def test_port_matches_reference(orc, refo):
    n = changed = 0
    for (sf, df, w, h, dw, dh, flags) in cases():
        pl = planes(sf, w, h, 11)
        a, b = run(refo, sf, pl, w, h, df, dw, dh, flags), run(orc, sf, pl, w, h, df, dw, dh, flags)
        assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), (sf, df, w, h, dw, dh, hex(flags))
        if (sf in J) != (df in J) and (w, h) == (dw, dh):        # same size but NOT a plane copy: the samples were rescaled
            changed += not np.array_equal(a[1][0][:, :dw], pl[0][:h, :w])
        n += 1
    assert n == len(PAIRS) * len(GEOMS) * len(FLAGS) and changed > 50


def mixed_planes(fmt, w, h, seed):
    from libav_b200 import synth
    if J.get(fmt, fmt) in SRC_SUB:
        return planes(fmt, w, h, seed)
    r = np.random.RandomState(seed)
    if fmt in (23, 24):
        return [synth.pad_rows(r.randint(0, 256, (h, w)).astype(np.uint8)), r.randint(0, 256, ((h + 1) // 2, 2 * ((w + 1) // 2) + 6)).astype(np.uint8)]
    bpp = {1: 2, 15: 2, 2: 3, 3: 3, 26: 4}[fmt]
    return [r.randint(0, 256, (h, bpp * w + 16)).astype(np.uint8)]


def mixed_outputs(df, dw, dh):
    if df in (23, 24):
        return [np.full((dh, dw + 3), 7, np.uint8), np.full(((dh + 1) // 2, 2 * ((dw + 1) // 2) + 6), 7, np.uint8)]
    if df in (1, 15):
        return [np.full((dh, 2 * dw + 10), 7, np.uint8)]
    from test_sws_planar_dst import outputs
    return outputs(df, dw, dh)


def mixed_run(o, sf, pl, w, h, df, dw, dh, flags):
    import ctypes as C
    out = mixed_outputs(df, dw, dh)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(sf, sp, ss, w, h, df, dp, ds, dw, dh, flags), out


def mixed_cases():
    for (sf, df) in MIXED:
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS:
                chr_dw = dw if J.get(df, df) == 5 else (dw + 1) // 2
                if (flags & 1) and sf in (23, 24, 1, 15, 2, 3, 26) and (dw > w or chr_dw > (w + 1) // 2):
                    continue            # fast bilinear up-scaling (luma or chroma) of a source that goes through the reference's uncleared formatConvBuffer
                yield sf, df, w, h, dw, dh, flags


def test_port_matches_reference_with_a_semi_planar_or_packed_side(orc, refo):
    n = 0
    for (sf, df, w, h, dw, dh, flags) in mixed_cases():
        pl = mixed_planes(sf, w, h, 13)
        a, b = mixed_run(refo, sf, pl, w, h, df, dw, dh, flags), mixed_run(orc, sf, pl, w, h, df, dw, dh, flags)
        assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), (sf, df, w, h, dw, dh, hex(flags))
        n += 1
    assert n > 230


def test_what_stays_refused(orc):
    pl = planes(12, 64, 48, 1)
    for df in (47, 49, 48):                       # 16-bit destinations of the other range (nv12 / packed 4:2:2: tests/test_sws_plan_cpu.py)
        assert run(orc, 12, pl, 64, 48, df, 96, 80, 4)[0] != 80, df
