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


def test_what_stays_refused(orc):
    pl = planes(12, 64, 48, 1)
    for df in (47, 49, 48):                       # 16-bit destinations of the other range (nv12 / packed 4:2:2: tests/test_sws_plan_cpu.py)
        assert run(orc, 12, pl, 64, 48, df, 96, 80, 4)[0] != 80, df
