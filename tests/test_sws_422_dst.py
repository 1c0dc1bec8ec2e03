"""Packed 4:2:2 destinations (yuyv422 1, uyvy422 15): yuv2422_X / _2 / _1 (output.c:448-576) -- the packed output stage without the
colour conversion -- and the reference's unscaled converters (yuv422p always, yuv420p with fast-bilinear / point, same-format copy).
CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

from test_sws_nv12_dst import src_planes

SRC = [0, 4, 5, 23, 1, 15, 2, 26]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 51, 67, 51), (66, 50, 33, 25), (64, 48, 64, 96)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 4, 0x10 | ACC, 1 | ACC, 4 | ACC | 0x2000)


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags):
    out = np.full((dh, dw * 2 + 6), 7, np.uint8)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


def combos():
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS:
                if flags & 1 and fmt in (23, 1, 15, 2, 26) and dw > w:
                    continue         # undefined right edge, see tests/test_sws_packed_sources.py
                yield fmt, w, h, dw, dh, flags


@pytest.mark.parametrize("dfmt", [1, 15])
def test_port_matches_reference(orc, refo, dfmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for (fmt, w, h, dw, dh, flags) in combos():
        pl = src_planes(fmt, w, h, 3)
        a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), (fmt, dfmt, w, h, dw, dh, hex(flags), np.argwhere(a[1] != b[1])[:4].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("dfmt", [1, 15])
def test_gpu_matches_checker(gpu, checker, dfmt):
    from libav_b200 import device
    for (fmt, w, h, dw, dh, flags) in list(combos()) + [(0, 1920, 1080, 1280, 720, 4 | ACC)]:
        pl = src_planes(fmt, w, h, 5)
        rc, want = run(checker, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
        got = ctx.scale(pl, dst_pad=6, fill=7)
        assert np.array_equal(got, want), (fmt, dfmt, w, h, dw, dh, hex(flags), np.argwhere(got != want)[:4].tolist())
        ctx.close()
