"""Planar 8-bit YUV sources other than 4:2:0 (yuv422p, yuv444p, yuv410p, yuv411p, yuv440p) -> rgb24 / bgr24 / yuv420p: the
source's chroma sub-sampling only changes chrSrcW / chrSrcH (getSubSampleFactors, utils.c:983) -- and, for yuv422p, which
chroma line the unscaled table converter reads (yuv2rgb.c:133-136).  CPU: port vs the compiled reference; GPU: product vs
checker."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr

SUB = {0: (1, 1), 4: (1, 0), 5: (0, 0), 6: (2, 2), 7: (2, 0), 31: (0, 1)}
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (100, 37, 333, 211), (66, 50, 33, 25)]
ACC = 0x40000 | 0x80000


def frame(fmt, w, h, seed):
    hs, vs = SUB[fmt]
    r = np.random.RandomState(seed)
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    return [synth.pad_rows(r.randint(0, 256, s).astype(np.uint8)) for s in ((h, w), (ch, cw), (ch, cw))]


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags):
    out = [np.full((dh, dw * 3 + 6), 7, np.uint8)] if dfmt == 2 else \
          [np.full((dh, dw), 7, np.uint8), np.full(((dh + 1) // 2, (dw + 1) // 2), 7, np.uint8), np.full(((dh + 1) // 2, (dw + 1) // 2), 7, np.uint8)]
    sp, ss = (C.c_void_p * 3)(*[a.ctypes.data for a in pl]), (C.c_int * 3)(*[a.strides[0] for a in pl])
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


@pytest.mark.parametrize("fmt", [4, 5, 6, 7, 31])
def test_port_matches_reference(orc, refo, fmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for (w, h, dw, dh) in GEOMS:
        pl = frame(fmt, w, h, 3)
        for flags in (4 | ACC, 2 | 0x80000, 0x10 | ACC, 1 | ACC, 4 | ACC | 0x2000):
            for dfmt in (2, 0):
                if dfmt == 0 and flags & 0x2000:
                    continue
                a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags)
                assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), (fmt, w, h, dw, dh, hex(flags), dfmt)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [4, 5, 6, 7, 31])
def test_gpu_matches_checker(gpu, checker, fmt):
    from libav_b200 import device
    for (w, h, dw, dh) in GEOMS + [(1920, 1080, 1280, 720)]:
        pl = frame(fmt, w, h, 5)
        for flags in (4 | ACC, 2 | 0x80000, 4 | ACC | 0x2000):
            if w * h > 10 ** 6 and flags != (4 | ACC):
                continue
            for dfmt in (device.PIX_FMT_RGB24, device.PIX_FMT_BGR24, device.PIX_FMT_YUV420P):
                if dfmt == device.PIX_FMT_YUV420P and flags & 0x2000:
                    continue
                rc, want = run(checker, fmt, pl, w, h, 2 if dfmt != device.PIX_FMT_YUV420P else 0, dw, dh, flags)
                assert rc == dh
                ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
                if dfmt == device.PIX_FMT_YUV420P:
                    for a, b in zip(ctx.scale(pl), want):
                        assert np.array_equal(a, b), (fmt, w, h, dw, dh, hex(flags), "yuv")
                else:
                    got = ctx.scale(pl, dst_pad=6)[:, :dw * 3].reshape(dh, dw, 3)
                    wr = want[0][:, :dw * 3].reshape(dh, dw, 3)
                    if dfmt == device.PIX_FMT_BGR24:
                        got = got[:, :, ::-1]
                    assert np.array_equal(got, wr), (fmt, w, h, dw, dh, hex(flags), dfmt)
                ctx.close()


@pytest.mark.gpu
def test_unscaled_special_converter_is_refused_loudly(gpu):
    from libav_b200 import device
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 64, 48, device.PIX_FMT_YUV420P, 2, src_fmt=6)     # yvu9ToYv12Wrapper territory
    gpu.lib.avb200_clear_error()
