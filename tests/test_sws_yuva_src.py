"""yuva420p (AV_PIX_FMT_YUVA420P = 33) sources: the alpha plane is read only when the destination has alpha as well (c->alpPixBuf,
libswscale/utils.c:1244; needAlpha, yuv2rgb.c:870); for every other destination the reference treats the format like yuv420p wherever it
tests formats (swscale_unscaled.c:1041-1153).  Product and port take those over (three planes, src[3] untouched) and refuse argb / rgba /
abgr / bgra.  CPU: reference(33) == reference(0) == port(33); GPU: product(33) == checker."""
import ctypes as C

import numpy as np
import pytest

import test_sws_rgb48_dst as R
from test_sws_planar_dst import PLANAR_FORMATS

ACC = 0x40000 | 0x80000
DSTS = [2, 3, 0, 4, 5, 23, 1, 37, 35, 8, 64, 12]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (101, 37, 333, 211), (66, 50, 33, 25)]
FLAGS = (4 | ACC, 4, 0x10, 1 | ACC, 2)


def dest(df, w, h):
    if df == 23:
        return [np.zeros((h, w + 8), np.uint8), np.zeros(((h + 1) // 2, 2 * ((w + 1) // 2) + 8), np.uint8)]
    if df in PLANAR_FORMATS:
        hs, vs, bits = PLANAR_FORMATS[df]
        cw, ch = -((-w) >> hs), -((-h) >> vs)
        return [np.zeros((h, w + 8), np.uint8 if bits == 8 else np.uint16) for (h, w) in ((h, w), (ch, cw), (ch, cw))]
    if df == 12:
        return [np.zeros((h, w + 8), np.uint8), np.zeros(((h + 1) // 2, (w + 1) // 2 + 8), np.uint8), np.zeros(((h + 1) // 2, (w + 1) // 2 + 8), np.uint8)]
    return [np.zeros((h, w * {8: 1, 1: 2, 15: 2, 37: 2, 35: 6}.get(df, 3) + 16), np.uint8)]


def run(o, sf, pl, sw, sh, df, dw, dh, flags):
    out = dest(df, dw, dh)
    sp = (C.c_void_p * 3)(*[a.ctypes.data for a in pl])
    ss = (C.c_int * 3)(*[a.strides[0] for a in pl])
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(sf, sp, ss, sw, sh, df, dp, ds, dw, dh, flags), out


def crop(df, dw, planes):
    """the pictures only, not the row padding (the reference's plane copies move whole strides when the pitches agree; port and reference differ
    in the room they need to write the pair of an odd last column of packed rgb)"""
    if len(planes) == 1:
        return [planes[0][:, :dw * {8: 1, 1: 2, 15: 2, 37: 2, 35: 6}.get(df, 3)]]
    if df == 23:
        return [planes[0][:, :dw], planes[1][:, :2 * ((dw + 1) // 2)]]
    hs = 1 if df == 12 else PLANAR_FORMATS[df][0]
    return [planes[0][:, :dw]] + [p[:, :-((-dw) >> hs)] for p in planes[1:]]


def test_yuva420p_is_yuv420p_without_alpha_destinations(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for df in DSTS:
        for (sw, sh, dw, dh) in GEOMS:
            for flags in FLAGS:
                pl = R.source(0, sw, sh, 5)
                (ra, a), (rb, b), (rc, c) = run(refo, 33, pl, sw, sh, df, dw, dh, flags), run(refo, 0, pl, sw, sh, df, dw, dh, flags), run(orc, 33, pl, sw, sh, df, dw, dh, flags)
                assert ra == rb == dh and all(np.array_equal(x, y) for x, y in zip(a, b)), (df, sw, sh, dw, dh, hex(flags))
                if rc == -1:
                    assert run(orc, 0, pl, sw, sh, df, dw, dh, flags)[0] == -1           # refused for yuv420p as well
                    continue
                assert rc == dh and all(np.array_equal(x, y) for x, y in zip(crop(df, dw, a), crop(df, dw, c))), (df, sw, sh, dw, dh, hex(flags))
                n += 1
    assert n > 200
    pl = R.source(0, 64, 48, 5)
    for df in (25, 26, 27, 28):
        assert run(orc, 33, pl, 64, 48, df, 128, 96, 4 | ACC)[0] == -1


@pytest.mark.gpu
def test_gpu_yuva420p_sources(gpu, checker):
    from libav_b200 import device
    for df in (2, 3, 35, 37):
        for (sw, sh, dw, dh) in GEOMS:
            for flags in (4 | ACC, 4):
                if df == 37 and (sw, sh) == (dw, dh) and not flags & 0x40000:
                    continue                     # (the ordered-dither table converter: refused for yuv420p too)
                pl = R.source(0, sw, sh, 7)
                bpp = {35: 6, 37: 2}.get(df, 3)
                rc, want = run(checker, 0, pl, sw, sh, df, dw, dh, flags)
                assert rc == dh
                ctx = device.SwsContext(sw, sh, dw, dh, df, flags, src_fmt=33)
                got = ctx.scale(pl, dst_pad=16, fill=0)
                assert np.array_equal(got[:, :bpp * dw], want[0][:, :bpp * dw]), (df, sw, sh, dw, dh, hex(flags))
                ctx.close()
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 128, 96, 26, 4 | ACC, src_fmt=33)
    gpu.lib.avb200_clear_error()
