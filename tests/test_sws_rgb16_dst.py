"""15 / 16 / 12-bpp packed rgb destinations (rgb565 / bgr565 / rgb555 / bgr555 / rgb444 / bgr444, LE and BE; libavutil/pixfmt.h 36-43,
54-57): yuv2rgb_write's branch that adds a 2 x 2 (4 x 4) ordered dither to the table index and sums three pre-shifted channel fields
(output.c:869-902, tables yuv2rgb.c:806-844), behind the X / 2 / 1 selection of swscale().
CPU: port vs the compiled reference; GPU: product vs checker (host-pointer and batched device-pointer calls)."""
import ctypes as C

import numpy as np
import pytest

from test_sws_planar_dst import source

DST = [37, 36, 41, 40, 39, 38, 43, 42, 54, 55, 56, 57]
SRC = [0, 4, 5, 23]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 50, 67, 50), (66, 50, 33, 25), (64, 48, 64, 96)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 4, 0x10 | ACC, 4 | ACC | 0x2000, 1 | ACC)


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags):
    out = np.full((dh, dw * 2 + 8), 7, np.uint8)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


def table_converter(fmt, w, h, dw, dh, flags):
    """the reference's ordered-dither unscaled converters (yuv2rgb.c:377-573): same size, 4:2:0 / 4:2:2 planar source, no SWS_ACCURATE_RND, even height"""
    return fmt in (0, 4) and (w, h) == (dw, dh) and not flags & 0x40000 and not dh & 1


def combos(dfmt):
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS:
                if flags & 1 and fmt == 23 and dw > w:
                    continue         # undefined right edge, see tests/test_sws_packed_sources.py
                if table_converter(fmt, w, h, dw, dh, flags):
                    continue         # refused (checked below)
                yield fmt, w, h, dw, dh, flags


@pytest.mark.parametrize("dfmt", DST)
def test_port_matches_reference(orc, refo, dfmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for (fmt, w, h, dw, dh, flags) in combos(dfmt):
        pl = source(fmt, w, h, 3)
        a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), (fmt, dfmt, w, h, dw, dh, hex(flags), np.argwhere(a[1] != b[1])[:4].tolist())
        n += 1
    assert n > 100
    pl = source(0, 64, 48, 3)
    assert run(orc, 0, pl, 64, 48, dfmt, 64, 48, 4)[0] == -1           # the table converter's case is not restated


def test_dither_is_exercised(orc):
    """the 565 output is not the 24-bit output cut down: the dither reaches the low bits"""
    pl = source(0, 64, 48, 9)
    rc, o16 = run(orc, 0, pl, 64, 48, 37, 128, 96, 4 | ACC)
    assert rc == 96
    px = o16[:, :256].view(np.uint16)
    assert len(np.unique(px)) > 500 and (px[0::2] != px[1::2]).any()


@pytest.mark.gpu
@pytest.mark.parametrize("dfmt", DST)
def test_gpu_matches_checker(gpu, checker, dfmt):
    from libav_b200 import device
    for (fmt, w, h, dw, dh, flags) in combos(dfmt):
        pl = source(fmt, w, h, 5)
        rc, want = run(checker, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
        got = ctx.scale(pl, dst_pad=8, fill=7)
        assert np.array_equal(got, want), (fmt, dfmt, w, h, dw, dh, hex(flags), np.argwhere(got != want)[:4].tolist())
        ctx.close()


@pytest.mark.gpu
def test_gpu_device_batch_and_refusals(gpu, checker):
    from libav_b200 import device
    w, h, dw, dh, dfmt = 640, 480, 1280, 720, 37
    frames = [source(0, w, h, 30 + k) for k in range(3)]
    tight = [[np.ascontiguousarray(p[:, :p.shape[1]]) for p in f] for f in frames]
    src = [device.DevBuf.from_numpy(np.stack([f[i] for f in tight])) for i in range(3)]
    dst = device.DevBuf(3 * dh * dw * 2)
    ctx = device.SwsContext(w, h, dw, dh, dfmt, 4 | ACC)
    ctx.scale_device([s.ptr for s in src], [f.strides[0] for f in tight[0]], [dst.ptr], [dw * 2], nframes=3,
                     src_frame=[f.nbytes for f in tight[0]], dst_frame=[dh * dw * 2])
    got = dst.download(np.uint8, (3, dh, dw * 2))
    device.sync()
    for k in range(3):
        rc, want = run(checker, 0, frames[k], w, h, dfmt, dw, dh, 4 | ACC)
        assert np.array_equal(got[k], want[:, :dw * 2]), k
    ctx.close()
    for args in ((64, 48, 64, 48, 37, 4), (64, 48, 128, 96, 37, 4 | ACC, 2)):      # the table converter's case; a packed rgb source
        with pytest.raises(Exception):
            device.SwsContext(*args[:6], **({"src_fmt": args[6]} if len(args) > 6 else {}))
        gpu.lib.avb200_clear_error()
