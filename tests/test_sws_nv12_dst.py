"""nv12 / nv21 destinations: yuv2nv12cX_c (output.c:267-303) = the planar chroma recipe written interleaved; yuv420p of the same
size through planarToNv12Wrapper (swscale_unscaled.c:138-156).  CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

from test_sws_planar_dst import source, undefined_edge

SRC = [0, 4, 5, 6, 23, 24, 1, 15, 2, 3, 26]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 51, 67, 51), (66, 50, 33, 25)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 4, 0x10 | ACC, 1 | ACC)


def src_planes(fmt, w, h, seed):
    if fmt in (3, 26, 15, 24, 6):
        r = np.random.RandomState(seed)
        if fmt == 6:
            from libav_b200 import synth
            cw, ch = -((-w) >> 2), -((-h) >> 2)
            return [synth.pad_rows(r.randint(0, 256, s).astype(np.uint8)) for s in ((h, w), (ch, cw), (ch, cw))]
        if fmt == 24:
            return source(23, w, h, seed)
        bpp = {3: 3, 26: 4, 15: 2}[fmt]
        return [r.randint(0, 256, (h, bpp * w + 12)).astype(np.uint8)]
    return source(fmt, w, h, seed)


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags):
    cw, ch = (dw + 1) // 2, (dh + 1) // 2
    out = [np.full((dh, dw + 3), 7, np.uint8), np.full((ch, 2 * cw + 4), 7, np.uint8)]
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(out[0].ctypes.data, out[1].ctypes.data, None), (C.c_int * 3)(out[0].strides[0], out[1].strides[0], 0)
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


def combos():
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS:
            if fmt in (23, 24) and (w, h) == (dw, dh):
                continue             # refused: the reference's plane copy skips the chroma plane for nv12 -> nv12
            if fmt == 6 and (w, h) == (dw, dh):
                continue
            for flags in FLAGS:
                if flags & 1 and fmt in (23, 24, 1, 15, 2, 3, 26) and (dw > w or (dw + 1) // 2 > (w + 1) // 2):
                    continue         # undefined right edge, see tests/test_sws_packed_sources.py
                yield fmt, w, h, dw, dh, flags


@pytest.mark.parametrize("dfmt", [23, 24])
def test_port_matches_reference(orc, refo, dfmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for (fmt, w, h, dw, dh, flags) in combos():
        pl = src_planes(fmt, w, h, 3)
        a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), (fmt, dfmt, w, h, dw, dh, hex(flags))


@pytest.mark.gpu
@pytest.mark.parametrize("dfmt", [23, 24])
def test_gpu_matches_checker(gpu, checker, dfmt):
    from libav_b200 import device
    for (fmt, w, h, dw, dh, flags) in list(combos()) + [(0, 1920, 1080, 1280, 720, 4 | ACC), (0, 640, 360, 1920, 1080, 4 | ACC)]:
        pl = src_planes(fmt, w, h, 5)
        rc, want = run(checker, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
        got = ctx.scale(pl, fill=7)
        for a, b in zip(got, want):
            assert np.array_equal(a, b[:, :a.shape[1]]), (fmt, dfmt, w, h, dw, dh, hex(flags))
        ctx.close()


@pytest.mark.gpu
def test_refusal(gpu):
    from libav_b200 import device
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 64, 48, 23, 4, src_fmt=24)
    gpu.lib.avb200_clear_error()
