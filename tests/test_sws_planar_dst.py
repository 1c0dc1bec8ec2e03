"""Planar yuv destinations other than yuv420p: 8-bit 422p / 444p / 410p / 411p / 440p and the little-endian 9 / 10 / 16-bit 420p /
422p / 444p (yuv2planeX_10_c / yuv2plane1_10_c, output.c:183-213; yuv2planeX_16_c on hScale8To19_c lines, :136-172; planarCopyWrapper's
8 -> 9 / 10 / 16 bit conversions, swscale_unscaled.c:946-992; yuyvtoyuv422_c / uyvytoyuv422_c).  CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from libav_b200.device import PLANAR_BE, PLANAR_FORMATS

SRC = {0: (1, 1), 4: (1, 0), 5: (0, 0), 23: None, 1: None, 2: None, 15: None}
DST = [4, 5, 6, 7, 31, 62, 64, 66, 68, 70, 72, 47, 49, 51, 61, 65, 69, 48, 52]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (66, 50, 33, 25)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 0x10 | ACC, 0x200 | ACC, 1 | ACC)


def source(fmt, w, h, seed):
    r = np.random.RandomState(seed)
    if fmt in (1, 15):
        return [r.randint(0, 256, (h, 2 * w + 10)).astype(np.uint8)]
    if fmt == 2:
        return [r.randint(0, 256, (h, 3 * w + 10)).astype(np.uint8)]
    if fmt == 23:
        cw, ch = (w + 1) // 2, (h + 1) // 2
        return [synth.pad_rows(r.randint(0, 256, (h, w)).astype(np.uint8)), r.randint(0, 256, (ch, 2 * cw + 6)).astype(np.uint8)]
    hs, vs = SRC[fmt]
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    return [synth.pad_rows(r.randint(0, 256, s).astype(np.uint8)) for s in ((h, w), (ch, cw), (ch, cw))]


def outputs(dfmt, dw, dh, pad=3):
    hs, vs, bits = PLANAR_FORMATS[dfmt]
    dt = np.uint8 if bits == 8 else np.dtype(">u2" if dfmt in PLANAR_BE else "<u2")
    cw, ch = -((-dw) >> hs), -((-dh) >> vs)
    return [np.full((dh, dw + pad), 7, dt), np.full((ch, cw + pad), 7, dt), np.full((ch, cw + pad), 7, dt)]


def undefined_edge(fmt, flags, sw, dw, dfmt):
    """fast-bilinear up-scaling of a source that goes through the reference's uncleared formatConvBuffer: its last pixels
    depend on uninitialised memory (tests/test_sws_packed_sources.py covers everything but that edge)"""
    chroma_up = -((-dw) >> PLANAR_FORMATS[dfmt][0]) > (sw + 1) // 2          # all four have half-width chroma lines
    return bool(flags & 1) and fmt in (23, 1, 2, 15) and (dw > sw or chroma_up) and PLANAR_FORMATS[dfmt][2] != 16


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags):
    out = outputs(dfmt, dw, dh)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(*[a.ctypes.data for a in out]), (C.c_int * 3)(*[a.strides[0] for a in out])
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


@pytest.mark.parametrize("dfmt", DST)
def test_port_matches_reference(orc, refo, dfmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS:
            pl = source(fmt, w, h, 3)
            for flags in FLAGS:
                if undefined_edge(fmt, flags, w, dw, dfmt):
                    continue
                a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags)
                assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), (fmt, dfmt, w, h, dw, dh, hex(flags))


@pytest.mark.gpu
@pytest.mark.parametrize("dfmt", DST)
def test_gpu_matches_checker(gpu, checker, dfmt):
    from libav_b200 import device
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS + [(1280, 720, 1920, 1080)]:
            pl = source(fmt, w, h, 5)
            for flags in FLAGS:
                if (w * h > 500000 and flags != 4 | ACC) or undefined_edge(fmt, flags, w, dw, dfmt):
                    continue
                rc, want = run(checker, fmt, pl, w, h, dfmt, dw, dh, flags)
                assert rc == dh
                ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
                got = ctx.scale(pl, fill=7)
                for a, b in zip(got, want):
                    assert np.array_equal(a, b[:, :a.shape[1]]), (fmt, dfmt, w, h, dw, dh, hex(flags))
                ctx.close()


@pytest.mark.gpu
def test_refusals_are_loud(gpu):
    from libav_b200 import device
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 64, 48, 33, 4)          # yuva420p
    gpu.lib.avb200_clear_error()
