"""9 / 10 / 16-bit planar yuv sources (hScale16To15_c, swscale.c:110-131; big-endian twins byte-swapped by the input stage; ordered dither on
8-bit planar outputs, swscale.c:389-390,553-556): port vs the compiled reference.  The product side runs in tests/test_hostsim_sws_frames_cpu.py
(host-compiled scaler) and tests/test_zz_gpu_late_slots.py."""
import ctypes as C

import numpy as np
import pytest

ACC = 0x40000 | 0x80000
# format -> (log2 chroma w, log2 chroma h, bits, big endian)
HBD = {62: (1, 1, 9, 0), 61: (1, 1, 9, 1), 64: (1, 1, 10, 0), 63: (1, 1, 10, 1), 47: (1, 1, 16, 0), 48: (1, 1, 16, 1),
       66: (1, 0, 10, 0), 65: (1, 0, 10, 1), 72: (1, 0, 9, 0), 49: (1, 0, 16, 0), 70: (0, 0, 10, 0), 69: (0, 0, 10, 1), 68: (0, 0, 9, 0), 51: (0, 0, 16, 0)}
DSTS = [2, 3, 28, 25, 0, 4, 5, 62, 64, 63, 23, 24, 1, 15]
GEOMS = [(64, 48, 96, 80), (96, 80, 64, 48), (66, 50, 66, 50), (101, 37, 64, 48)]
FLAGS = (4 | ACC, 2, 0x10, 1 | ACC, 4 | ACC | 0x2000)


def planes(fmt, w, h, seed):
    hs, vs, bits, be = HBD[fmt]
    r = np.random.RandomState(seed + fmt)
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    dt = np.dtype(">u2" if be else "<u2")
    out = []
    for (pw, ph) in ((w, h), (cw, ch), (cw, ch)):
        a = np.zeros((ph, pw + 9), dt)
        a[:, :pw] = r.randint(0, 1 << bits, (ph, pw))
        a[:, pw:] = a[:, pw - 1:pw]
        out.append(a)
    return out


def outputs(df, dw, dh):
    from libav_b200.device import PLANAR_BE, PLANAR_FORMATS
    if df in (23, 24):
        return [np.full((dh, dw + 3), 7, np.uint8), np.full(((dh + 1) // 2, 2 * ((dw + 1) // 2) + 6), 7, np.uint8)]
    if df in PLANAR_FORMATS:
        hs, vs, bits = PLANAR_FORMATS[df]
        dt = np.uint8 if bits == 8 else np.dtype(">u2" if df in PLANAR_BE else "<u2")
        cw, ch = -((-dw) >> hs), -((-dh) >> vs)
        return [np.full((dh, dw + 3), 7, dt), np.full((ch, cw + 3), 7, dt), np.full((ch, cw + 3), 7, dt)]
    bpp = 4 if 25 <= df <= 28 else 2 if df in (1, 15) else 3
    return [np.full((dh, dw * bpp + 10), 7, np.uint8)]


def run(o, sf, pl, w, h, df, dw, dh, flags):
    out = outputs(df, dw, dh)
    sp = (C.c_void_p * 3)(*[a.ctypes.data for a in pl])
    ss = (C.c_int * 3)(*[a.strides[0] for a in pl])
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(sf, sp, ss, w, h, df, dp, ds, dw, dh, flags), out


def taken(sf, df, w, h, dw, dh, flags):
    """what port and product take over for these sources: no planarCopyWrapper cases (same size and sub-sampling, planar destination)"""
    from libav_b200.device import PLANAR_FORMATS
    if (flags & 0x2000) and df not in (2, 3, 28, 25):
        return False
    if df in PLANAR_FORMATS and (w, h) == (dw, dh) and PLANAR_FORMATS[df][:2] == HBD[sf][:2]:
        return False
    return True


def cases(formats=tuple(HBD)):
    for sf in formats:
        for df in DSTS:
            for (w, h, dw, dh) in GEOMS:
                for flags in FLAGS:
                    if taken(sf, df, w, h, dw, dh, flags):
                        yield sf, df, w, h, dw, dh, flags


@pytest.mark.parametrize("sf", list(HBD))
def test_port_matches_reference(orc, refo, sf):
    n = 0
    for (s, df, w, h, dw, dh, flags) in cases((sf,)):
        pl = planes(s, w, h, 5)
        a, b = run(refo, s, pl, w, h, df, dw, dh, flags), run(orc, s, pl, w, h, df, dw, dh, flags)
        assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), (s, df, w, h, dw, dh, hex(flags), a[0], b[0])
        n += 1
    assert n > 150


def test_refusals(orc):
    pl = planes(64, 64, 48, 1)
    assert run(orc, 64, pl, 64, 48, 64, 64, 48, 4)[0] != 48          # yuv420p10 -> yuv420p10 of the same size: planarCopyWrapper
    assert run(orc, 64, pl, 64, 48, 0, 64, 48, 4)[0] != 48           # -> yuv420p: planarCopyWrapper's dithered depth conversion
    assert run(orc, 64, pl, 64, 48, 47, 96, 80, 4)[0] != 80          # 16-bit destination: the 19-bit line functions are not restated for these sources
