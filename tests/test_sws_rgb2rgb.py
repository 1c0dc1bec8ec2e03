"""Same-size conversions between the byte-wise packed rgb formats with 32 bits on at least one side (rgb24 2, bgr24 3, argb 25, rgba 26, abgr 27,
bgra 28): the reference's rgbToRgbWrapper (libswscale/swscale_unscaled.c:590-705) with the converters of rgb2rgb.c:139-175,335-352 and
rgb2rgb_template.c:31-78,338-350 -- channel remaps (alpha copied 32 -> 32, 255 for 24 -> 32, dropped 32 -> 24).  A 24-bit source to argb / abgr is
refused: the reference runs its 4-byte writer one byte into the row there (ALT32_CORR) and writes past it.
Pictures are compared, not the row padding (the reference converts that too when the pitches are proportional, :694-697).
CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

FMTS = [2, 3, 25, 26, 27, 28]
BPP = {2: 3, 3: 3, 25: 4, 26: 4, 27: 4, 28: 4}
GEOMS = [(64, 48), (101, 37), (9, 2), (1920, 8)]


def pairs():
    for sf in FMTS:
        for df in FMTS:
            if BPP[sf] == 3 and BPP[df] == 3:
                continue                         # 24 <-> 24: tests/test_sws_packed_sources.py
            yield sf, df, BPP[sf] == 3 and df in (25, 27)


def picture(fmt, w, h, seed, pad):
    return np.random.RandomState(seed).randint(0, 256, (h, BPP[fmt] * w + pad)).astype(np.uint8)


def run(o, sf, src, w, h, df, flags, pad):
    out = np.full((h, BPP[df] * w + pad), 7, np.uint8)
    sp, ss = (C.c_void_p * 3)(src.ctypes.data, None, None), (C.c_int * 3)(src.strides[0], 0, 0)
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    return o.sws_planar(sf, sp, ss, w, h, df, dp, ds, w, h, flags), out


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for sf, df, refused in pairs():
        for (w, h) in GEOMS:
            for flags in (4, 4 | 0x40000 | 0x80000, 0x10):
                for spad, dpad in ((0, 0), (5, 9)):
                    src = picture(sf, w, h, 3 + w, spad)
                    rb, b = run(orc, sf, src, w, h, df, flags, dpad)
                    if refused:
                        assert rb == -1
                        continue
                    ra, a = run(refo, sf, src, w, h, df, flags, dpad)
                    assert ra == rb == h and np.array_equal(a[:, :BPP[df] * w], b[:, :BPP[df] * w]), (sf, df, w, h, hex(flags), spad, dpad)
                    assert (b[:, BPP[df] * w:] == 7).all()                      # the port writes the picture only
                    n += 1
    assert n > 500


def test_alpha_rules(orc):
    src = picture(2, 16, 4, 1, 0)
    rc, o = run(orc, 2, src, 16, 4, 28, 4, 0)                                      # rgb24 -> bgra: alpha 255, channels swapped
    px, sp = o.reshape(4, 16, 4), src.reshape(4, 16, 3)
    assert rc == 4 and (px[..., 3] == 255).all() and np.array_equal(px[..., 2], sp[..., 0]) and np.array_equal(px[..., 0], sp[..., 2])
    src = picture(25, 16, 4, 2, 0)
    rc, o = run(orc, 25, src, 16, 4, 28, 4, 0)                                     # argb -> bgra: alpha copied
    assert rc == 4 and np.array_equal(o.reshape(4, 16, 4)[..., 3], src.reshape(4, 16, 4)[..., 0])


@pytest.mark.gpu
def test_gpu_matches_checker(gpu, checker):
    from libav_b200 import device
    n = 0
    for sf, df, refused in pairs():
        for (w, h) in GEOMS:
            if refused:
                with pytest.raises(Exception):
                    device.SwsContext(w, h, w, h, df, 4, src_fmt=sf)
                gpu.lib.avb200_clear_error()
                continue
            src = picture(sf, w, h, 11 + w, 6)
            rc, want = run(checker, sf, src, w, h, df, 4, 8)
            assert rc == h
            ctx = device.SwsContext(w, h, w, h, df, 4, src_fmt=sf)
            got = ctx.scale([src], dst_pad=8, fill=7)
            assert np.array_equal(got[:, :BPP[df] * w], want[:, :BPP[df] * w]), (sf, df, w, h)
            ctx.close()
            n += 1
    assert n > 100
