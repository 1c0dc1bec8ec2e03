"""sws_setColorspaceDetails for packed rgb destinations (BT.709 / full range / brightness, contrast, saturation: ff_yuv2rgb_c_init_tables,
yuv2rgb.c:671-863) and the full-range yuvj source formats (handle_jpeg, utils.c:855-873).
CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

from test_sws_planar_dst import source

ITU709 = (117504, 138453, 13954, 34903)
FCC = (104448, 132798, 24759, 53109)
SMPTE240M = (117579, 136230, 16907, 35559)
ITU601 = (104597, 132201, 25675, 53279)
SETTINGS = [(ITU709, 0, 0, 1 << 16, 1 << 16), (ITU709, 1, 0, 1 << 16, 1 << 16), (ITU601, 1, 0, 1 << 16, 1 << 16), (FCC, 0, 3000, 70000, 60000),
            (SMPTE240M, 1, -4000, 60000, 80000), (ITU601, 0, 0, 1 << 16, 0)]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (101, 37, 333, 211), (640, 480, 352, 288)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 4, 2 | ACC | 0x2000, 1 | ACC)


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags, cs):
    bpp = 4 if dfmt >= 25 else 3
    out = np.full((dh, dw * bpp + 8), 7, np.uint8)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    if cs is not None:
        o.sws_set_colorspace((C.c_int * 4)(*cs[0]), cs[1], cs[2], cs[3], cs[4])
    try:
        return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out
    finally:
        o.sws_set_colorspace(None, 0, 0, 0, 0)


def combos():
    for cs in SETTINGS:
        for fmt in (0, 4, 23, 1):
            for (w, h, dw, dh) in GEOMS:
                for flags in FLAGS:
                    if flags & 1 and fmt in (23, 1) and dw > w:
                        continue
                    yield cs, fmt, w, h, dw, dh, flags, 2 if (w + flags) % 3 else 28
    for fmt in (12, 13, 14):                    # yuvj: full range without any call
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS[:3]:
                yield None, fmt, w, h, dw, dh, flags, 2


def planes(fmt, w, h, seed):
    return source({12: 0, 13: 4, 14: 5}.get(fmt, fmt), w, h, seed)


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for (cs, fmt, w, h, dw, dh, flags, dfmt) in combos():
        pl = planes(fmt, w, h, 3)
        a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags, cs), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags, cs)
        assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), (cs, fmt, w, h, dw, dh, hex(flags), dfmt, np.argwhere(a[1] != b[1])[:4].tolist())
        n += 1
    assert n > 300
    # the settings change the picture
    pl = planes(0, 64, 48, 3)
    assert not np.array_equal(run(orc, 0, pl, 64, 48, 2, 64, 48, 4 | ACC, SETTINGS[0])[1], run(orc, 0, pl, 64, 48, 2, 64, 48, 4 | ACC, None)[1])


@pytest.mark.gpu
def test_gpu_matches_checker(gpu, checker):
    from libav_b200 import device
    for (cs, fmt, w, h, dw, dh, flags, dfmt) in list(combos()) + [(SETTINGS[0], 0, 1920, 1080, 1920, 1080, 4 | ACC, 2), (SETTINGS[1], 0, 1280, 720, 1920, 1080, 4 | ACC, 2)]:
        pl = planes(fmt, w, h, 5)
        rc, want = run(checker, fmt, pl, w, h, dfmt, dw, dh, flags, cs)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
        if cs is not None:
            tab = (C.c_int * 4)(*cs[0])
            assert gpu.lib.sws_setColorspaceDetails_cuda(ctx.ctx, tab, cs[1], tab, 0, cs[2], cs[3], cs[4]) == 0
        got = ctx.scale(pl, dst_pad=8, fill=7)
        assert np.array_equal(got, want), (cs, fmt, w, h, dw, dh, hex(flags), dfmt, np.argwhere(got != want)[:4].tolist())
        ctx.close()


@pytest.mark.gpu
def test_gpu_refusals(gpu):
    from libav_b200 import device
    tab = (C.c_int * 4)(*ITU709)
    ctx = device.SwsContext(64, 48, 128, 96, 0, 4)                       # yuv destination: -1 like the reference
    assert gpu.lib.sws_setColorspaceDetails_cuda(ctx.ctx, tab, 0, tab, 0, 0, 1 << 16, 1 << 16) == -1
    ctx.close()
    ctx = device.SwsContext(64, 48, 128, 96, 2, 4)
    assert gpu.lib.sws_setColorspaceDetails_cuda(ctx.ctx, tab, 0, tab, 0, 0, 1 << 16, 1 << 20) == -1      # saturation 16: leaves the colour table
    assert "colour table" in gpu.last_error()
    gpu.lib.avb200_clear_error()
    ctx.close()
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 128, 96, 47, 4, src_fmt=12)             # yuvj420p -> 16-bit yuv: the 16-bit range conversion is not taken over
    gpu.lib.avb200_clear_error()
