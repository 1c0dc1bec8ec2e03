"""32-bit packed rgb destinations (argb 25, rgba 26, abgr 27, bgra 28): the 32-bit output functions and colour tables hold the
24-bit channel values plus a constant alpha of 255 (yuv2rgb.c:763-800, output.c yuv2rgb_write / yuv2rgb_full_X_c).
CPU: port vs the compiled reference; GPU: product vs checker (host-pointer and batched device-pointer calls)."""
import ctypes as C

import numpy as np
import pytest

from test_sws_planar_dst import source, undefined_edge

SRC = [0, 4, 5, 23, 1, 2]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 50, 67, 50), (66, 50, 33, 25)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 4, 0x10 | ACC, 4 | ACC | 0x2000, 1 | ACC)


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags):
    out = np.full((dh, dw * 4 + 8), 7, np.uint8)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


def combos(dfmt):
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS:
            if fmt == 2 and (w, h) == (dw, dh):
                continue             # rgb24 -> 32-bit rgb of the same size: the reference's rgb2rgb converters, refused
            for flags in FLAGS:
                if flags & 1 and fmt in (23, 1, 2) and dw > w:
                    continue         # undefined right edge, see tests/test_sws_packed_sources.py
                if dfmt == 27 and flags & 0x2000:
                    continue         # abgr + SWS_FULL_CHR_H_INT: the reference overruns the destination (output.c:1231-1237); refused
                yield fmt, w, h, dw, dh, flags


@pytest.mark.parametrize("dfmt", [25, 26, 27, 28])
def test_port_matches_reference(orc, refo, dfmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for (fmt, w, h, dw, dh, flags) in combos(dfmt):
        pl = source(fmt, w, h, 3)
        a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags)
        assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), (fmt, dfmt, w, h, dw, dh, hex(flags), np.argwhere(a[1] != b[1])[:4].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("dfmt", [25, 26, 27, 28])
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
def test_gpu_device_batch_and_refusal(gpu, checker):
    from libav_b200 import device
    w, h, dw, dh, dfmt = 640, 480, 1280, 720, 28
    frames = [source(0, w, h, 30 + k) for k in range(3)]
    tight = [[np.ascontiguousarray(p[:, :p.shape[1]]) for p in f] for f in frames]
    src = [device.DevBuf.from_numpy(np.stack([f[i] for f in tight])) for i in range(3)]
    dst = device.DevBuf(3 * dh * dw * 4)
    ctx = device.SwsContext(w, h, dw, dh, dfmt, 4 | ACC)
    ctx.scale_device([s.ptr for s in src], [f.strides[0] for f in tight[0]], [dst.ptr], [dw * 4], nframes=3,
                     src_frame=[f.nbytes for f in tight[0]], dst_frame=[dh * dw * 4])
    got = dst.download(np.uint8, (3, dh, dw * 4))
    device.sync()
    for k in range(3):
        rc, want = run(checker, 0, frames[k], w, h, dfmt, dw, dh, 4 | ACC)
        assert np.array_equal(got[k], want[:, :dw * 4]), k
    ctx.close()
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 64, 48, 27, 4, src_fmt=2)          # rgb24 -> abgr of the same size (the reference's converter writes past the row)
    gpu.lib.avb200_clear_error()
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 128, 96, 27, 4 | ACC | 0x2000)
    gpu.lib.avb200_clear_error()
