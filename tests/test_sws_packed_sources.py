"""Packed sources (yuyv422, uyvy422, rgb24, bgr24, argb, rgba, abgr, bgra) -> rgb24 / bgr24 / yuv420p: the reference's input readers (libswscale/input.c)
in front of the scaler, and its unscaled special converters (rgb24 <-> bgr24, rgb24toyv12_c, yuyvtoyuv420_c / uyvytoyuv420_c).
CPU: port vs the compiled reference; GPU: product vs checker, host-pointer and device-pointer (batched) calls."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth

BPP = {1: 2, 15: 2, 2: 3, 3: 3, 25: 4, 26: 4, 27: 4, 28: 4}
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 51, 67, 51), (66, 50, 33, 25),
         (64, 48, 160, 48)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 0x10 | ACC, 1 | ACC, 4 | ACC | 0x2000, 4 | ACC | 0x4000, 0x200 | 0x80000)


def frame(fmt, w, h, seed, pad=10):
    r = np.random.RandomState(seed)
    buf = r.randint(0, 256, (h, w * BPP[fmt] + pad)).astype(np.uint8)
    return buf


def outputs(dfmt, dw, dh):
    if dfmt:
        return [np.full((dh, dw * 3 + 6), 7, np.uint8)]
    return [np.full((dh, dw), 7, np.uint8), np.full(((dh + 1) // 2, (dw + 1) // 2), 7, np.uint8), np.full(((dh + 1) // 2, (dw + 1) // 2), 7, np.uint8)]


def run(o, fmt, buf, w, h, dfmt, dw, dh, flags):
    out = outputs(dfmt, dw, dh)
    sp, ss = (C.c_void_p * 3)(buf.ctypes.data, None, None), (C.c_int * 3)(buf.strides[0], 0, 0)
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


def combos(fmt):
    for (w, h, dw, dh) in GEOMS:
        for flags in FLAGS:
            for dfmt in (2, 3, 0):
                if dfmt == 0 and flags & 0x2000:
                    continue
                if fmt == 3 and dfmt == 0 and (w, h) == (dw, dh) and not flags & 0x40000 and h & 1:
                    continue        # rgb24toyv12_c converts rows in pairs: an odd height is refused
                if fmt >= 25 and dfmt and (w, h) == (dw, dh):
                    continue        # 32-bit rgb -> packed rgb of the same size: the rgb2rgb converter family, refused
                yield w, h, dw, dh, flags, dfmt


def same(a, b, dfmt, dw, skip=0):
    """skip: pixel columns at the right edge left out of the comparison (see fast_bilinear_margin)"""
    if dfmt:                     # the pixel area; the reference's whole-buffer copies also touch row padding
        return np.array_equal(a[0][:, :(dw - skip) * 3], b[0][:, :(dw - skip) * 3])
    return all(np.array_equal(x[:, :x.shape[1] - skip], y[:, :y.shape[1] - skip]) for x, y in zip(a, b))


def fast_bilinear_margin(flags, sw, dw):
    """hyscale_fast_c / hcscale_fast_c (swscale.c:238-300) read src[xx + 1] one sample past the converted line; for a packed source
    that line lives in formatConvBuffer, which the reference allocates without clearing (utils.c:1064): the last up-scaled
    pixels depend on uninitialised memory there.  Product and port define that sample as a copy of the last one."""
    return 2 * -(-dw // sw) + 2 if (flags & 1) and dw > sw else 0


@pytest.mark.parametrize("fmt", [1, 15, 2, 3, 25, 26, 27, 28])
def test_port_matches_reference(orc, refo, fmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for (w, h, dw, dh, flags, dfmt) in combos(fmt):
        buf = frame(fmt, w, h, 3)
        a, b = run(refo, fmt, buf, w, h, dfmt, dw, dh, flags), run(orc, fmt, buf, w, h, dfmt, dw, dh, flags)
        assert a[0] == b[0] == dh and same(a[1], b[1], dfmt, dw, fast_bilinear_margin(flags, w, dw)), (fmt, w, h, dw, dh, hex(flags), dfmt)
        n += 1
    assert n > 100


def test_port_tight_rows(orc, refo):
    """no row padding: the sample the chroma readers take past an odd width is the next row's first one"""
    if refo is None:
        pytest.skip("oracle/_ref not built")
    for fmt in (1, 15, 2, 3):
        for (w, h, dw, dh) in ((101, 36, 333, 210), (67, 50, 67, 50)):
            buf = frame(fmt, w, h + 1, 9, pad=0)         # one spare row keeps the reference's over-read inside the array
            for dfmt in (2, 0):
                a, b = run(refo, fmt, buf, w, h, dfmt, dw, dh, 4 | ACC), run(orc, fmt, buf, w, h, dfmt, dw, dh, 4 | ACC)
                # the last source row's over-read would leave the frame: port and product repeat the pixel instead, the reference
                # reads the spare row -- output rows that depend on the last source row are left out
                m = 3 * -(-dh // h) + 2
                cut = lambda planes: [p[:(dh - m) * p.shape[0] // dh] for p in planes]
                assert a[0] == b[0] == dh and same(cut(a[1]), cut(b[1]), dfmt, dw), (fmt, w, h, dfmt)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [1, 15, 2, 3, 25, 26, 27, 28])
def test_gpu_matches_checker(gpu, checker, orc, fmt):
    from libav_b200 import device
    for (w, h, dw, dh, flags, dfmt) in combos(fmt):
        buf = frame(fmt, w, h, 5)
        rc, want = run(orc if fast_bilinear_margin(flags, w, dw) else checker, fmt, buf, w, h, dfmt, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
        got = ctx.scale([buf], dst_pad=6, fill=7)
        got = [got] if dfmt else got
        assert same(got, want, dfmt, dw), (fmt, w, h, dw, dh, hex(flags), dfmt)
        if dfmt == 0:                # rows / columns the special converters leave alone stay untouched
            assert all(np.array_equal(x, y) for x, y in zip(got, want))
        ctx.close()


@pytest.mark.gpu
def test_gpu_device_batch(gpu, checker):
    """device pointers, 3 frames per call: rgb24 1080p -> yuv420p 720p (encoder ingest) and yuyv422 -> rgb24"""
    from libav_b200 import device
    for fmt, dfmt, (w, h, dw, dh) in ((2, 0, (1920, 1080, 1280, 720)), (1, 2, (640, 480, 640, 480)), (3, 0, (640, 480, 640, 480))):
        flags = 4 | (0 if (fmt, dfmt) == (3, 0) else ACC)
        frames = [frame(fmt, w, h, 20 + k, pad=0) for k in range(3)]
        src = device.DevBuf.from_numpy(np.stack(frames))
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags, src_fmt=fmt)
        if dfmt:
            dst = device.DevBuf(3 * dh * dw * 3)
            ctx.scale_device([src.ptr], [frames[0].strides[0]], [dst.ptr], [dw * 3], nframes=3, src_frame=[frames[0].nbytes], dst_frame=[dh * dw * 3])
            got = dst.download(np.uint8, (3, dh, dw * 3))
        else:
            cw, ch = (dw + 1) // 2, (dh + 1) // 2
            fb = dw * dh + 2 * cw * ch
            dst = device.DevBuf(3 * fb)
            ctx.scale_device([src.ptr], [frames[0].strides[0]], [dst.ptr, dst.ptr + dw * dh, dst.ptr + dw * dh + cw * ch], [dw, cw, cw], nframes=3,
                             src_frame=[frames[0].nbytes], dst_frame=[fb, fb, fb])
            got = dst.download(np.uint8, (3, fb))
        device.sync()
        for k in range(3):
            rc, want = run(checker, fmt, frames[k], w, h, dfmt, dw, dh, flags)
            if dfmt:
                assert np.array_equal(got[k], want[0][:, :dw * 3]), (fmt, dfmt, k)
            else:
                assert np.array_equal(got[k], np.concatenate([p.ravel() for p in want])), (fmt, dfmt, k)
        ctx.close()


@pytest.mark.gpu
def test_refusals_are_loud(gpu):
    from libav_b200 import device
    with pytest.raises(Exception):
        device.SwsContext(64, 49, 64, 49, device.PIX_FMT_YUV420P, 4, src_fmt=3)      # rgb24toyv12_c with an odd height
    gpu.lib.avb200_clear_error()
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 64, 48, 25, 4, src_fmt=2)                          # rgb24 -> argb of the same size: the reference's writer runs past the row
    gpu.lib.avb200_clear_error()
    with pytest.raises(Exception):
        device.SwsContext(64, 48, 128, 96, 28, 4, src_fmt=26)                        # rgba -> bgra: the alpha plane would be scaled too
    gpu.lib.avb200_clear_error()
