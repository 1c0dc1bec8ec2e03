"""48-bit packed rgb destinations (rgb48be 34, rgb48le 35, bgr48be 59, bgr48le 60; libavutil/pixfmt.h): behind swscale() the lines are
hScale8To19_c's (swscale.c:62-80,728-741) and the output stage is yuv2rgb48_X / _2 / _1_c_template (output.c:593-760) under the X / 2 / 1
selection of swscale.c:658-683; the same-size case without SWS_ACCURATE_RND is the table converter yuv2rgb_c_48 / _bgr48 (yuv2rgb.c:106-236:
the 8-bit value in both bytes).  Planar 8-bit yuv sources (full-range ones included).
CPU: port vs the compiled reference; GPU: product vs checker (host-pointer and batched device-pointer calls).
Rows get room for one more pixel: for an odd width the reference's pair loop writes a whole pixel past the row (output.c:601,640-647)."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth

DST = [35, 34, 60, 59]
SUB = {0: (1, 1), 4: (1, 0), 5: (0, 0), 6: (2, 2), 7: (2, 0), 31: (0, 1), 12: (1, 1), 13: (1, 0), 14: (0, 0), 32: (0, 1)}
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 50, 67, 50), (66, 50, 33, 25), (64, 48, 64, 96), (64, 48, 65, 48)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 4, 0x10 | ACC, 4 | ACC | 0x2000, 1 | ACC, 0x200 | ACC)


def source(fmt, w, h, seed):
    r = np.random.RandomState(seed)
    hs, vs = SUB[fmt]
    cw, ch = -((-w) >> hs), -((-h) >> vs)
    return [synth.pad_rows(r.randint(0, 256, s).astype(np.uint8)) for s in ((h, w), (ch, cw), (ch, cw))]


def run(o, fmt, pl, w, h, dfmt, dw, dh, flags, pad=8):
    out = np.full((dh, dw * 6 + pad), 7, np.uint8)
    sp = (C.c_void_p * 3)(*[a.ctypes.data for a in pl])
    ss = (C.c_int * 3)(*[a.strides[0] for a in pl])
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    return o.sws_planar(fmt, sp, ss, w, h, dfmt, dp, ds, dw, dh, flags), out


def combos(thin=1):
    n = 0
    for fmt in sorted(SUB):
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS:
                n += 1
                if n % thin == 0:
                    yield fmt, w, h, dw, dh, flags


@pytest.mark.parametrize("dfmt", DST)
def test_port_matches_reference(orc, refo, dfmt):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for (fmt, w, h, dw, dh, flags) in combos():
        pl = source(fmt, w, h, 3)
        for pad in ((8, 0) if not dw & 1 else (8,)):
            a, b = run(refo, fmt, pl, w, h, dfmt, dw, dh, flags, pad), run(orc, fmt, pl, w, h, dfmt, dw, dh, flags, pad)
            assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), (fmt, dfmt, w, h, dw, dh, hex(flags), np.argwhere(a[1] != b[1])[:4].tolist())
            n += 1
    assert n > 500


def test_sixteen_bits_are_used_and_refusals(orc):
    """the scaled output is not the 8-bit value doubled; packed / semi-planar / high-bit-depth sources are refused"""
    pl = source(0, 64, 48, 9)
    rc, o = run(orc, 0, pl, 64, 48, 35, 128, 96, 4 | ACC)
    assert rc == 96
    px = o[:, :128 * 6].reshape(96, -1, 2)
    assert (px[..., 0] != px[..., 1]).mean() > 0.5          # (random input clips a good part of the samples to 0 / 0xFFFF)
    rc, o = run(orc, 0, pl, 64, 48, 35, 64, 48, 4)                       # the table converter: both bytes equal
    px = o[:, :64 * 6].reshape(48, -1, 2)
    assert rc == 48 and (px[..., 0] == px[..., 1]).all()
    r = np.random.RandomState(1)
    for sf, planes in ((2, [r.randint(0, 256, (48, 3 * 64 + 10)).astype(np.uint8)]), (23, [synth.pad_rows(r.randint(0, 256, (48, 64)).astype(np.uint8)), r.randint(0, 256, (24, 70)).astype(np.uint8)])):
        out = np.zeros((96, 128 * 6 + 8), np.uint8)
        sp = (C.c_void_p * 3)(*([a.ctypes.data for a in planes] + [None] * (3 - len(planes))))
        ss = (C.c_int * 3)(*([a.strides[0] for a in planes] + [0] * (3 - len(planes))))
        assert orc.sws_planar(sf, sp, ss, 64, 48, 35, (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0), 128, 96, 4 | ACC) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("dfmt", DST)
def test_gpu_matches_checker(gpu, checker, dfmt):
    from libav_b200 import device
    for (fmt, w, h, dw, dh, flags) in combos(thin=2):
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
    w, h, dw, dh = 640, 480, 1280, 720
    frames = [source(0, w, h, 30 + k) for k in range(3)]
    tight = [[np.ascontiguousarray(p[:, :p.shape[1]]) for p in f] for f in frames]
    src = [device.DevBuf.from_numpy(np.stack([f[i] for f in tight])) for i in range(3)]
    for dfmt, flags in ((35, 4 | ACC), (59, 2 | 0x80000)):
        dst = device.DevBuf(3 * dh * dw * 6)
        ctx = device.SwsContext(w, h, dw, dh, dfmt, flags)
        ctx.scale_device([s.ptr for s in src], [f.strides[0] for f in tight[0]], [dst.ptr], [dw * 6], nframes=3,
                         src_frame=[f.nbytes for f in tight[0]], dst_frame=[dh * dw * 6])
        got = dst.download(np.uint8, (3, dh, dw * 6))
        device.sync()
        for k in range(3):
            rc, want = run(checker, 0, frames[k], w, h, dfmt, dw, dh, flags)
            assert np.array_equal(got[k], want[:, :dw * 6]), (dfmt, k)
        ctx.close()
    for args in ((64, 48, 128, 96, 35, 4 | ACC, 2), (64, 48, 128, 96, 60, 4 | ACC, 23), (64, 48, 128, 96, 34, 4 | ACC, 64)):      # packed rgb, nv12, 10-bit sources
        with pytest.raises(Exception):
            device.SwsContext(*args[:6], src_fmt=args[6])
        gpu.lib.avb200_clear_error()


def test_colourspace_details(orc, refo, built):
    """sws_setColorspaceDetails with a 48-bit destination: the 16-bit coefficients of yuv2rgb.c:735-740 (other matrix, full range, brightness /
    contrast / saturation) -- port vs the compiled reference, and the product's kernels (host-compiled, tests/hostsim/) vs the reference"""
    import test_sws_colorspace as CS
    from test_hostsim_slots_cpu import sim as _sim_fixture            # noqa: F401
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")

    def run_cs(o, pl, w, h, dfmt, dw, dh, flags, cs):
        o.sws_set_colorspace((C.c_int * 4)(*cs[0]), cs[1], cs[2], cs[3], cs[4])
        try:
            return run(o, 0, pl, w, h, dfmt, dw, dh, flags)
        finally:
            o.sws_set_colorspace(None, 0, 0, 0, 0)
    n = 0
    for cs in CS.SETTINGS:
        for (w, h, dw, dh) in ((64, 48, 64, 48), (101, 37, 333, 211), (64, 48, 96, 80)):
            for flags in (4 | ACC, 4, 2 | 0x80000):
                for dfmt in (35, 59):
                    pl = source(0, w, h, 3)
                    a, b = run_cs(refo, pl, w, h, dfmt, dw, dh, flags, cs), run_cs(orc, pl, w, h, dfmt, dw, dh, flags, cs)
                    assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), (cs, w, h, dw, dh, hex(flags), dfmt)
                    n += 1
    assert n > 100
