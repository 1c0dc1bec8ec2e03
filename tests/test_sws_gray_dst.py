"""gray8 destination (AV_PIX_FMT_GRAY8 = 8): the luma plane alone -- swscale() skips the chroma of a gray destination (swscale.c:618-630), the
same-size case is the plane copy for every planar yuv source (swscale_unscaled.c:1155), a full-range (yuvj) source gets lumRangeFromJpeg.
CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

from test_sws_planar_dst import source

SRC = [0, 4, 5, 23, 1, 2, 12, 3, 15, 26]        # 3 = bgr24: at the same size without SWS_ACCURATE_RND a yuv420p destination would get rgb24toyv12 -- a gray8 one does not
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (101, 37, 333, 211), (67, 51, 67, 51), (66, 50, 33, 25)]
ACC = 0x40000 | 0x80000
FLAGS = (4 | ACC, 2 | 0x80000, 4, 0x10 | ACC, 1 | ACC, 0x200 | ACC)


def planes(fmt, w, h, seed):
    if fmt == 26:
        return [np.random.RandomState(seed).randint(0, 256, (h, 4 * w + 12)).astype(np.uint8)]
    return source({12: 0, 3: 2}.get(fmt, fmt), w, h, seed)


def run(o, fmt, pl, w, h, dw, dh, flags):
    out = np.full((dh, dw + 8), 7, np.uint8)
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp, ds = (C.c_void_p * 3)(out.ctypes.data, None, None), (C.c_int * 3)(out.strides[0], 0, 0)
    return o.sws_planar(fmt, sp, ss, w, h, 8, dp, ds, dw, dh, flags), out


def combos():
    for fmt in SRC:
        for (w, h, dw, dh) in GEOMS:
            for flags in FLAGS:
                if flags & 1 and fmt in (23, 1, 2, 3, 15, 26) and dw > w:
                    continue         # undefined right edge, see tests/test_sws_packed_sources.py
                yield fmt, w, h, dw, dh, flags


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for (fmt, w, h, dw, dh, flags) in combos():
        pl = planes(fmt, w, h, 3)
        a, b = run(refo, fmt, pl, w, h, dw, dh, flags), run(orc, fmt, pl, w, h, dw, dh, flags)
        assert a[0] == b[0] == dh and np.array_equal(a[1][:, :dw], b[1][:, :dw]), (fmt, w, h, dw, dh, hex(flags), np.argwhere(a[1] != b[1])[:4].tolist())
        assert (b[1][:, dw:] == 7).all()
        n += 1
    assert n > 200


@pytest.mark.gpu
def test_gpu_matches_checker(gpu, checker):
    from libav_b200 import device
    for (fmt, w, h, dw, dh, flags) in combos():
        pl = planes(fmt, w, h, 5)
        rc, want = run(checker, fmt, pl, w, h, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, 8, flags, src_fmt=fmt)
        got = ctx.scale(pl, dst_pad=8, fill=7)
        assert np.array_equal(got[:, :dw], want[:, :dw]) and (got[:, dw:] == 7).all(), (fmt, w, h, dw, dh, hex(flags), np.argwhere(got != want)[:4].tolist())
        ctx.close()


@pytest.mark.gpu
def test_gpu_device_batch(gpu, checker):
    from libav_b200 import device
    w, h, dw, dh = 640, 480, 1280, 720
    frames = [source(0, w, h, 30 + k) for k in range(3)]
    tight = [[np.ascontiguousarray(p[:, :p.shape[1]]) for p in f] for f in frames]
    src = [device.DevBuf.from_numpy(np.stack([f[i] for f in tight])) for i in range(3)]
    dst = device.DevBuf(3 * dh * dw)
    ctx = device.SwsContext(w, h, dw, dh, 8, 4 | ACC)
    ctx.scale_device([s.ptr for s in src], [f.strides[0] for f in tight[0]], [dst.ptr], [dw], nframes=3,
                     src_frame=[f.nbytes for f in tight[0]], dst_frame=[dh * dw])
    got = dst.download(np.uint8, (3, dh, dw))
    device.sync()
    for k in range(3):
        rc, want = run(checker, 0, frames[k], w, h, dw, dh, 4 | ACC)
        assert np.array_equal(got[k], want[:, :dw]), k
    ctx.close()
