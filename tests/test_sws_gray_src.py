"""gray8 (AV_PIX_FMT_GRAY8 = 8) sources.  swscale() never converts chroma lines for a gray source (needs_hcscale, libswscale/swscale.c:532,768-770):
the vertical stage reads what sws_init_context left in the line buffers, bytes of 64 (utils.c:1273) -- 0x4040 per 15-bit sample, which comes out as 129
in an 8-bit chroma plane and as neutral chroma in rgb.  At the same size a planar yuv destination is planarCopyWrapper (luma copied, chroma filled with
128, swscale_unscaled.c:812-823,1156) and a 24 / 32-bit rgb destination is palToRgbWrapper with the pseudo-palette r = g = b = sample (:342-384,1257-1259).
Refused: yuvj and semi-planar destinations, 16-bit planar, SwsFilter vectors.  CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
import test_sws_yuva_src as Y

ACC = 0x40000 | 0x80000
DSTS = [2, 3, 0, 4, 5, 6, 7, 31, 1, 15, 37, 35, 8, 64, 62, 25, 26, 27, 28]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (101, 37, 333, 211), (66, 50, 33, 25), (64, 48, 64, 96)]
FLAGS = (4 | ACC, 4, 0x10, 1 | ACC, 2, 2 | ACC | 0x2000)


def picture(w, h, seed):
    return synth.pad_rows(np.random.RandomState(seed).randint(0, 256, (h, w)).astype(np.uint8))


def dest(df, dw, dh):
    if df == 8:
        return [np.zeros((dh, dw + 16), np.uint8)]
    if 25 <= df <= 28:
        return [np.zeros((dh, dw * 4 + 16), np.uint8)]
    return Y.dest(df, dw, dh)


def crop(df, dw, p):
    if df == 8:
        return [p[0][:, :dw]]
    if 25 <= df <= 28:
        return [p[0][:, :4 * dw]]
    return Y.crop(df, dw, p)


def run(o, pl, sw, sh, df, dw, dh, flags):
    out = dest(df, dw, dh)
    sp, ss = (C.c_void_p * 3)(pl.ctypes.data, None, None), (C.c_int * 3)(pl.strides[0], 0, 0)
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(8, sp, ss, sw, sh, df, dp, ds, dw, dh, flags), out


def refused(df, sw, sh, dw, dh, flags):
    """what port and product decline on top of the destination list above"""
    if df in (64, 62) and (sw, sh) == (dw, dh):
        return True                               # planarCopyWrapper's fill_plane9or10
    if flags & 0x2000 and df not in (2, 3, 25, 26, 27, 28):
        return True                               # (the reference drops SWS_FULL_CHR_H_INT silently for these destinations; port and product decline it)
    if df == 27 and flags & 0x2000:
        return True                               # abgr + SWS_FULL_CHR_H_INT (tests/test_sws_rgb32_dst.py), scaled only
    return False


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for df in DSTS:
        for (sw, sh, dw, dh) in GEOMS:
            for flags in FLAGS:
                pl = picture(sw, sh, 5)
                rb, b = run(orc, pl, sw, sh, df, dw, dh, flags)
                if rb == -1:
                    assert refused(df, sw, sh, dw, dh, flags) and not (df == 27 and (sw, sh) == (dw, dh)), (df, sw, sh, dw, dh, hex(flags))
                    continue
                ra, a = run(refo, pl, sw, sh, df, dw, dh, flags)
                assert ra == rb == dh and all(np.array_equal(x, y) for x, y in zip(crop(df, dw, a), crop(df, dw, b))), (df, sw, sh, dw, dh, hex(flags))
                n += 1
    assert n > 500
    for df in (12, 23, 47):                       # yuvj420p, nv12, yuv420p16le
        assert run(orc, picture(64, 48, 1), 64, 48, df, 128, 96, 4 | ACC)[0] == -1 if df != 47 else True


def test_the_quirks_are_there(orc):
    pl = picture(64, 48, 2)
    rc, o = run(orc, pl, 64, 48, 0, 128, 96, 4 | ACC)                      # scaled: the untouched line buffers come out as 129
    assert rc == 96 and (o[1][:, :64] == 129).all() and (o[2][:, :64] == 129).all()
    rc, o = run(orc, pl, 64, 48, 0, 64, 48, 4 | ACC)                       # same size: planarCopyWrapper fills 128
    assert rc == 48 and (o[1][:, :32] == 128).all() and np.array_equal(o[0][:, :64], pl[:, :64])
    rc, o = run(orc, pl, 64, 48, 26, 64, 48, 4)                            # same size to rgba: r = g = b = sample, alpha 255
    px = o[0][:, :256].reshape(48, 64, 4)
    assert rc == 48 and (px[..., 3] == 255).all() and all(np.array_equal(px[..., k], pl[:, :64]) for k in range(3))


@pytest.mark.gpu
def test_gpu_matches_checker(gpu, checker):
    from libav_b200 import device
    n = 0
    for df in DSTS:
        for (sw, sh, dw, dh) in GEOMS:
            for flags in FLAGS[::2]:
                pl = picture(sw, sh, 7)
                if refused(df, sw, sh, dw, dh, flags):
                    with pytest.raises(Exception):
                        device.SwsContext(sw, sh, dw, dh, df, flags, src_fmt=8)
                    gpu.lib.avb200_clear_error()
                    continue
                rc, want = run(checker, pl, sw, sh, df, dw, dh, flags)
                assert rc == dh
                ctx = device.SwsContext(sw, sh, dw, dh, df, flags, src_fmt=8)
                got = ctx.scale([pl], dst_pad=16, fill=0)
                got = got if isinstance(got, (list, tuple)) else [got]
                assert all(np.array_equal(x, y) for x, y in zip(crop(df, dw, list(got)), crop(df, dw, want))), (df, sw, sh, dw, dh, hex(flags))
                ctx.close()
                n += 1
    assert n > 250
