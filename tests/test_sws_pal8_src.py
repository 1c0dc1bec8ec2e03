"""pal8 (AV_PIX_FMT_PAL8 = 11) sources: plane 0 = one index per pixel, plane 1 = 256 native-endian 0xAARRGGBB entries.  sws_scale() converts the palette
to limited-range y / u / v per entry on every call (libswscale/swscale_unscaled.c:1236-1268) and the readers palToY_c / palToUV_c look the samples up
(input.c:321-343), chroma at full resolution; at the same size a 24 / 32-bit rgb destination is palToRgbWrapper (:342-384), a lookup of r, g, b with
alpha 255.  The palette's alpha byte is never used.  Refused: 15 / 16 / 48-bit rgb destinations, SwsFilter vectors.
(SWS_FAST_BILINEAR up-scaling is left out: the reference reads its uncleared conversion buffer past the line there, tests/test_sws_packed_sources.py.)
CPU: port vs the compiled reference; GPU: product vs checker."""
import ctypes as C

import numpy as np
import pytest

import test_sws_gray_src as G

ACC = 0x40000 | 0x80000
DSTS = [2, 3, 0, 4, 5, 6, 7, 31, 1, 15, 8, 64, 62, 25, 26, 27, 28, 12, 23]
GEOMS = G.GEOMS
FLAGS = (4 | ACC, 4, 0x10, 1 | ACC, 2, 2 | ACC | 0x2000)


def picture(w, h, seed):
    r = np.random.RandomState(seed)
    idx = r.randint(0, 256, (h, w + 16)).astype(np.uint8)
    idx[:, w:] = idx[:, w - 1:w]
    return idx, r.randint(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)


def dest(df, dw, dh):
    if df == 23:
        return [np.zeros((dh, dw + 8), np.uint8), np.zeros(((dh + 1) // 2, 2 * ((dw + 1) // 2) + 8), np.uint8)]
    return G.dest(df, dw, dh)


def run(o, idx, pal, sw, sh, df, dw, dh, flags):
    out = dest(df, dw, dh)
    sp, ss = (C.c_void_p * 3)(idx.ctypes.data, pal.ctypes.data, None), (C.c_int * 3)(idx.strides[0], 1024, 0)
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in out] + [None] * (3 - len(out))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in out] + [0] * (3 - len(out))))
    return o.sws_planar(11, sp, ss, sw, sh, df, dp, ds, dw, dh, flags), out


def skipped(df, sw, sh, dw, dh, flags):
    if flags & 1 and dw > sw:
        return True                               # undefined right edge in the reference (see the module docstring)
    if flags & 0x2000 and df not in (2, 3, 25, 26, 27, 28):
        return True                               # (the reference drops the flag silently there; port and product decline it)
    return False


def refused(df, sw, sh, dw, dh, flags):
    return df == 27 and flags & 0x2000 and (sw, sh) != (dw, dh)          # abgr + SWS_FULL_CHR_H_INT (tests/test_sws_rgb32_dst.py), scaled only


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    n = 0
    for df in DSTS:
        for (sw, sh, dw, dh) in GEOMS:
            for flags in FLAGS:
                if skipped(df, sw, sh, dw, dh, flags):
                    continue
                idx, pal = picture(sw, sh, 5)
                rb, b = run(orc, idx, pal, sw, sh, df, dw, dh, flags)
                if refused(df, sw, sh, dw, dh, flags):
                    assert rb == -1
                    continue
                ra, a = run(refo, idx, pal, sw, sh, df, dw, dh, flags)
                assert ra == rb == dh and all(np.array_equal(x, y) for x, y in zip(G.crop(df, dw, a), G.crop(df, dw, b))), (df, sw, sh, dw, dh, hex(flags))
                n += 1
    assert n > 350
    idx, pal = picture(64, 48, 1)
    for df in (37, 35):
        assert run(orc, idx, pal, 64, 48, df, 128, 96, 4 | ACC)[0] == -1


def test_palette_semantics(orc):
    idx, pal = picture(64, 48, 2)
    rc, o = run(orc, idx, pal, 64, 48, 28, 64, 48, 4)                      # same size to bgra: b, g, r of the entry, alpha 255 whatever the entry's top byte
    px = o[0][:, :256].reshape(48, 64, 4)
    e = pal[idx[:, :64]]
    assert rc == 48 and (px[..., 3] == 255).all() and np.array_equal(px[..., 0], e & 255) and np.array_equal(px[..., 2], (e >> 16) & 255)


@pytest.mark.gpu
def test_gpu_matches_checker(gpu, checker):
    from libav_b200 import device
    n = 0
    for df in DSTS:
        for (sw, sh, dw, dh) in GEOMS:
            for flags in FLAGS[::2]:
                if skipped(df, sw, sh, dw, dh, flags):
                    continue
                idx, pal = picture(sw, sh, 7)
                rc, want = run(checker, idx, pal, sw, sh, df, dw, dh, flags)
                assert rc == dh
                ctx = device.SwsContext(sw, sh, dw, dh, df, flags, src_fmt=11)
                got = ctx.scale([idx, pal.view(np.uint8).reshape(1, 1024)], dst_pad=16, fill=0)
                got = got if isinstance(got, (list, tuple)) else [got]
                assert all(np.array_equal(x, y) for x, y in zip(G.crop(df, dw, list(got)), G.crop(df, dw, want))), (df, sw, sh, dw, dh, hex(flags))
                ctx.close()
                n += 1
    assert n > 200
    for df in (37, 35):
        with pytest.raises(Exception):
            device.SwsContext(64, 48, 128, 96, df, 4 | ACC, src_fmt=11)
        gpu.lib.avb200_clear_error()
