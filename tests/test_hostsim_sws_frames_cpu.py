"""CPU: whole frames through the product's scaler compiled for the host (tests/hostsim/swscale_hostsim.cpp: every kernel of swscale.cu except
the two shared-memory tile kernels is thread-independent; the general scaler runs on its two-pass path) -- sws_getContext_cuda /
sws_scale_cuda of that build against the compiled reference.  This is where the parts of the frame path written without a GPU at hand are
checked end to end: the yuvj <-> yuv range conversion, SwsFilter vectors, bottom-up pictures; plus the fused same-size kernels, the
unscaled converters, the source readers and every output stage the two-pass path has.  16-bit destinations (tile kernel only) stay GPU-only."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from libav_b200.device import PLANAR_BE, PLANAR_FORMATS
from test_hostsim_slots_cpu import sim          # noqa: F401
from test_sws_planar_dst import source

ACC = 0x40000 | 0x80000


def outputs(df, dw, dh, pad=8, fill=7):
    if df in (23, 24):
        return [np.full((dh, dw + pad), fill, np.uint8), np.full(((dh + 1) // 2, 2 * ((dw + 1) // 2) + pad), fill, np.uint8)]
    if df in PLANAR_FORMATS:
        hs, vs, bits = PLANAR_FORMATS[df]
        dt = np.uint8 if bits == 8 else np.dtype(">u2" if df in PLANAR_BE else "<u2")
        cw, ch = -((-dw) >> hs), -((-dh) >> vs)
        return [np.full((dh, dw + pad), fill, dt), np.full((ch, cw + pad), fill, dt), np.full((ch, cw + pad), fill, dt)]
    bpp = 6 if df in (34, 35, 59, 60) else 4 if 25 <= df <= 28 else 2 if df in (1, 15) or 36 <= df <= 43 or 54 <= df <= 57 else 1 if df == 8 else 3
    return [np.full((dh, dw * bpp + 2 * pad), fill, np.uint8)]


def arrays(planes):
    return (C.c_void_p * 4)(*([a.ctypes.data for a in planes] + [None] * (4 - len(planes)))), (C.c_int * 4)(*([a.strides[0] for a in planes] + [0] * (4 - len(planes))))


def product(sim, sf, pl, w, h, df, dw, dh, flags, src_filter=None, dst_filter=None, outs=None):
    ctx = sim.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, src_filter, dst_filter, None)
    assert ctx, (sf, df, w, h, dw, dh, hex(flags), sim.avb200_last_error())
    outs = outputs(df, dw, dh) if outs is None else outs
    sp, ss = arrays(pl)
    dp, ds = arrays(outs)
    assert sim.sws_scale_cuda(ctx, sp, ss, 0, h, dp, ds) == dh, sim.avb200_last_error()
    sim.sws_freeContext_cuda(ctx)
    return outs


def reference(refo, sf, pl, w, h, df, dw, dh, flags, outs=None):
    outs = outputs(df, dw, dh) if outs is None else outs
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in outs] + [None] * (3 - len(outs))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in outs] + [0] * (3 - len(outs))))
    assert refo.sws_planar(sf, sp, ss, w, h, df, dp, ds, dw, dh, flags) == dh
    return outs


def same(a, b, what, crop=0):
    """crop > 0: compare the pictures only, not the row padding (the reference's plane copies move whole strides when the pitches agree)"""
    for x, y in zip(a, b):
        if crop:
            x, y = x[:, :x.shape[1] - crop], y[:, :y.shape[1] - crop]
        assert np.array_equal(x, y), (what, np.argwhere(x != y)[:4].tolist())


def test_known_answer_crc_fused_kernel(sim):
    """av_crc(AV_CRC_32_IEEE) of the rgb24 picture of the LFG-seed-1 640x480 frame = f046d710 -- here through the fused dp4a kernel"""
    yuv = synth.yuv420p_frame(640, 480, 1)
    out = product(sim, 0, yuv, 640, 480, 2, 640, 480, 4 | ACC, outs=[np.zeros((480, 640 * 3), np.uint8)])
    assert synth.crc32_ieee_be(out[0].tobytes()) == 0xF046D710


def test_format_matrix(sim, refo):
    n = 0
    for sf in (0, 4, 5, 23, 1, 2, 15):
        for df in (2, 3, 28, 25, 0, 5, 62, 63, 23, 24, 1, 15, 4):
            for (w, h, dw, dh) in ((64, 48, 64, 48), (66, 50, 33, 25), (64, 48, 96, 80), (101, 37, 64, 48)):
                for flags in (4 | ACC, 2, 0x10 | ACC, 4 | ACC | 0x2000):
                    if (flags & 0x2000) and df not in (2, 3, 28, 25):
                        continue
                    if sf in (23,) and df in (23, 24) and (w, h) == (dw, dh):
                        continue                                   # refused: the reference's plane copy skips the chroma plane
                    if sf in (2,) and df in (28, 25) and (w, h) == (dw, dh):
                        continue                                   # rgb2rgb converter family: refused
                    if sf == 3 and df == 0 and (w, h) == (dw, dh) and not (flags & 0x40000) and h & 1:
                        continue
                    pl = source(sf, w, h, 41)
                    same(product(sim, sf, pl, w, h, df, dw, dh, flags), reference(refo, sf, pl, w, h, df, dw, dh, flags), (sf, df, w, h, dw, dh, hex(flags)), crop=8)
                    n += 1
    assert n > 900


def test_rgb16_destinations(sim, refo):
    """rgb565 / bgr565 / rgb555 / bgr555 / rgb444 / bgr444, LE and BE (tests/test_sws_rgb16_dst.py): the dithered 16-bpp output stage of the
    two-pass path, every source it is taken over for, full-range sources and 10-bit sources included"""
    import test_sws_rgb16_dst as R
    n = 0
    for df in R.DST:
        for (sf, w, h, dw, dh, flags) in R.combos(df):
            if n % 3 and (w, h, dw, dh) != (67, 50, 67, 50):
                n += 1
                continue                         # (a third of the matrix per format: the port covers all of it against the reference)
            pl = source(sf, w, h, 43)
            same(product(sim, sf, pl, w, h, df, dw, dh, flags), reference(refo, sf, pl, w, h, df, dw, dh, flags), (sf, df, w, h, dw, dh, hex(flags)), crop=8)
            n += 1
    assert n > 1000
    import test_sws_hbd_sources_cpu as H
    for sf in (12, 64):                          # yuvj420p (the range folds into the colour constants), yuv420p10le (hScale16To15 lines)
        for df in (37, 43, 55):
            pl = source(0, 66, 50, 44) if sf == 12 else H.planes(sf, 66, 50, 5)
            same(product(sim, sf, pl, 66, 50, df, 100, 80, 4 | ACC), reference(refo, sf, pl, 66, 50, df, 100, 80, 4 | ACC), (sf, df), crop=8)
    # refusals: the ordered-dither table converter's case, packed rgb sources, the per-line slots
    for args in ((64, 48, 0, 64, 48, 37, 4), (64, 48, 2, 128, 96, 37, 4 | ACC), (64, 48, 1, 128, 96, 41, 4 | ACC)):
        assert not sim.sws_getContext_cuda(*args, None, None, None)
        sim.avb200_clear_error()


def test_rgb48_destinations(sim, refo):
    """rgb48 / bgr48, LE and BE (tests/test_sws_rgb48_dst.py): the direct 19-bit-line kernel and the byte-doubling table converter"""
    import test_sws_rgb48_dst as R
    n = 0
    for k, df in enumerate(R.DST):
        for (sf, w, h, dw, dh, flags) in R.combos(thin=2):
            if (n + k) % 2 and (w, h, dw, dh) not in ((67, 50, 67, 50), (64, 48, 65, 48)):
                n += 1
                continue
            pl = R.source(sf, w, h, 43)
            same(product(sim, sf, pl, w, h, df, dw, dh, flags), reference(refo, sf, pl, w, h, df, dw, dh, flags), (sf, df, w, h, dw, dh, hex(flags)), crop=8)
            n += 1
    assert n > 500
    for args in ((64, 48, 2, 128, 96, 35, 4 | ACC), (64, 48, 23, 128, 96, 59, 4 | ACC), (64, 48, 64, 128, 96, 34, 4 | ACC)):
        assert not sim.sws_getContext_cuda(*args, None, None, None)
        sim.avb200_clear_error()


def test_rgb48_colourspace_details(sim, refo):
    """sws_setColorspaceDetails_cuda reaches the 48-bit output stage (the 16-bit coefficients of yuv2rgb.c:735-740) and its byte-doubling table converter"""
    import test_sws_colorspace as CS
    import test_sws_rgb48_dst as R
    for cs in CS.SETTINGS[:5]:
        tab = (C.c_int * 4)(*cs[0])
        for (w, h, dw, dh, flags) in ((64, 48, 64, 48, 4), (101, 37, 333, 211, 4 | ACC), (64, 48, 96, 80, 2 | 0x80000)):
            for df in (35, 59):
                pl = R.source(0, w, h, 3)
                refo.sws_set_colorspace(tab, cs[1], cs[2], cs[3], cs[4])
                try:
                    rc, want = R.run(refo, 0, pl, w, h, df, dw, dh, flags)
                finally:
                    refo.sws_set_colorspace(None, 0, 0, 0, 0)
                assert rc == dh
                ctx = sim.sws_getContext_cuda(w, h, 0, dw, dh, df, flags, None, None, None)
                assert ctx and sim.sws_setColorspaceDetails_cuda(ctx, tab, cs[1], tab, 0, cs[2], cs[3], cs[4]) == 0, sim.avb200_last_error()
                out = np.full((dh, dw * 6 + 8), 7, np.uint8)
                sp, ss = arrays(pl)
                dp, ds = arrays([out])
                assert sim.sws_scale_cuda(ctx, sp, ss, 0, h, dp, ds) == dh, sim.avb200_last_error()
                sim.sws_freeContext_cuda(ctx)
                assert np.array_equal(out[:, :6 * dw], want[:, :6 * dw]), (cs, w, h, dw, dh, hex(flags), df)


def test_frames_call_with_one_plane_sources(sim, refo):
    """sws_scale_frames_cuda (the device-pointer batch call; here "device memory" is host memory): a gray8 batch passes one plane, the entry point hands
    the kernels the luma plane for the two nobody reads; two frames, every path a gray source has"""
    import test_sws_gray_src as G
    sim.sws_scale_frames_cuda.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
    for df, (w, h, dw, dh), flags in ((2, (64, 48, 96, 80), 4 | ACC), (0, (64, 48, 64, 48), 4), (26, (64, 48, 64, 48), 4), (35, (66, 50, 33, 25), 2), (0, (64, 48, 96, 80), 4)):
        frames = [np.ascontiguousarray(G.picture(w, h, 60 + k)[:, :w]) for k in range(2)]
        src = np.stack(frames)
        want = [G.run(refo, f, w, h, df, dw, dh, flags)[1] for f in frames]
        outs = [np.stack([np.zeros_like(p) for _ in range(2)]) for p in G.dest(df, dw, dh)]
        ctx = sim.sws_getContext_cuda(w, h, 8, dw, dh, df, flags, None, None, None)
        assert ctx, sim.avb200_last_error()
        sp, ss, sf = (C.c_void_p * 3)(src.ctypes.data, None, None), (C.c_int * 3)(src.strides[1], 0, 0), (C.c_size_t * 3)(src.strides[0], 0, 0)
        dp = (C.c_void_p * 3)(*([o.ctypes.data for o in outs] + [None] * (3 - len(outs))))
        ds = (C.c_int * 3)(*([o.strides[1] for o in outs] + [0] * (3 - len(outs))))
        dfr = (C.c_size_t * 3)(*([o.strides[0] for o in outs] + [0] * (3 - len(outs))))
        assert sim.sws_scale_frames_cuda(ctx, sp, ss, sf, dp, ds, dfr, 2, None) == 2 * dh, sim.avb200_last_error()
        sim.sws_freeContext_cuda(ctx)
        for k in range(2):
            assert all(np.array_equal(x, y) for x, y in zip(G.crop(df, dw, [o[k] for o in outs]), G.crop(df, dw, want[k]))), (df, k)


def test_colourspace_details_with_the_late_sources(sim, refo):
    """sws_setColorspaceDetails_cuda on contexts with gray8 / pal8 / yuva420p sources and rgb destinations (the colour tables change, the constant
    chroma lines of a gray source and the palette conversion do not)"""
    import test_sws_colorspace as CS
    import test_sws_gray_src as G
    import test_sws_pal8_src as P
    import test_sws_rgb48_dst as R
    for cs in CS.SETTINGS[:4]:
        tab = (C.c_int * 4)(*cs[0])
        for sf in (8, 11, 33):
            for df, (w, h, dw, dh), flags in ((2, (64, 48, 96, 80), 4 | ACC), (28, (101, 37, 333, 211), 2), (35, (64, 48, 96, 80), 4 | ACC), (3, (64, 48, 64, 48), 4)):
                if (sf == 11 and df == 35) or (sf == 33 and df == 28):
                    continue                     # refused pairs (pal8 -> 48 bit, yuva420p -> a destination with alpha)
                if sf == 8:
                    pl = [G.picture(w, h, 9)]
                elif sf == 11:
                    idx, pal = P.picture(w, h, 9)
                    pl = [idx, pal.view(np.uint8).reshape(1, 1024)]
                else:
                    pl = R.source(0, w, h, 9)
                refo.sws_set_colorspace(tab, cs[1], cs[2], cs[3], cs[4])
                try:
                    want = reference(refo, sf, pl, w, h, df, dw, dh, flags, outs=outputs(df, dw, dh))
                finally:
                    refo.sws_set_colorspace(None, 0, 0, 0, 0)
                ctx = sim.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, None, None, None)
                assert ctx and sim.sws_setColorspaceDetails_cuda(ctx, tab, cs[1], tab, 0, cs[2], cs[3], cs[4]) == 0, sim.avb200_last_error()
                got = outputs(df, dw, dh)
                sp, ss = arrays(pl)
                dp, ds = arrays(got)
                assert sim.sws_scale_cuda(ctx, sp, ss, 0, h, dp, ds) == dh, sim.avb200_last_error()
                sim.sws_freeContext_cuda(ctx)
                same(got, want, (cs, sf, df, w, h, dw, dh, hex(flags)), crop=8)


def test_yuva420p_sources(sim, refo):
    """yuva420p (33) to destinations without alpha: the product reads three planes like the reference does (tests/test_sws_yuva_src.py)"""
    import test_sws_rgb48_dst as R
    for df in (2, 3, 0, 37, 35):
        for (w, h, dw, dh) in ((64, 48, 64, 48), (352, 288, 640, 480), (101, 37, 333, 211)):
            for flags in (4 | ACC, 4):
                if df == 37 and (w, h) == (dw, dh) and not flags & 0x40000:
                    continue                     # (the ordered-dither table converter: refused for yuv420p too)
                pl = R.source(0, w, h, 47)
                same(product(sim, 33, pl, w, h, df, dw, dh, flags), reference(refo, 33, pl, w, h, df, dw, dh, flags), (df, w, h, dw, dh, hex(flags)), crop=8)
    assert not sim.sws_getContext_cuda(64, 48, 33, 128, 96, 26, 4 | ACC, None, None, None)
    sim.avb200_clear_error()


def test_rgb2rgb_converters(sim, refo):
    """same-size 24 <-> 32 and 32 <-> 32 bit packed rgb (tests/test_sws_rgb2rgb.py): the channel-remap kernel against rgbToRgbWrapper"""
    import test_sws_rgb2rgb as R
    n = 0
    for sf, df, refused in R.pairs():
        for (w, h) in R.GEOMS:
            if refused:
                assert not sim.sws_getContext_cuda(w, h, sf, w, h, df, 4, None, None, None)
                sim.avb200_clear_error()
                continue
            src = R.picture(sf, w, h, 21 + w, 6)
            same(product(sim, sf, [src], w, h, df, w, h, 4), reference(refo, sf, [src], w, h, df, w, h, 4), (sf, df, w, h), crop=8)
            n += 1
    assert n > 100


def test_gray8_sources(sim, refo):
    """gray8 sources (tests/test_sws_gray_src.py): constant chroma lines in the two-pass path, the plane copy with 128 fill, the pseudo-palette converter"""
    import test_sws_gray_src as G
    n = 0
    for df in G.DSTS:
        for (w, h, dw, dh) in G.GEOMS:
            for flags in G.FLAGS[::2]:
                pl = G.picture(w, h, 49)
                if G.refused(df, w, h, dw, dh, flags):
                    assert not sim.sws_getContext_cuda(w, h, 8, dw, dh, df, flags, None, None, None), (df, w, h, dw, dh, hex(flags))
                    sim.avb200_clear_error()
                    continue
                rc, want = G.run(refo, pl, w, h, df, dw, dh, flags)
                assert rc == dh
                got = product(sim, 8, [pl], w, h, df, dw, dh, flags, outs=G.dest(df, dw, dh))
                assert all(np.array_equal(x, y) for x, y in zip(G.crop(df, dw, got), G.crop(df, dw, want))), (df, w, h, dw, dh, hex(flags))
                n += 1
    assert n > 250
    for df in (12, 23, 47):
        assert not sim.sws_getContext_cuda(64, 48, 8, 128, 96, df, 4 | ACC, None, None, None)
        sim.avb200_clear_error()


def test_pal8_sources(sim, refo):
    """pal8 sources (tests/test_sws_pal8_src.py): the palette reader in front of the planar kernels, palToRgbWrapper's lookup; whole frames, a batch through the
    device-pointer call with one palette per frame, a bottom-up picture"""
    import test_sws_pal8_src as P
    import test_sws_gray_src as G
    n = 0
    for df in P.DSTS:
        for (w, h, dw, dh) in P.GEOMS:
            for flags in P.FLAGS[::2]:
                if P.skipped(df, w, h, dw, dh, flags):
                    continue
                idx, pal = P.picture(w, h, 51)
                rc, want = P.run(refo, idx, pal, w, h, df, dw, dh, flags)
                assert rc == dh
                got = product(sim, 11, [idx, pal.view(np.uint8).reshape(1, 1024)], w, h, df, dw, dh, flags, outs=P.dest(df, dw, dh))
                assert all(np.array_equal(x, y) for x, y in zip(G.crop(df, dw, got), G.crop(df, dw, want))), (df, w, h, dw, dh, hex(flags))
                n += 1
    assert n > 200
    for df in (37, 35):
        assert not sim.sws_getContext_cuda(64, 48, 11, 128, 96, df, 4 | ACC, None, None, None)
        sim.avb200_clear_error()
    # two frames, two palettes, one call
    sim.sws_scale_frames_cuda.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
    for df, (w, h, dw, dh), flags in ((2, (64, 48, 96, 80), 4 | ACC), (26, (64, 48, 64, 48), 4), (0, (66, 50, 33, 25), 2)):
        pics = [P.picture(w, h, 70 + k) for k in range(2)]
        src = np.stack([np.ascontiguousarray(p[0]) for p in pics])
        pals = np.stack([p[1] for p in pics])
        want = [P.run(refo, p[0], p[1], w, h, df, dw, dh, flags)[1] for p in pics]
        outs = [np.stack([np.zeros_like(q) for _ in range(2)]) for q in P.dest(df, dw, dh)]
        ctx = sim.sws_getContext_cuda(w, h, 11, dw, dh, df, flags, None, None, None)
        assert ctx, sim.avb200_last_error()
        sp, ss, sf = (C.c_void_p * 3)(src.ctypes.data, pals.ctypes.data, None), (C.c_int * 3)(src.strides[1], 1024, 0), (C.c_size_t * 3)(src.strides[0], 1024, 0)
        dp = (C.c_void_p * 3)(*([o.ctypes.data for o in outs] + [None] * (3 - len(outs))))
        ds = (C.c_int * 3)(*([o.strides[1] for o in outs] + [0] * (3 - len(outs))))
        dfr = (C.c_size_t * 3)(*([o.strides[0] for o in outs] + [0] * (3 - len(outs))))
        assert sim.sws_scale_frames_cuda(ctx, sp, ss, sf, dp, ds, dfr, 2, None) == 2 * dh, sim.avb200_last_error()
        sim.sws_freeContext_cuda(ctx)
        for k in range(2):
            assert all(np.array_equal(x, y) for x, y in zip(G.crop(df, dw, [o[k] for o in outs]), G.crop(df, dw, want[k]))), (df, k)
    # bottom-up source and destination
    idx, pal = P.picture(64, 48, 77)
    want = product(sim, 11, [idx, pal.view(np.uint8).reshape(1, 1024)], 64, 48, 2, 96, 80, 4)
    up = np.ascontiguousarray(idx[::-1])[::-1]
    store = np.ascontiguousarray(outputs(2, 96, 80)[0][::-1])
    got = [store[::-1]]
    assert up.strides[0] < 0 and got[0].strides[0] < 0
    product(sim, 11, [up, pal.view(np.uint8).reshape(1, 1024)], 64, 48, 2, 96, 80, 4, outs=got)
    same(got, want, "pal8 bottom-up")


def test_gray8_destination(sim, refo):
    """gray8 (tests/test_sws_gray_dst.py): the luma plane of the planar conversion, chroma into the context's scratch"""
    import test_sws_gray_dst as G
    n = 0
    for (sf, w, h, dw, dh, flags) in G.combos():
        pl = G.planes(sf, w, h, 47)
        got = product(sim, sf, pl, w, h, 8, dw, dh, flags)
        want = reference(refo, sf, pl, w, h, 8, dw, dh, flags)
        assert np.array_equal(got[0][:, :dw], want[0][:, :dw]) and (got[0][:, dw:] == 7).all(), (sf, w, h, dw, dh, hex(flags), np.argwhere(got[0] != want[0])[:4].tolist())
        n += 1
    assert n > 200


def test_range_conversion_frames(sim, refo):
    import test_sws_range_cpu as R
    n = 0
    for (sf, df, w, h, dw, dh, flags) in R.cases():
        pl = R.planes(sf, w, h, 21)
        same(product(sim, sf, pl, w, h, df, dw, dh, flags), reference(refo, sf, pl, w, h, df, dw, dh, flags), (sf, df, w, h, dw, dh, hex(flags)))
        n += 1
    assert n == len(R.PAIRS) * len(R.GEOMS) * len(R.FLAGS)


def test_sws_filter_frames(sim, refo):
    import test_sws_filter_cpu as F
    for k, s in enumerate(F.SETS):
        sf, keep1 = F.make_filter(s.get("src"))
        df, keep2 = F.make_filter(s.get("dst"))
        keep = F.set_oracle(refo, s)
        try:
            for (w, h, dw, dh) in F.GEOMS:
                yuv = synth.yuv420p_frame(w, h, 3 + k)
                for flags in (4 | ACC, 4, 2):
                    for dfmt in (2, 0, 28, 5):
                        got = product(sim, 0, yuv, w, h, dfmt, dw, dh, flags, C.byref(sf), C.byref(df))
                        same(got, reference(refo, 0, yuv, w, h, dfmt, dw, dh, flags), (k, dfmt, w, h, dw, dh, hex(flags)))
        finally:
            F.set_oracle(refo, None)
        del keep, keep1, keep2


def test_negative_strides(sim):
    for (sf, df, w, h, dw, dh, flags) in [(0, 2, 64, 48, 96, 80, 4), (0, 2, 128, 64, 128, 64, 4 | ACC), (0, 0, 101, 37, 64, 48, 4), (4, 5, 64, 48, 64, 48, 2),
                                           (1, 2, 64, 48, 96, 80, 4), (0, 28, 64, 48, 33, 25, 4), (0, 23, 64, 48, 96, 80, 4),
                                           (0, 35, 64, 48, 96, 80, 4 | ACC), (0, 60, 64, 48, 64, 48, 4), (8, 2, 64, 48, 96, 80, 4), (8, 0, 64, 48, 64, 48, 4),
                                           (8, 28, 64, 48, 64, 48, 4), (26, 3, 64, 48, 64, 48, 4), (2, 26, 64, 48, 64, 48, 4)]:
        if sf == 8:
            pl = [synth.pad_rows(np.random.RandomState(31).randint(0, 256, (h, w)).astype(np.uint8))]
        elif 25 <= sf <= 28:
            pl = [np.random.RandomState(31).randint(0, 256, (h, 4 * w + 12)).astype(np.uint8)]
        else:
            pl = source(sf, w, h, 31)
        want = product(sim, sf, pl, w, h, df, dw, dh, flags)
        src_up = [np.ascontiguousarray(a[::-1])[::-1] for a in pl]
        store = [np.ascontiguousarray(o[::-1]) for o in outputs(df, dw, dh)]
        got = [g[::-1] for g in store]
        assert all(a.strides[0] < 0 for a in src_up + got)
        product(sim, sf, src_up, w, h, df, dw, dh, flags, outs=got)
        same(got, want, ("bottom-up", sf, df))
        store = [np.ascontiguousarray(o[::-1]) for o in outputs(df, dw, dh)]
        got = [g[::-1] for g in store]
        product(sim, sf, pl, w, h, df, dw, dh, flags, outs=got)
        same(got, want, ("bottom-up destination", sf, df))


def test_range_conversion_with_a_semi_planar_or_packed_side(sim, refo):
    import test_sws_range_cpu as R
    n = 0
    for (sf, df, w, h, dw, dh, flags) in R.mixed_cases():
        pl = R.mixed_planes(sf, w, h, 13)
        rc, want = R.mixed_run(refo, sf, pl, w, h, df, dw, dh, flags)
        assert rc == dh
        got = product(sim, sf, pl, w, h, df, dw, dh, flags, outs=R.mixed_outputs(df, dw, dh))
        same(got, want, (sf, df, w, h, dw, dh, hex(flags)))
        n += 1
    assert n > 230


def test_high_bit_depth_sources(sim, refo):
    """9 / 10 / 16-bit planar sources (hScale16To15 in the two-pass path, ordered dither on 8-bit planar / nv12 outputs)"""
    import test_sws_hbd_sources_cpu as H
    n = 0
    for (sf, df, w, h, dw, dh, flags) in H.cases():
        pl = H.planes(sf, w, h, 5)
        rc, want = H.run(refo, sf, pl, w, h, df, dw, dh, flags)
        assert rc == dh
        got = product(sim, sf, pl, w, h, df, dw, dh, flags, outs=H.outputs(df, dw, dh))
        same(got, want, (sf, df, w, h, dw, dh, hex(flags)))
        n += 1
    assert n > 2000
    # what stays refused, with a reason
    for (sf, df, w, h, dw, dh) in ((64, 64, 64, 48, 64, 48), (64, 0, 64, 48, 64, 48), (64, 47, 64, 48, 96, 80), (64, 12, 64, 48, 96, 80)):
        assert not sim.sws_getContext_cuda(w, h, sf, dw, dh, df, 4, None, None, None)
        assert sim.avb200_last_error()
        sim.avb200_clear_error()
