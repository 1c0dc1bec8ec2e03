"""CPU: sws_scale() called slice by slice (libswscale/swscale_unscaled.c:1212-1340, swscale.c:449-496) -- the product's scaler compiled for the
host (tests/hostsim) against the compiled reference given the same slices: the return value of every call and the whole destination picture
after every call (the rows a slice completes, nothing else touched) must agree.  Covers the scaler loop's "enough_lines" progression for up- and
down-scaling, every vertical filter size, sub-sampled chroma on either side, and the row-mapped unscaled converters.

Left out, because the reference's sliced result is not a function of the picture there: vertical down-scaling with SWS_POINT or SWS_FAST_BILINEAR
and slices shorter than the step (its line ring, sized by utils.c:1190-1213 for whole chroma row groups, is over-run while a slice is buffered:
48x64 -> 48x20 fast-bilinear in 4-row slices differs from its own whole-frame picture in rows 6 and 16, 66x50 -> 33x25 point in 7-row slices
crashes; a randomised run also met it with SWS_BILINEAR at 48x64 -> 64x25 in 4-row slices, row 10), and slice boundaries inside a source chroma row (rows before the slice are addressed).  The product returns the whole-frame rows."""
import ctypes as C

import numpy as np
import pytest

from test_hostsim_slots_cpu import sim          # noqa: F401
from test_hostsim_sws_frames_cpu import ACC, arrays, outputs
from libav_b200 import synth
from test_sws_planar_dst import source as _source

VSUB = {61: 1, 47: 1, 48: 1, 66: 0, 70: 0, 0: 1, 4: 0, 5: 0, 23: 1, 24: 1, 12: 1, 13: 0, 14: 0, 6: 2, 7: 0, 8: 1, 62: 1, 63: 1, 64: 1}


def source(sf, w, h, seed):
    r = np.random.RandomState(seed)
    if sf in (62, 61, 64, 63, 47, 48, 66, 70):
        import test_sws_hbd_sources_cpu as H
        return H.planes(sf, w, h, seed)
    if sf in (12, 13, 14):
        import test_sws_range_cpu as R
        return R.planes(sf, w, h, seed)
    if sf == 3:
        return [r.randint(0, 256, (h, 3 * w + 10)).astype(np.uint8)]
    if sf == 11:                                                      # pal8: indices + the palette (which every slice call passes again)
        idx = r.randint(0, 256, (h, w + 16)).astype(np.uint8)
        return [idx, r.randint(0, 256, (1, 1024)).astype(np.uint8)]
    if sf == 8:                                                       # gray8: one plane
        return [synth.pad_rows(r.randint(0, 256, (h, w)).astype(np.uint8))]
    if 25 <= sf <= 28:
        return [r.randint(0, 256, (h, 4 * w + 12)).astype(np.uint8)]
    if sf == 6:
        cw, ch = -((-w) >> 2), -((-h) >> 2)
        return [synth.pad_rows(r.randint(0, 256, d).astype(np.uint8)) for d in ((h, w), (ch, cw), (ch, cw))]
    return _source(sf, w, h, seed)


def slice_planes(sf, pl, y0):
    if sf == 11:
        return [pl[0][y0:], pl[1]]
    vs = VSUB.get(sf, 0)
    return [pl[0][y0:]] + [p[y0 >> vs:] for p in pl[1:]]


def run_slices(scale, ctx, sf, pl, plan, df, dw, dh):
    outs = outputs(df, dw, dh)
    trace, y0 = [], 0
    for sh in plan:
        sp, ss = arrays(slice_planes(sf, pl, y0))
        dp, ds = arrays(outs)
        r = scale(ctx, sp, ss, y0, sh, dp, ds)
        trace.append((r, [o.copy() for o in outs]))
        y0 += sh
    return trace


def plans(h, align):
    out = [[h // 2 // align * align, h - h // 2 // align * align], [16] * (h // 16) + ([h % 16] if h % 16 else [])]
    if align == 1:
        out += [[7] * (h // 7) + ([h % 7] if h % 7 else []), [1] * h]
    else:
        out += [[2 * align] * (h // (2 * align)) + ([h % (2 * align)] if h % (2 * align) else [])]
    return [p for p in out if all(x > 0 for x in p)]


def compare(sim, refo, sf, df, w, h, dw, dh, flags, plan, seed=5):
    pl = source(sf, w, h, seed)
    ctx = sim.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, None, None, None)
    assert ctx, sim.avb200_last_error()
    rc = refo.sws_open(sf, w, h, df, dw, dh, flags)
    assert rc
    try:
        for frame in range(2):                                      # a second picture through the same contexts: the slice state restarts
            got = run_slices(sim.sws_scale_cuda, ctx, sf, pl, plan, df, dw, dh)
            want = run_slices(lambda c, sp, ss, y0, sh, dp, ds: refo.sws_run_slice(c, sp, ss, y0, sh, dp, ds), rc, sf, pl, plan, df, dw, dh)
            assert [g[0] for g in got] == [x[0] for x in want], (sf, df, w, h, dw, dh, hex(flags), plan, [g[0] for g in got], [x[0] for x in want], sim.avb200_last_error())
            assert sum(g[0] for g in got) == dh
            for k, (g, x) in enumerate(zip(got, want)):
                for a, b in zip(g[1], x[1]):
                    a, b = a[:, :a.shape[1] - 8], b[:, :b.shape[1] - 8]
                    assert np.array_equal(a, b), (sf, df, w, h, dw, dh, hex(flags), plan, k, np.argwhere(a != b)[:4].tolist())
    finally:
        sim.sws_freeContext_cuda(ctx)
        refo.sws_close(rc)


def test_scaler_loop_slices(sim, refo):
    n = 0
    for sf, df in ((0, 2), (0, 0), (0, 5), (4, 0), (5, 28), (0, 23), (23, 0), (2, 0), (1, 3), (0, 1), (0, 63)):
        for (w, h, dw, dh) in ((64, 48, 96, 80), (66, 50, 33, 25), (64, 48, 64, 48), (48, 64, 48, 20)):
            for flags in (4 | ACC, 2, 0x10 | ACC, 1):
                if (w, h) == (dw, dh) and not (flags & ACC == ACC and sf in (0, 4, 5) and df in (2, 28)):
                    continue                                         # same-size pairs mostly install an unscaled converter: next test
                if flags in (1, 0x10 | ACC) and dh < h:
                    continue                                         # see the module docstring: the reference's own sliced output is not its whole-frame output here
                if flags == 1 and dh > h:
                    continue                                         # fast-bilinear up-scaling reads the reference's uncleared conversion buffer
                for plan in plans(h, 1 << VSUB.get(sf, 0)):
                    compare(sim, refo, sf, df, w, h, dw, dh, flags, plan)
                    n += 1
    assert n > 200


def test_range_conversion_and_deep_source_slices(sim, refo):
    """the two-pass-only stages under slices: yuvj <-> yuv range conversion on the scaled lines, 9 / 10 / 16-bit sources (hScale16To15_c)"""
    import test_sws_hbd_sources_cpu as H
    n = 0
    for sf, df in ((12, 0), (0, 12), (14, 4), (12, 2), (64, 0), (62, 2), (47, 5), (63, 28), (64, 62)):
        for (w, h, dw, dh) in ((64, 48, 96, 80), (66, 50, 33, 25), (64, 48, 64, 48)):
            for flags in (4 | ACC, 2):
                if sf > 40 and not H.taken(sf, df, w, h, dw, dh, flags):
                    continue                                         # (their same-size plane copies are refused at sws_getContext_cuda)
                for plan in plans(h, 1 << VSUB.get(sf, 0)):
                    compare(sim, refo, sf, df, w, h, dw, dh, flags, plan)
                    n += 1
    assert n > 140


def test_unscaled_converter_slices(sim, refo):
    n = 0
    for sf, df, flags, align in ((0, 0, 4, 2), (4, 4, 4, 1), (0, 2, 4, 2), (4, 3, 4, 2), (0, 23, 4, 2), (23, 0, 4, 2), (2, 3, 4, 1), (3, 0, 4, 2), (1, 0, 4, 2),
                                 (15, 4, 4, 1), (4, 1, 4, 1), (0, 1, 1, 2), (1, 1, 4, 1), (6, 6, 4, 4), (0, 62, 4, 2)):
        for (w, h) in ((64, 48), (66, 52)):
            for plan in plans(h, align):
                compare(sim, refo, sf, df, w, h, w, h, flags, plan)
                n += 1
    assert n >= 80


def test_slices_of_the_late_formats(sim, refo):
    """rgb48 destinations, gray8 sources and the same-size rgb2rgb converters under slices"""
    n = 0
    for sf, df in ((0, 35), (5, 60), (8, 2), (8, 0), (8, 37), (11, 2), (11, 0), (11, 28)):
        for (w, h, dw, dh) in ((64, 48, 96, 80), (66, 50, 33, 25)):
            for flags in (4 | ACC, 2):
                for plan in plans(h, 1 << (VSUB.get(sf, 0) if sf not in (8, 11) else 0)):
                    compare(sim, refo, sf, df, w, h, dw, dh, flags, plan)
                    n += 1
    for sf, df, flags, align in ((0, 35, 4, 2), (4, 59, 4, 2), (8, 0, 4, 2), (8, 5, 4, 1), (8, 26, 4, 1), (8, 3, 4, 1), (26, 2, 4, 1), (2, 28, 4, 1), (25, 27, 4, 1), (11, 3, 4, 1), (11, 25, 4, 1)):
        for (w, h) in ((64, 48), (66, 52)):
            for plan in plans(h, align):
                compare(sim, refo, sf, df, w, h, w, h, flags, plan)
                n += 1
    assert n > 100


def test_slice_refusals(sim):
    pl = source(0, 64, 48, 1)
    ctx = sim.sws_getContext_cuda(64, 48, 0, 96, 80, 8, 4, None, None, None)             # gray8 destination: whole frames only
    sp, ss = arrays(slice_planes(0, pl, 0))
    gray = np.zeros((80, 96), np.uint8)
    dp, ds = arrays([gray])
    sim.avb200_clear_error()
    assert sim.sws_scale_cuda(ctx, sp, ss, 0, 16, dp, ds) == 0 and b"gray8 destination" in sim.avb200_last_error()
    assert sim.sws_scale_cuda(ctx, sp, ss, 0, 48, dp, ds) == 80
    sim.sws_freeContext_cuda(ctx)
    sim.avb200_clear_error()
    ctx = sim.sws_getContext_cuda(64, 48, 0, 96, 80, 2, 4, None, None, None)
    outs = outputs(2, 96, 80)
    dp, ds = arrays(outs)
    try:
        sp, ss = arrays(slice_planes(0, pl, 16))
        sim.avb200_clear_error()
        assert sim.sws_scale_cuda(ctx, sp, ss, 16, 16, dp, ds) == 0 and b"top-down" in sim.avb200_last_error()      # starts in the middle
        sp, ss = arrays(slice_planes(0, pl, 0))
        assert sim.sws_scale_cuda(ctx, sp, ss, 0, 16, dp, ds) > 0
        sp, ss = arrays(slice_planes(0, pl, 32))
        assert sim.sws_scale_cuda(ctx, sp, ss, 32, 16, dp, ds) == 0                                                     # a gap
        sim.avb200_clear_error()
        assert sim.sws_scale_cuda(ctx, sp, ss, 40, 16, dp, ds) == 0 and b"outside" in sim.avb200_last_error()
    finally:
        sim.sws_freeContext_cuda(ctx)
    sim.avb200_clear_error()
    ctx = sim.sws_getContext_cuda(64, 48, 0, 64, 48, 0, 4, None, None, None)
    outs = outputs(0, 64, 48)
    dp, ds = arrays(outs)
    try:
        sp, ss = arrays(slice_planes(0, pl, 0))
        assert sim.sws_scale_cuda(ctx, sp, ss, 0, 7, dp, ds) == 0 and b"aligned" in sim.avb200_last_error()
    finally:
        sim.sws_freeContext_cuda(ctx)
    sim.avb200_clear_error()
    ctx = sim.sws_getContext_cuda(64, 48, 0, 96, 80, 2, 4, None, None, None)                  # the scaler loop: whole source chroma rows
    outs = outputs(2, 96, 80)
    dp, ds = arrays(outs)
    try:
        assert sim.sws_scale_cuda(ctx, sp, ss, 0, 7, dp, ds) == 0 and b"aligned" in sim.avb200_last_error()
    finally:
        sim.sws_freeContext_cuda(ctx)
        sim.avb200_clear_error()
