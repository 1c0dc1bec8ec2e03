"""SwsFilter (srcFilter / dstFilter of sws_getContext: caller-made vectors that initFilter() folds into the filter banks, utils.c:444-474):
the banks the PRODUCT's host set-up designs (sws_debug_filter2_cuda, no GPU involved) and the port's, against the compiled reference; whole
frames port vs reference.  Vectors: a normalised blur (what sws_getDefaultFilter builds for lumaGBlur), a sharpen kernel, a shifted identity."""
import ctypes as C

import numpy as np
import pytest

import libav_b200._lib as L
from oracle.loader import ptr
from test_oracle_sws_cpu import ACC, _bank, to_rgb, to_yuv
from libav_b200 import synth

BLUR = np.array([0.05, 0.24, 0.42, 0.24, 0.05])
SHARP = np.array([-0.1, 1.2, -0.1])
SHIFT = np.array([0.0, 0.0, 1.0])
ONE = np.array([1.0])
# (lumH, lumV, chrH, chrV) for the source side; the destination side only ever widens the rows
G21 = np.exp(-np.linspace(-2.5, 2.5, 21) ** 2) / np.exp(-np.linspace(-2.5, 2.5, 21) ** 2).sum()
SETS = [dict(src=(BLUR, BLUR, None, None)), dict(src=(SHARP, None, BLUR, SHARP)), dict(src=(SHIFT, None, SHIFT, G21), dst=(BLUR, None, None, SHARP)),
        dict(src=(ONE, ONE, ONE, ONE))]
GEOMS = [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (100, 37, 333, 211)]


class Vec(C.Structure):
    _fields_ = [("coeff", C.POINTER(C.c_double)), ("length", C.c_int)]


class Filt(C.Structure):
    _fields_ = [("lumH", C.POINTER(Vec)), ("lumV", C.POINTER(Vec)), ("chrH", C.POINTER(Vec)), ("chrV", C.POINTER(Vec))]


def make_filter(vectors):
    """a reference-layout SwsFilter (plus the objects that must stay alive)"""
    keep, f = [], Filt()
    for name, v in zip(("lumH", "lumV", "chrH", "chrV"), vectors or (None,) * 4):
        if v is None:
            continue
        arr = np.ascontiguousarray(v, np.float64)
        vec = Vec(arr.ctypes.data_as(C.POINTER(C.c_double)), len(arr))
        keep += [arr, vec]
        setattr(f, name, C.pointer(vec))
    return f, keep


def set_oracle(o, s):
    keep = []
    for side, key in ((0, "src"), (1, "dst")):
        for which, v in enumerate((s or {}).get(key) or (None,) * 4):
            arr = None if v is None else np.ascontiguousarray(v, np.float64)
            keep.append(arr)
            o.sws_set_filter(which, side, None if arr is None else arr.ctypes.data, 0 if arr is None else len(arr))
    return keep


@pytest.mark.parametrize("k", range(len(SETS)))
def test_banks_match_reference(refo, orc, k):
    s = SETS[k]
    sf, keep1 = make_filter(s.get("src"))
    df, keep2 = make_filter(s.get("dst"))
    keep = [set_oracle(o, s) for o in (refo, orc)]
    try:
        for flags in (4 | ACC, 2, 0x200 | ACC, 0x10):
            for rgb in (1, 0):
                for (sw, sh, dw, dh) in GEOMS:
                    for which in range(4):
                        a = _bank(refo.sws_get_filter, which, rgb, sw, sh, dw, dh, flags)
                        b = _bank(orc.sws_get_filter, which, rgb, sw, sh, dw, dh, flags)
                        c = _bank(lambda *args: L.lib.sws_debug_filter2_cuda(*args[:7], C.byref(sf), C.byref(df), *args[7:]), which, sw, sh, dw, dh, 2 if rgb else 0, flags)
                        L.lib.avb200_clear_error()
                        if a[0] == -3:
                            continue
                        for other, who in ((b, "port"), (c, "product")):
                            assert a[0] == other[0], (who, k, hex(flags), rgb, sw, sh, dw, dh, which, a[0], other[0])
                            if a[0] > 0:
                                assert np.array_equal(a[1], other[1]) and np.array_equal(a[2], other[2]), (who, k, hex(flags), rgb, sw, sh, dw, dh, which)
    finally:
        for o in (refo, orc):
            set_oracle(o, None)
    del keep, keep1, keep2


@pytest.mark.parametrize("k", range(len(SETS)))
def test_frames_port_matches_reference(refo, orc, k):
    keep = [set_oracle(o, SETS[k]) for o in (refo, orc)]
    try:
        for (sw, sh, dw, dh) in GEOMS:
            yuv = synth.yuv420p_frame(sw, sh, 3 + k)
            for flags in (4 | ACC, 4, 2):
                a, b = to_rgb(refo, yuv, dw, dh, flags, pad=3), to_rgb(orc, yuv, dw, dh, flags, pad=3)
                assert a[0] == b[0] == dh and np.array_equal(a[1], b[1]), ("rgb", k, sw, sh, dw, dh, hex(flags))
                a, b = to_yuv(refo, yuv, dw, dh, flags), to_yuv(orc, yuv, dw, dh, flags)
                assert a[0] == b[0] == dh and all(np.array_equal(x, y) for x, y in zip(a[1], b[1])), ("yuv", k, sw, sh, dw, dh, hex(flags))
    finally:
        for o in (refo, orc):
            set_oracle(o, None)
    del keep


def test_filters_change_the_picture_and_length_one_vectors_do_not(refo):
    yuv = synth.yuv420p_frame(64, 48, 1)
    plain = to_rgb(refo, yuv, 64, 48, 4 | ACC)[1]
    keep = set_oracle(refo, SETS[0])
    blurred = to_rgb(refo, yuv, 64, 48, 4 | ACC)[1]
    set_oracle(refo, SETS[3])
    same = to_rgb(refo, yuv, 64, 48, 4 | ACC)[1]
    set_oracle(refo, None)
    assert not np.array_equal(plain, blurred) and np.array_equal(plain, same)
    del keep


def test_product_refuses_what_it_does_not_take(built):
    sf, keep = make_filter((BLUR, None, None, None))
    out = (C.c_int32 * 8)()
    # (no device: sws_getContext_cuda itself needs one; the decision is the host code of make_context, reached through the bank probe)
    n = C.c_int()
    buf, pos = np.zeros(1 << 16, np.int16), np.zeros(1 << 12, np.int32)
    assert L.lib.sws_debug_filter2_cuda(0, 64, 48, 96, 80, 23, 4, C.byref(sf), None, buf.ctypes.data, pos.ctypes.data, 1 << 12, C.byref(n)) < 0      # nv12 destination
    assert "SwsFilter" in L.last_error()
    L.lib.avb200_clear_error()
    assert L.lib.sws_debug_filter2_cuda(0, 64, 48, 96, 80, 2, 4, C.byref(sf), None, buf.ctypes.data, pos.ctypes.data, 1 << 12, C.byref(n)) > 0
    sv, keep2 = make_filter((None, SHIFT, None, None))                # a shifted vertical vector: the reference's bottom rows depend on its ring buffer
    assert L.lib.sws_debug_filter2_cuda(0, 64, 48, 96, 80, 2, 4, C.byref(sv), None, buf.ctypes.data, pos.ctypes.data, 1 << 12, C.byref(n)) < 0
    assert "asymmetric" in L.last_error()
    L.lib.avb200_clear_error()
    del keep2
    del keep, out
