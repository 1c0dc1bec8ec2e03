"""The product's per-line SwsContext slots installed in a REAL reference SwsContext and driven by the reference's own, unmodified line
scheduler swscale() (libswscale/swscale.c:343-721): sws_getContext() -> ff_sws_init_swscale_cuda(c, ...) -> the slot fields of `c` are
replaced (oracle/refbuild/refapi.c ref_sws_set_slots, the assignment INTEGRATION.md section 4 shows) -> sws_scale().  The picture must be
the one the untouched reference produces.  Shared by the host simulation (CPU) and the product library (GPU).  Needs oracle/_ref (the
compiled reference); TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from libav_b200.device import PLANAR_BE, PLANAR_FORMATS
from test_sws_planar_dst import source

ACC = 0x40000 | 0x80000
# (source format, destination format, flags): every slot family gets work -- hScale8To15 / 19, the fast-bilinear pair, plane1 / planeX at
# 8 / 10 / 16 bit, nv12, packed 1 / 2 / X for rgb and 4:2:2, the full-chroma X function, both range conversions
CASES = [(0, 2, 4 | ACC), (0, 3, 2), (0, 28, 4), (0, 25, 0x10), (0, 2, 4 | ACC | 0x2000), (0, 1, 4), (0, 15, 2 | ACC), (0, 0, 4 | ACC), (0, 0, 1 | ACC),
         (0, 2, 1), (4, 5, 0x200 | ACC), (0, 64, 4), (0, 63, 2), (0, 47, 4 | ACC), (0, 23, 4), (0, 24, 2), (12, 0, 4), (0, 12, 2 | ACC), (5, 2, 0x400)]
GEOMS = [(64, 48, 96, 80), (96, 80, 64, 48), (66, 50, 66, 50), (64, 48, 64, 80), (64, 48, 96, 48)]


def outputs(df, dw, dh):
    if df in (23, 24):
        return [np.full((dh, dw + 16), 7, np.uint8), np.full(((dh + 1) // 2, 2 * ((dw + 1) // 2) + 16), 7, np.uint8)]
    if df in PLANAR_FORMATS:
        hs, vs, bits = PLANAR_FORMATS[df]
        dt = np.uint8 if bits == 8 else np.dtype(">u2" if df in PLANAR_BE else "<u2")
        cw, ch = -((-dw) >> hs), -((-dh) >> vs)
        return [np.full((dh, dw + 16), 7, dt), np.full((ch, cw + 16), 7, dt), np.full((ch, cw + 16), 7, dt)]
    return [np.full((dh, 4 * dw + 32), 7, np.uint8)]


def run(refo, ctx, pl, h, outs):
    sp = (C.c_void_p * 3)(*([a.ctypes.data for a in pl] + [None] * (3 - len(pl))))
    ss = (C.c_int * 3)(*([a.strides[0] for a in pl] + [0] * (3 - len(pl))))
    dp = (C.c_void_p * 3)(*([a.ctypes.data for a in outs] + [None] * (3 - len(outs))))
    ds = (C.c_int * 3)(*([a.strides[0] for a in outs] + [0] * (3 - len(outs))))
    return refo.sws_run(ctx, sp, ss, h, dp, ds)


def check(refo, lib, make_ctx, free_ctx, cases=CASES, geoms=GEOMS):
    """make_ctx(src_fmt, w, h, dst_fmt, dw, dh, flags) -> the product-side SwsContextCUDA for the same request"""
    from libav_b200 import tables
    n = 0
    for (sf, df, flags) in cases:
        for (w, h, dw, dh) in geoms:
            base = {12: 0, 13: 4, 14: 5}.get(sf, sf)
            pl = source(base, w, h, 77)
            plain = refo.sws_open(sf, w, h, df, dw, dh, flags)
            assert plain
            want = outputs(df, dw, dh)
            assert run(refo, plain, pl, h, want) == dh
            mask = refo.sws_slot_mask(plain)
            refo.sws_close(plain)
            if not (mask & 1):
                continue                                   # an unscaled special converter: swscale() and its slots are not used at all
            c = refo.sws_open(sf, w, h, df, dw, dh, flags)
            cuda = make_ctx(sf, w, h, df, dw, dh, flags)
            t = tables.SwsLineSlotsCUDA()
            assert lib.ff_sws_init_swscale_cuda(C.c_void_p(c), C.c_void_p(cuda), C.byref(t)) == 0
            mine = sum(1 << i for i, (name, _) in enumerate(tables.SwsLineSlotsCUDA._fields_) if C.cast(getattr(t, name), C.c_void_p).value)
            assert mine == mask, (sf, df, hex(flags), bin(mine), bin(mask))          # the same set of slots as the reference installs
            assert refo.sws_set_slots(c, C.byref(t)) == bin(mask).count("1")
            got = outputs(df, dw, dh)
            assert run(refo, c, pl, h, got) == dh
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (sf, df, hex(flags), w, h, dw, dh, np.argwhere(a != b)[:4].tolist())
            refo.sws_close(c)
            free_ctx(cuda)
            n += 1
    return n
