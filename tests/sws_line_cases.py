"""Cases for libswscale's per-line SwsContext slots, shared by the oracle pinning (port vs compiled reference), the host simulation
and the GPU tests: caller-made lines and coefficients shaped like the ones swscale() hands to the slots (15-bit lines from 8-bit
samples, 14-bit horizontal / 12-bit vertical coefficients with negative lobes), every function compared byte for byte with the
function the reference itself installs for that destination format.  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from oracle.loader import ptr

FULL, FAST, BICUBIC = 0x2000, 1, 4
PLANE_FMTS = [0, 62, 61, 64, 63, 47, 48]            # yuv420p 8 bit, 9 LE / BE, 10 LE / BE, 16 LE / BE
PACKED_FMTS = [2, 3, 1, 15, 25, 26, 27, 28]         # rgb24 bgr24 yuyv422 uyvy422 argb rgba abgr bgra
ITU709 = (117504, 138453, 13954, 34903)
FCC = (104448, 132798, 24759, 53109)          # with brightness 3000, contrast 70000, saturation 60000: inside the reference's colour table


def vp(a, off=0):
    return C.c_void_p(a.ctypes.data + off)


def ptrs(lines):
    return (C.c_void_p * len(lines))(*[l.ctypes.data for l in lines])


def coeffs(rng, n, one):
    """n taps summing to `one`; from four taps up the outer two are negative lobes (-8 % and -5 %) like the designed filters'.
    The sums stay inside (-256, 512) for 8-bit samples: beyond that the reference indexes outside its colour tables (the
    "clip only when bit 8 is set" rule of output.c:966-971 does not catch such values)."""
    if n == 1:
        return np.array([one], np.int16)
    w = rng.random(n) + 0.2
    if n >= 4:
        w[0] = w[-1] = 0
        w *= 1.13 / w.sum()
        w[0], w[-1] = -0.08, -0.05
    else:
        w /= w.sum()
    w = np.round(w * one).astype(np.int64)
    w[n // 2] += one - w.sum()
    return w.astype(np.int16)


def lines15(rng, n, width, wide=False):
    """hscaled lines: samples << 7 (<< 11 for the 19-bit lines of 16-bit destinations) plus noise, a few overshoots"""
    sh = 11 if wide else 7
    top = (1 << (sh + 8)) - 1
    out = []
    for _ in range(n):
        l = (rng.integers(0, 256, size=width + 8).astype(np.int64) << sh) + rng.integers(-64, 64, size=width + 8)
        l[rng.integers(0, width, size=3)] = top
        l[rng.integers(0, width, size=3)] = -rng.integers(1, 300)
        out.append(np.clip(l, -2000, top).astype(np.int32 if wide else np.int16))
    return out


def hscale_cases(rng):
    for dst_fmt in (0, 47):
        for (src_w, dst_w, fs) in ((64, 37, 1), (200, 128, 4), (333, 333, 2), (90, 250, 8), (500, 121, 11)):
            src = rng.integers(0, 256, size=src_w + 16, dtype=np.uint8)
            pos = np.sort(rng.integers(0, src_w - fs + 1, size=dst_w)).astype(np.int32)
            if dst_w > 4:
                pos[:2] = pos[:2][::-1]                     # positions need not be monotonic for the function itself
            filt = np.concatenate([coeffs(rng, fs, 1 << 14) for _ in range(dst_w)]).astype(np.int16)
            yield dst_fmt, dst_w, src, filt, pos, fs


def run_hscale(call, case):
    dst_fmt, dst_w, src, filt, pos, fs = case
    out = np.full(dst_w + 4, -7, np.int32 if dst_fmt == 47 else np.int16)
    call(dst_fmt, out, dst_w, src, filt, pos, fs)
    return out


def hfast_cases(rng):
    for (src_w, dst_w) in ((100, 173), (64, 64), (321, 200), (17, 300)):
        x_inc = ((src_w << 16) + (dst_w >> 1)) // dst_w
        yield src_w, dst_w, x_inc, rng.integers(0, 256, size=src_w + 8, dtype=np.uint8), rng.integers(0, 256, size=src_w + 8, dtype=np.uint8)


def plane_cases(rng):
    for fmt in PLANE_FMTS:
        wide = fmt in (47, 48)
        for fs in (0, 1, 2, 4, 9):
            for width in (45, 128):
                yield fmt, fs, width, lines15(rng, max(fs, 1), width, wide), coeffs(rng, max(fs, 1), 1 << 12), \
                    rng.integers(0, 128, size=8, dtype=np.uint8), int(rng.integers(0, 2)) * 3


def run_plane(call, case):
    fmt, fs, width, lines, filt, dither, offset = case
    out = np.full(2 * width + 8, 0xA5, np.uint8)
    call(fmt, filt, fs, lines, out, width, dither, offset)
    return out


def nv12_cases(rng):
    for fmt in (23, 24):
        for fs in (1, 2, 4, 7):
            for width in (33, 96):
                yield fmt, fs, width, lines15(rng, fs, width), lines15(rng, fs, width), coeffs(rng, fs, 1 << 12)


def packed_cases(rng):
    for fmt in PACKED_FMTS:
        for flags in (BICUBIC, BICUBIC | FULL):
            full = bool(flags & FULL) and fmt not in (1, 15)
            if full and fmt == 27:
                continue                    # abgr + SWS_FULL_CHR_H_INT: the reference advances its pointer twice per pixel (output.c:1231-1237); refused by the product
            for kind in ((0,) if full else (1, 2, 0)):
                for width in (46, 77):
                    nl = 1 if kind == 1 else 2 if kind == 2 else int(rng.choice([1, 2, 4, 6]))
                    nc = 2 if kind else int(rng.choice([1, 2, 4, 5]))
                    yalpha, uvalpha = int(rng.integers(0, 4096)), int(rng.choice([0, 1500, 2048, 3000])) if kind == 1 else int(rng.integers(0, 4096))
                    yield fmt, flags, kind, width, lines15(rng, nl, width), lines15(rng, nc, width), lines15(rng, nc, width), \
                        coeffs(rng, nl, 1 << 12), coeffs(rng, nc, 1 << 12), yalpha, uvalpha


def run_packed(call, case):
    fmt, flags, kind, width, lum, cu, cv, lf, cf, yalpha, uvalpha = case
    out = np.full(4 * (width + 2) + 8, 0xA5, np.uint8)
    call(fmt, flags, kind, lf, lum, cf, cu, cv, out, width, yalpha, uvalpha)
    return out


def range_cases(rng):
    for kind in range(4):
        for width in (37, 200):
            a = rng.integers(-300, 32768, size=width + 4).astype(np.int16)
            a[:3] = (32767, 30189, 30775)
            yield kind, width, a, rng.integers(-300, 32768, size=width + 4).astype(np.int16)


# ---- the two kinds of callee -------------------------------------------------------------------------------------------------
class OracleCalls:
    """the oracle entry points (oracle/oracle_api.h sws_line_*)"""

    def __init__(self, o):
        self.o = o

    def hscale(self, dst_fmt, out, dst_w, src, filt, pos, fs):
        assert self.o.sws_line_hscale(dst_fmt, BICUBIC, ptr(out), dst_w, ptr(src), ptr(filt), ptr(pos), fs) == 0

    def hfast(self, chroma, d1, d2, dst_w, s1, s2, src_w, x_inc):
        assert self.o.sws_line_hfast(chroma, ptr(d1), ptr(d2), dst_w, ptr(s1), ptr(s2), src_w, x_inc) == 0

    def plane(self, fmt, filt, fs, lines, out, width, dither, offset):
        assert self.o.sws_line_plane(fmt, ptr(filt), fs, ptrs(lines), ptr(out), width, ptr(dither), offset) == 0

    def nv12(self, fmt, cf, fs, cu, cv, out, width):
        assert self.o.sws_line_nv12(fmt, ptr(cf), fs, ptrs(cu), ptrs(cv), ptr(out), width) == 0

    def range(self, kind, d1, d2, width):
        assert self.o.sws_line_range(kind, ptr(d1), ptr(d2), width) == 0

    def packed(self, fmt, flags, kind, lf, lum, cf, cu, cv, out, width, yalpha, uvalpha):
        assert self.o.sws_line_packed(fmt, flags, kind, ptr(lf), ptrs(lum), len(lum), ptr(cf), ptrs(cu), ptrs(cv), len(cu), ptr(out), width,
                                      yalpha, uvalpha, 5) == 0


class SlotCalls:
    """the product's slots: make_ctx(dst_fmt, flags) -> (SwsContextCUDA *, free); init = ff_sws_init_swscale_cuda of the library under test"""

    def __init__(self, lib, make_ctx, last_error):
        self.lib, self.make_ctx, self.last_error = lib, make_ctx, last_error
        self.cache = {}

    def slots(self, dst_fmt, flags, src_fmt=0):
        from libav_b200 import tables
        key = (dst_fmt, flags, src_fmt)
        if key not in self.cache:
            ctx, free = self.make_ctx(dst_fmt, flags) if not src_fmt else self.make_ctx(dst_fmt, flags, src_fmt)
            t = tables.SwsLineSlotsCUDA()
            handle = C.c_void_p(0x5000 + 16 * len(self.cache))         # stands for the caller's struct SwsContext *
            assert self.lib.ff_sws_init_swscale_cuda(handle, C.c_void_p(ctx), C.byref(t)) == 0
            self.cache[key] = (handle, t, ctx, free)
        return self.cache[key][:2]

    def close(self):
        for (_, _, ctx, free) in self.cache.values():
            free(ctx)
        self.cache = {}

    def hscale(self, dst_fmt, out, dst_w, src, filt, pos, fs):
        h, t = self.slots(dst_fmt, BICUBIC)
        t.hyScale(h, vp(out), dst_w, vp(src), vp(filt), vp(pos), fs)
        assert C.cast(t.hcScale, C.c_void_p).value == C.cast(t.hyScale, C.c_void_p).value and not t.hyscale_fast and not t.lumConvertRange

    def hfast(self, chroma, d1, d2, dst_w, s1, s2, src_w, x_inc):
        h, t = self.slots(0, FAST)
        if chroma:
            t.hcscale_fast(h, vp(d1), vp(d2), dst_w, vp(s1), vp(s2), src_w, x_inc)
        else:
            t.hyscale_fast(h, vp(d1), dst_w, vp(s1), src_w, x_inc)

    def plane(self, fmt, filt, fs, lines, out, width, dither, offset):
        h, t = self.slots(fmt, BICUBIC)
        if fs:
            t.yuv2planeX(vp(filt), fs, ptrs(lines), vp(out), width, vp(dither), offset)
        else:
            t.yuv2plane1(vp(lines[0]), vp(out), width, vp(dither), offset)
        assert not t.yuv2packedX and not t.yuv2nv12cX

    def nv12(self, fmt, cf, fs, cu, cv, out, width):
        h, t = self.slots(fmt, BICUBIC)
        t.yuv2nv12cX(h, vp(cf), fs, ptrs(cu), ptrs(cv), vp(out), width)

    def range(self, kind, d1, d2, width):
        h, t = self.slots(0 if kind < 2 else 12, BICUBIC, src_fmt=12 if kind < 2 else 0)       # yuvj420p -> yuv420p / yuv420p -> yuvj420p
        if kind & 1:
            t.chrConvertRange(vp(d1), vp(d2), width)
        else:
            t.lumConvertRange(vp(d1), width)

    def packed(self, fmt, flags, kind, lf, lum, cf, cu, cv, out, width, yalpha, uvalpha):
        h, t = self.slots(fmt, flags)
        if (flags & FULL) and fmt not in (1, 15):
            assert not t.yuv2packed1 and not t.yuv2packed2          # output.c:1392-1472
        if kind == 1:
            t.yuv2packed1(h, vp(lum[0]), ptrs(cu), ptrs(cv), None, vp(out), width, uvalpha, 5)
        elif kind == 2:
            t.yuv2packed2(h, ptrs(lum), ptrs(cu), ptrs(cv), None, vp(out), width, yalpha, uvalpha, 5)
        else:
            t.yuv2packedX(h, vp(lf), ptrs(lum), len(lum), vp(cf), ptrs(cu), ptrs(cv), len(cu), None, vp(out), width, 5)


def compare(a, b, seed=0, colourspace=None):
    """every case through callee a and callee b; returns the number of comparisons"""
    n = 0
    rng = np.random.default_rng(seed)
    for case in hscale_cases(rng):
        x, y = run_hscale(a.hscale, case), run_hscale(b.hscale, case)
        assert np.array_equal(x, y), ("hscale", case[0], case[1], case[5]); n += 1
    for (src_w, dst_w, x_inc, s1, s2) in hfast_cases(rng):
        for chroma in (0, 1):
            outs = []
            for c in (a, b):
                d1, d2 = np.full(dst_w + 4, -7, np.int16), np.full(dst_w + 4, -7, np.int16)
                c.hfast(chroma, d1, d2, dst_w, s1, s2, src_w, x_inc)
                outs.append((d1, d2))
            assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), ("hfast", chroma, src_w, dst_w); n += 1
    for case in plane_cases(rng):
        x, y = run_plane(a.plane, case), run_plane(b.plane, case)
        assert np.array_equal(x, y), ("plane",) + case[:3]; n += 1
    for (fmt, fs, width, cu, cv, cf) in nv12_cases(rng):
        outs = []
        for c in (a, b):
            out = np.full(2 * width + 8, 0xA5, np.uint8)
            c.nv12(fmt, cf, fs, cu, cv, out, width)
            outs.append(out)
        assert np.array_equal(outs[0], outs[1]), ("nv12", fmt, fs, width); n += 1
    for (kind, width, l1, l2) in range_cases(rng):
        outs = []
        for c in (a, b):
            d1, d2 = l1.copy(), l2.copy()
            c.range(kind, d1, d2, width)
            outs.append((d1, d2))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), ("range", kind, width)
        assert np.array_equal(outs[0][0][width:], l1[width:]) and ((kind & 1) or np.array_equal(outs[0][1], l2)); n += 1
    for case in packed_cases(rng):
        x, y = run_packed(a.packed, case), run_packed(b.packed, case)
        assert np.array_equal(x, y), ("packed",) + case[:4] + (np.argwhere(x != y)[:4].tolist(),); n += 1
    return n
