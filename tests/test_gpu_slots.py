"""GPU: the reference's own tables, filled by ff_*_init_cuda(), called slot by slot with HOST pointers exactly like a codec
(or tests/checkasm) calls them, and byte-compared with the CPU oracle's slot -- outputs AND clobbered inputs (checkasm's
rule, tests/checkasm/h264dsp.c:213-215)."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr
from h264_util import at
import slot_cases

pytestmark = pytest.mark.gpu
u8p, i16p, i8p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int8)


def P(a, off=0, t=u8p):
    return C.cast(a.ctypes.data + int(off), t)


def test_fdct_slots(gpu, checker):
    from libav_b200 import tables
    rng = np.random.default_rng(0)
    for algo, base in ((0, 0), (2, 0), (1, 2)):                 # FF_DCT_AUTO, FF_DCT_INT -> islow ; FF_DCT_FASTINT -> ifast
        c = tables.FDCTDSPContext()
        gpu.lib.ff_fdctdsp_init_cuda(C.byref(c), algo, 8, 0)
        for _ in range(20):
            blk = rng.integers(-256, 256, size=64).astype(np.int16)
            for slot, which in ((c.fdct, base), (c.fdct248, base + 1)):
                a, b = blk.copy(), blk.copy()
                slot(P(a, 0, i16p))
                checker.fdct(which, ptr(b))
                assert np.array_equal(a, b)
    assert gpu.last_error() == ""


def test_me_cmp_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.MECmpContext()
    gpu.lib.ff_me_cmp_init_cuda(C.byref(c))
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=(24, 48), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    cases = [(c.pix_abs[s][d], 0, s, d) for s in (0, 1) for d in range(4)] + \
            [(c.sad[0], 1, 0, 0), (c.sad[1], 1, 1, 0), (c.sse[0], 2, 0, 0), (c.sse[1], 2, 1, 0), (c.sse[2], 2, 2, 0),
             (c.hadamard8_diff[0], 3, 0, 0), (c.hadamard8_diff[1], 3, 1, 0), (c.hadamard8_diff[4], 7, 0, 0), (c.hadamard8_diff[5], 7, 1, 0),
             (c.vsad[0], 4, 0, 0), (c.vsad[4], 8, 0, 0), (c.vsad[5], 8, 1, 0), (c.vsse[0], 5, 0, 0), (c.vsse[4], 9, 0, 0), (c.vsse[5], 9, 1, 0),
             (c.nsse[0], 6, 0, 0), (c.nsse[1], 6, 1, 0)]
    for slot, kind, sidx, dxy in cases:
        w = 16 >> sidx
        h = 8 if w < 16 else 16
        got = slot(None, P(a, 48 * 2 + 16), P(b, 48 * 3 + 5), 48, h)
        assert got == checker.me_cmp(kind, sidx, dxy, at(a, 48 * 2 + 16), at(b, 48 * 3 + 5), 48, h), (kind, sidx, dxy)
    blk = rng.integers(-500, 500, size=64).astype(np.int16)
    assert c.sum_abs_dctelem(P(blk, 0, i16p)) == checker.me_cmp(10, 0, 0, ptr(blk), None, 0, 0)
    assert not c.dct_sad[0] and not c.rd[0] and not c.bit[0]      # encoder-state metrics are not taken over
    assert gpu.last_error() == ""


def test_hpel_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.HpelDSPContext()
    gpu.lib.ff_hpeldsp_init_cuda(C.byref(c), 0)
    rng = np.random.default_rng(2)
    tabs = {0: c.put_pixels_tab, 1: c.avg_pixels_tab, 2: c.put_no_rnd_pixels_tab}
    for tab in range(4):
        for sidx in range(4):
            for dxy in range(4):
                if tab == 3:
                    slot = c.avg_no_rnd_pixels_tab[dxy] if sidx == 0 else None
                else:
                    slot = tabs[tab][sidx][dxy]
                if (tab == 2 and sidx > 1) or (tab == 3 and sidx):
                    assert not slot
                    continue
                w = 16 >> sidx
                h = max(w // 2, 2)
                pix = rng.integers(0, 256, size=(h + 2, 40), dtype=np.uint8)
                blk = rng.integers(0, 256, size=(h, 40), dtype=np.uint8)
                x, y = blk.copy(), blk.copy()
                slot(P(x, 8), P(pix, 3), 40, h)
                assert checker.hpel(tab, sidx, dxy, at(y, 8), at(pix, 3), 40, h) == 0
                assert np.array_equal(x, y), (tab, sidx, dxy)
    assert gpu.last_error() == ""


def test_h264_qpel_and_chroma_slots(gpu, checker):
    slot_cases.qpel_and_chroma_cases(gpu.lib, gpu.last_error, checker)


def test_h264dsp_slots(gpu, checker):
    slot_cases.h264dsp_cases(gpu.lib, gpu.last_error, checker)
