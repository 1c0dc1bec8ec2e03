"""QpelDSPContext (MPEG-4 quarter-pel MC): the one-formula restatement against the reference's 96 macro-generated
functions (CPU); the GPU table slots and the record batch against the checker."""
import ctypes as C

import numpy as np
import pytest

from oracle.loader import ptr

STRIDE = 48


def at(a, off):
    return a.ctypes.data + int(off)


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = np.random.RandomState(3)
    for it in range(12):
        src = r.randint(0, 256, (40, STRIDE)).astype(np.uint8)
        if it % 4 == 0:
            src[:] = r.choice([0, 255]); src[::2, ::3] = r.randint(0, 256)      # overshoot of the 8-tap filter -> clipping
        d0 = r.randint(0, 256, (40, STRIDE)).astype(np.uint8)
        for kind in range(3):
            for sidx in range(2):
                for mc in range(16):
                    a, b = d0.copy(), d0.copy()
                    refo.mpeg4_qpel(kind, sidx, mc, at(a, 4 * STRIDE + 8), at(src, 4 * STRIDE + 8), STRIDE)
                    orc.mpeg4_qpel(kind, sidx, mc, at(b, 4 * STRIDE + 8), at(src, 4 * STRIDE + 8), STRIDE)
                    assert np.array_equal(a, b), (kind, sidx, mc, np.argwhere(a != b)[:3].tolist())


@pytest.mark.gpu
def test_table_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.QpelDSPContext()
    gpu.lib.ff_qpeldsp_init_cuda(C.byref(c))
    tabs = [c.put_qpel_pixels_tab, c.put_no_rnd_qpel_pixels_tab, c.avg_qpel_pixels_tab]
    r = np.random.RandomState(4)
    for it in range(2):
        src = r.randint(0, 256, (40, STRIDE)).astype(np.uint8)
        if it:
            src[:] = 255; src[::2, ::3] = 0
        d0 = r.randint(0, 256, (40, STRIDE)).astype(np.uint8)
        for kind in range(3):
            for sidx in range(2):
                for mc in range(16):
                    want, got = d0.copy(), d0.copy()
                    checker.mpeg4_qpel(kind, sidx, mc, at(want, 4 * STRIDE + 8), at(src, 4 * STRIDE + 8), STRIDE)
                    p8 = lambda a: C.cast(a, C.POINTER(C.c_uint8))
                    tabs[kind][sidx][mc](p8(at(got, 4 * STRIDE + 8)), p8(at(src, 4 * STRIDE + 8)), STRIDE)
                    assert gpu.last_error() == ""
                    assert np.array_equal(got, want), (kind, sidx, mc, np.argwhere(got != want)[:3].tolist())


@pytest.mark.gpu
def test_record_batch(gpu, checker):
    from libav_b200 import device
    r = np.random.RandomState(5)
    w, h = 352, 288
    src = r.randint(0, 256, (h + 1, w + 8)).astype(np.uint8)
    base = r.randint(0, 256, (h, w + 8)).astype(np.uint8)
    st = src.strides[0]
    dt = np.dtype([("dst_off", "<u4"), ("src_off", "<u4"), ("kind", "u1"), ("sidx", "u1"), ("mc", "u1"), ("pad", "u1")])
    recs = []
    for by in range(0, h - 16, 16):                       # one 16x16 (or four 8x8) destination per macroblock cell, random source
        for bx in range(0, w - 16, 16):
            if r.rand() < 0.5:
                recs.append((by * st + bx, r.randint(0, h - 17) * st + r.randint(0, w - 17), r.randint(0, 3), 0, r.randint(0, 16), 0))
            else:
                for k in range(4):
                    recs.append(((by + 8 * (k >> 1)) * st + bx + 8 * (k & 1), r.randint(0, h - 9) * st + r.randint(0, w - 9), r.randint(0, 3), 1, r.randint(0, 16), 0))
    rec = np.array(recs, dt)
    want = base.copy()
    for q in rec:
        checker.mpeg4_qpel(int(q["kind"]), int(q["sidx"]), int(q["mc"]), at(want, q["dst_off"]), at(src, q["src_off"]), st)
    d_dst, d_src, d_rec = device.DevBuf.from_numpy(base), device.DevBuf.from_numpy(src), device.DevBuf.from_numpy(rec)
    gpu.check(gpu.lib.ff_mpeg4_qpel_batch_cuda(d_rec.ptr, len(rec), d_dst.ptr, d_src.ptr, st, None))
    device.sync()
    got = d_dst.download(np.uint8, base.shape)
    assert np.array_equal(got, want), np.argwhere(got != want)[:4].tolist()
