"""PixblockDSPContext (get_pixels / diff_pixels): port vs compiled reference on the CPU; on the GPU the table slots and the
batched fetch(-difference) -> forward-DCT kernel against the checker (pixblock followed by the oracle's fdct)."""
import ctypes as C

import numpy as np
import pytest

from oracle.loader import ptr


def test_port_matches_reference(orc, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = np.random.RandomState(1)
    for it in range(200):
        a, b = r.randint(0, 256, (16, 40)).astype(np.uint8), r.randint(0, 256, (16, 40)).astype(np.uint8)
        for kind in (0, 1):
            o1, o2 = np.zeros(64, np.int16), np.zeros(64, np.int16)
            refo.pixblock(kind, ptr(o1), a.ctypes.data + 8, b.ctypes.data + 8, 40)
            orc.pixblock(kind, ptr(o2), a.ctypes.data + 8, b.ctypes.data + 8, 40)
            assert np.array_equal(o1, o2)


@pytest.mark.gpu
def test_slots(gpu, checker):
    from libav_b200 import tables
    c = tables.PixblockDSPContext()
    gpu.lib.ff_pixblockdsp_init_cuda(C.byref(c), 0)
    r = np.random.RandomState(2)
    for it in range(20):
        a, b = r.randint(0, 256, (16, 40)).astype(np.uint8), r.randint(0, 256, (16, 40)).astype(np.uint8)
        off = int(r.randint(0, 20))
        want, got = np.zeros(64, np.int16), np.zeros(64, np.int16)
        checker.pixblock(0, ptr(want), a.ctypes.data + off, None, 40)
        c.get_pixels(ptr(got), a.ctypes.data + off, 40)
        assert np.array_equal(got, want) and gpu.last_error() == ""
        checker.pixblock(1, ptr(want), a.ctypes.data + off, b.ctypes.data + off, 40)
        c.diff_pixels(ptr(got), a.ctypes.data + off, b.ctypes.data + off, 40)
        assert np.array_equal(got, want) and gpu.last_error() == ""
    c2 = tables.PixblockDSPContext()
    gpu.lib.ff_pixblockdsp_init_cuda(C.byref(c2), 1)
    assert not C.cast(c2.get_pixels, C.c_void_p).value


@pytest.mark.gpu
@pytest.mark.parametrize("which", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("diff", [False, True])
def test_fetch_fdct_batch(gpu, checker, which, diff):
    from libav_b200 import device
    r = np.random.RandomState(10 + which)
    w, h, n = 352, 288, 3000
    s1, s2 = r.randint(0, 256, (h, w)).astype(np.uint8), r.randint(0, 256, (h, w)).astype(np.uint8)
    ys, xs = r.randint(0, h - 8, n), r.randint(0, w - 8, n)
    xs[::2] &= ~7                                            # half of the blocks take the 8-byte aligned path
    off1 = (ys * w + xs).astype(np.uint32)
    off2 = (r.randint(0, h - 8, n) * w + r.randint(0, w - 8, n)).astype(np.uint32) if diff else None
    want = np.zeros((n, 64), np.int16)
    for i in range(n):
        checker.pixblock(1 if diff else 0, ptr(want[i]), s1.ctypes.data + int(off1[i]), s2.ctypes.data + int(off2[i]) if diff else None, w)
        if which >= 0:
            checker.fdct(which, ptr(want[i]))
    d1, d2, do1 = device.DevBuf.from_numpy(s1), device.DevBuf.from_numpy(s2), device.DevBuf.from_numpy(off1)
    do2 = device.DevBuf.from_numpy(off2) if diff else None
    out = device.DevBuf(n * 128)
    gpu.check(gpu.lib.ff_pixblock_fdct_batch_cuda(which, d1.ptr, d2.ptr if diff else None, do1.ptr, do2.ptr if diff else None, w, out.ptr, n, None))
    device.sync()
    got = out.download(np.int16, (n, 64))
    assert np.array_equal(got, want), np.argwhere((got != want).any(axis=1))[:4].ravel().tolist()
