"""CPU: pin oracle/port/orc_misc.c (forward DCTs, me_cmp metrics, full search, half-pel MC) byte-for-byte against the
unmodified reference in oracle/_ref on random inputs."""
import numpy as np
import pytest

from oracle.loader import ptr
from h264_util import at


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_fdct(orc, refo, which):
    rng = np.random.default_rng(which)
    for it in range(400):
        blk = rng.integers(-256, 256, size=64).astype(np.int16)
        if it % 10 == 0:
            blk = rng.integers(-1024, 1024, size=64).astype(np.int16)
        a, b = blk.copy(), blk.copy()
        orc.fdct(which, ptr(a))
        refo.fdct(which, ptr(b))
        assert np.array_equal(a, b), (which, it)


KINDS = [(0, s, d) for s in (0, 1) for d in range(4)] + [(1, 0, 0), (1, 1, 0), (2, 0, 0), (2, 1, 0), (2, 2, 0), (3, 0, 0), (3, 1, 0),
                                                         (4, 0, 0), (5, 0, 0), (6, 0, 0), (6, 1, 0), (7, 0, 0), (7, 1, 0), (8, 0, 0),
                                                         (8, 1, 0), (9, 0, 0), (9, 1, 0)]


@pytest.mark.parametrize("kind,sidx,dxy", KINDS)
def test_me_cmp(orc, refo, kind, sidx, dxy):
    rng = np.random.default_rng(kind * 100 + sidx * 10 + dxy)
    for it in range(60):
        a = rng.integers(0, 256, size=(20, 40), dtype=np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8) if it % 2 else rng.integers(0, 256, size=a.shape, dtype=np.uint8)
        w = 16 >> sidx
        h = 8 if kind in (3, 7) and sidx == 1 else int(rng.choice([8, 16])) if w == 16 else 8
        if kind in (3, 7) and sidx == 0:
            h = int(rng.choice([8, 16]))
        x = orc.me_cmp(kind, sidx, dxy, at(a, 40 + 3), at(b, 40 + 5), 40, h)
        y = refo.me_cmp(kind, sidx, dxy, at(a, 40 + 3), at(b, 40 + 5), 40, h)
        assert x == y and x >= 0, (kind, sidx, dxy, h, x, y)


def test_sum_abs_dctelem(orc, refo):
    rng = np.random.default_rng(0)
    for _ in range(50):
        blk = rng.integers(-2048, 2048, size=64).astype(np.int16)
        assert orc.me_cmp(10, 0, 0, ptr(blk), None, 0, 0) == refo.me_cmp(10, 0, 0, ptr(blk), None, 0, 0)


def test_full_search(orc, refo):
    rng = np.random.default_rng(4)
    w, h = 96, 64
    ref = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    cur = np.roll(ref, (3, -5), axis=(0, 1))
    cur = np.clip(cur.astype(int) + rng.integers(-3, 4, size=cur.shape), 0, 255).astype(np.uint8)
    flat = np.full((h, w), 128, dtype=np.uint8)                # all candidates tie: the first in raster order must win
    for c, r in ((cur, ref), (flat, flat)):
        outs = []
        for o in (orc, refo):
            out = np.zeros((h // 16) * (w // 16) * 3, np.int32)
            o.full_search(ptr(c), ptr(r), w, w, h, 16, 0, h // 16, ptr(out), 3)
            outs.append(out)
        assert np.array_equal(outs[0], outs[1])
    mv = outs[1].reshape(-1, 3)
    assert (mv[:, 2] == 0).all() and mv[0, 0] == 0 and mv[0, 1] == 0       # clipped window of the top-left MB starts at (0, 0)


@pytest.mark.parametrize("tab", [0, 1, 2, 3])
def test_hpel(orc, refo, tab):
    rng = np.random.default_rng(tab)
    for sidx in range(4):
        for dxy in range(4):
            w = 16 >> sidx
            for h in ([8, 16] if w == 16 else [4, 8] if w == 8 else [4] if w == 4 else [2]):
                pix = rng.integers(0, 256, size=(h + 2, 40), dtype=np.uint8)
                blk = rng.integers(0, 256, size=(h, 40), dtype=np.uint8)
                a, b = blk.copy(), blk.copy()
                ra = orc.hpel(tab, sidx, dxy, ptr(a), ptr(pix), 40, h)
                rb = refo.hpel(tab, sidx, dxy, ptr(b), ptr(pix), 40, h)
                assert ra == rb, (tab, sidx, dxy)
                assert np.array_equal(a, b), (tab, sidx, dxy, h)


def test_dct_metrics_port_matches_reference(orc, refo):
    """dct_sad / dct_max / dct264_sad (me_cmp.c:538-621): the port against the reference's functions called with an
    MpegEncContext that carries pdsp / fdsp / mecc like an encoder's, both fdct selections, 16x16 / 16x8 / 8x8"""
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.RandomState(1)
    for it in range(120):
        a = rng.randint(0, 256, (40, 48)).astype(np.uint8)
        b = rng.randint(0, 256, (40, 48)).astype(np.uint8) if it % 3 else np.clip(a.astype(int) + rng.randint(-6, 7, a.shape), 0, 255).astype(np.uint8)
        for kind in (11, 12, 13):
            for sidx, h in ((0, 16), (0, 8), (1, 8)):
                for sel in ((0, 2) if kind != 13 else (0,)):
                    x = refo.me_cmp(kind, sidx, sel, a.ctypes.data + 8, b.ctypes.data + 8, 48, h)
                    y = orc.me_cmp(kind, sidx, sel, a.ctypes.data + 8, b.ctypes.data + 8, 48, h)
                    assert x == y and x >= 0, (kind, sidx, h, sel, x, y)
