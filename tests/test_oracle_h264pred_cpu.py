"""H.264 intra prediction: the edge-array restatement (oracle/port/orc_h264pred.c) against the reference's own
H264PredContext (oracle/_ref, ff_h264_pred_init for codec H.264 / 8 bit / 4:2:0) -- every mode of pred4x4, pred8x8l
(all four has_topleft / has_topright combinations), pred8x8, pred16x16 and the ten lossless *_add predictors, on random
pictures; the whole buffer is compared so a stray write shows up too."""
import numpy as np
import pytest

from oracle.loader import ptr

TABS = {0: (4, 12), 1: (8, 12), 2: (8, 11), 3: (16, 7)}
STRIDE, H = 64, 48


def at(a, off):
    return a.ctypes.data + int(off)


def apply_pred(o, tab, mode, img, x, y, tl, tr, own_topright):
    img = img.copy()
    n = TABS[tab][0]
    trbuf = np.ascontiguousarray(img[y - 1, x + n:x + n + 4].copy()) if not own_topright else np.full(4, img[y - 1, x + n - 1], np.uint8)
    o.h264_pred(tab, mode, at(img, y * STRIDE + x), ptr(trbuf), tl, tr, STRIDE)
    return img


@pytest.mark.parametrize("tab", list(TABS))
def test_predictors_match_reference(orc, refo, tab):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = np.random.RandomState(tab)
    n, modes = TABS[tab]
    for it in range(40):
        img = r.randint(0, 256, (H, STRIDE)).astype(np.uint8)
        if it % 5 == 0:
            img[:] = r.choice([0, 255, 128])                       # flat pictures: plane prediction clipping
            img[::3, ::5] = r.randint(0, 256)
        x, y = 16 + 4 * r.randint(0, 3), 16 + 4 * r.randint(0, 3)
        for mode in range(modes):
            for tl in (0, 1):
                for tr in (0, 1):
                    if tab != 1 and (tl or tr):
                        continue
                    a = apply_pred(refo, tab, mode, img, x, y, tl, tr, it & 1)
                    b = apply_pred(orc, tab, mode, img, x, y, tl, tr, it & 1)
                    assert np.array_equal(a, b), (tab, mode, tl, tr, np.argwhere(a != b)[:4].tolist())
                    if mode not in (1,) or tab != 0:
                        pass


def block_offsets(stride, chroma):
    # frame-macroblock defaults (h264_slice.c:486-493): 4x4 block k at scan8-derived position
    offs = []
    for k in range(4 if chroma else 16):
        bx = (k & 1) + 2 * ((k >> 2) & 1) if not chroma else (k & 1)
        by = ((k >> 1) & 1) + 2 * (k >> 3) if not chroma else (k >> 1)
        offs.append(4 * bx + 4 * by * stride)
    return np.array(offs, np.int32)


@pytest.mark.parametrize("tab", range(5))
def test_lossless_add_predictors_match_reference(orc, refo, tab):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = np.random.RandomState(20 + tab)
    nco = {0: 16, 1: 64, 2: 64, 3: 64, 4: 256}[tab]
    bo = block_offsets(STRIDE, tab == 3)
    for it in range(60):
        img = r.randint(0, 256, (H, STRIDE)).astype(np.uint8)
        blk = r.randint(-300, 301, nco).astype(np.int16)
        for mode in (0, 1):
            for tl, tr in ((0, 0), (1, 1), (1, 0), (0, 1)) if tab == 2 else ((0, 0),):
                outs = []
                for o in (refo, orc):
                    i2, b2 = img.copy(), blk.copy()
                    o.h264_pred_add(tab, mode, at(i2, 16 * STRIDE + 16), ptr(bo), ptr(b2), tl, tr, STRIDE)
                    outs.append((i2, b2))
                assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), (tab, mode, tl, tr)
                assert not outs[1][1].any()
