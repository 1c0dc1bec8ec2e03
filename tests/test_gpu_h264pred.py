"""GPU parity for H264PredContext (intra half): the table filled by ff_h264_pred_init_cuda driven like
the decoder drives it -- every mode of pred4x4 / pred8x8l (all availability flag pairs) / pred8x8 / pred16x16 and the
ten lossless *_add slots -- against the CPU checker on random pictures, whole buffers compared."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import tables
from oracle.loader import ptr
from test_oracle_h264pred_cpu import TABS, STRIDE, H, at, block_offsets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hpc(gpu):
    h = tables.H264PredContext()
    gpu.lib.ff_h264_pred_init_cuda(C.byref(h), 27, 8, 1)
    return h


def test_other_codecs_are_left_alone(gpu):
    h = tables.H264PredContext()
    gpu.lib.ff_h264_pred_init_cuda(C.byref(h), 27, 12, 1)       # (9 / 10 bit are taken over: tests/test_zz_gpu_late_slots.py)
    gpu.lib.ff_h264_pred_init_cuda(C.byref(h), 139, 8, 1)
    assert not any(C.cast(f, C.c_void_p).value for f in h.pred4x4)


@pytest.mark.parametrize("tab", list(TABS))
def test_prediction_slots(gpu, checker, hpc, tab):
    r = np.random.RandomState(50 + tab)
    n, modes = TABS[tab]
    table = [hpc.pred4x4, hpc.pred8x8l, hpc.pred8x8, hpc.pred16x16][tab]
    for it in range(6):
        img = r.randint(0, 256, (H, STRIDE)).astype(np.uint8)
        if it == 5:
            img[:] = 255; img[::3, ::5] = 0
        x, y = 16 + 4 * r.randint(0, 3), 16 + 4 * r.randint(0, 3)
        for mode in range(modes):
            for tl, tr in (((0, 0), (1, 1), (1, 0), (0, 1)) if tab == 1 else ((0, 0),)):
                trbuf = np.ascontiguousarray(img[y - 1, x + n:x + n + 4].copy()) if it & 1 else np.full(4, img[y - 1, x + n - 1], np.uint8)
                want, got = img.copy(), img.copy()
                checker.h264_pred(tab, mode, at(want, y * STRIDE + x), ptr(trbuf), tl, tr, STRIDE)
                p = at(got, y * STRIDE + x)
                if tab == 0: table[mode](p, ptr(trbuf), STRIDE)
                elif tab == 1: table[mode](p, tl, tr, STRIDE)
                else: table[mode](p, STRIDE)
                assert gpu.last_error() == "", gpu.last_error()
                assert np.array_equal(got, want), (tab, mode, tl, tr, np.argwhere(got != want)[:4].tolist())


@pytest.mark.parametrize("tab", range(5))
def test_lossless_add_slots(gpu, checker, hpc, tab):
    r = np.random.RandomState(70 + tab)
    nco = {0: 16, 1: 64, 2: 64, 3: 64, 4: 256}[tab]
    bo = block_offsets(STRIDE, tab == 3)
    for it in range(8):
        img = r.randint(0, 256, (H, STRIDE)).astype(np.uint8)
        blk = r.randint(-300, 301, nco).astype(np.int16)
        for mode in (0, 1):
            for tl, tr in (((0, 0), (1, 1), (1, 0), (0, 1)) if tab == 2 else ((0, 0),)):
                wi, wb, gi, gb = img.copy(), blk.copy(), img.copy(), blk.copy()
                checker.h264_pred_add(tab, mode, at(wi, 16 * STRIDE + 16), ptr(bo), ptr(wb), tl, tr, STRIDE)
                p = at(gi, 16 * STRIDE + 16)
                idx8 = 1 if mode else 2                      # HOR_PRED8x8 = 1, VERT_PRED8x8 = 2
                if tab == 0: hpc.pred4x4_add[mode](p, ptr(gb), STRIDE)
                elif tab == 1: hpc.pred8x8l_add[mode](p, ptr(gb), STRIDE)
                elif tab == 2: hpc.pred8x8l_filter_add[mode](p, ptr(gb), tl, tr, STRIDE)
                elif tab == 3: hpc.pred8x8_add[idx8](p, ptr(bo), ptr(gb), STRIDE)
                else: hpc.pred16x16_add[idx8](p, ptr(bo), ptr(gb), STRIDE)
                assert gpu.last_error() == "", gpu.last_error()
                assert np.array_equal(gi, wi) and np.array_equal(gb, wb), (tab, mode, tl, tr)
