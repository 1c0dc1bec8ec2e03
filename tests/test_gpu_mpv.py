"""GPU parity for the inverse quantisers (mpegvideo half): ff_mpeg_dequant_batch_cuda against the CPU
checker block by block, and the fused dequantise + simple IDCT kernel against checker-dequant followed by the oracle's
idct_put / idct_add -- every kind, both scans, AIC / AC prediction, full-range levels, skipped inter blocks, fused
clear, offset-addressed and tile-addressed destinations."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import tables
from oracle.loader import ptr

pytestmark = pytest.mark.gpu
DQ_DT = np.dtype([("qscale", "u1"), ("last", "i1"), ("dc", "u1"), ("flags", "u1")])
INTER = (1, 4, 6)


def _dev(a):
    from libav_b200 import device
    return device.DevBuf.from_numpy(np.ascontiguousarray(a))


def make_case(checker, kind, n, seed, alt, aic, wild):
    r = np.random.RandomState(seed)
    blocks = np.zeros((n, 64), np.int16)
    for i in range(n):
        nz = r.randint(1, 40)
        pos = r.permutation(64)[:nz]
        blocks[i, pos] = r.randint(-32768, 32768, nz) if (wild and i % 4 == 0) else r.randint(-300, 301, nz)
    rec = np.zeros(n, DQ_DT)
    rec["qscale"] = r.randint(1, 32, n)
    rec["last"] = r.randint(-1 if kind in INTER else 0, 64, n)
    rec["dc"] = r.choice([1, 2, 4, 8, 13, 25], n)
    rec["flags"] = r.randint(0, 2, n)
    t = tables.FFMpegDequantTables()
    perm, rend = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
    checker.mpeg_scantables(alt, ptr(perm), ptr(rend))
    intra, inter = r.randint(1, 256, 64).astype(np.uint16), r.randint(1, 256, 64).astype(np.uint16)
    for k in range(64):
        t.intra_matrix[k], t.inter_matrix[k], t.permutated[k], t.raster_end[k] = int(intra[k]), int(inter[k]), int(perm[k]), int(rend[k])
    t.alternate_scan, t.h263_aic = alt, aic
    return blocks, rec, t, intra, inter


def oracle_dequant(checker, kind, blocks, rec, intra, inter, alt, aic):
    out = blocks.copy()
    for i in range(blocks.shape[0]):
        last = int(rec["last"][i])
        if kind in INTER and last < 0:
            continue                                       # add_dequant_dct: untouched
        b = np.ascontiguousarray(out[i])
        # n = 0 selects y_dc_scale: the record's dc_scale is passed as both
        checker.mpeg_dequant(kind, ptr(b), 0, int(rec["qscale"][i]), last, int(rec["dc"][i]), int(rec["dc"][i]), ptr(intra), ptr(inter),
                             alt, aic, int(rec["flags"][i]) & 1)
        out[i] = b
    return out


@pytest.mark.parametrize("alt,aic", [(0, 0), (1, 1)])
@pytest.mark.parametrize("kind", range(7))
def test_dequant_batch(gpu, checker, kind, alt, aic):
    from libav_b200 import device
    n = 1500
    blocks, rec, t, intra, inter = make_case(checker, kind, n, 7 + kind, alt, aic, wild=True)
    want = oracle_dequant(checker, kind, blocks, rec, intra, inter, alt, aic)
    d_b, d_r = _dev(blocks), _dev(rec)
    gpu.check(gpu.lib.ff_mpeg_dequant_batch_cuda(kind, C.byref(t), d_r.ptr, d_b.ptr, n, None))
    device.sync()
    got = d_b.download(np.int16, blocks.shape)
    bad = np.argwhere((got != want).any(axis=1))
    assert not len(bad), (kind, bad[:3].ravel().tolist(), rec[bad[0, 0]], blocks[bad[0, 0]].tolist(), got[bad[0, 0]].tolist(), want[bad[0, 0]].tolist())


@pytest.mark.parametrize("clear,use_off", [(0, False), (1, True)])
@pytest.mark.parametrize("kind", range(7))
def test_dequant_idct_fused(gpu, checker, kind, clear, use_off):
    from libav_b200 import device
    n, tpr = 1000 + 13, 37                                  # ragged last warp, ragged last tile row
    alt, aic = kind & 1, (kind >> 1) & 1
    blocks, rec, t, intra, inter = make_case(checker, kind, n, 40 + kind, alt, aic, wild=True)
    rows = (n + tpr - 1) // tpr
    stride = tpr * 8 + 24
    frame = np.random.RandomState(3).randint(0, 256, (rows * 8, stride)).astype(np.uint8)
    off = np.array([(i // tpr) * 8 * stride + (i % tpr) * 8 for i in range(n)], np.uint32)
    if use_off:
        off = off[np.random.RandomState(4).permutation(n)]
    # oracle: dequantise, then idct_put (intra kinds) / idct_add (inter kinds, skipping blocks with last < 0)
    deq = oracle_dequant(checker, kind, blocks, rec, intra, inter, alt, aic)
    want = frame.copy()
    mode = 1 if kind in INTER else 0
    keep = np.ones(n, bool) if mode == 0 else (rec["last"] >= 0)
    idx = np.flatnonzero(keep)
    sub = np.ascontiguousarray(deq[idx]); suboff = np.ascontiguousarray(off[idx])
    checker.idct_batch(mode, ptr(sub), ptr(want), ptr(suboff), stride, len(idx), 1)
    d_b, d_r, d_f, d_o = _dev(blocks), _dev(rec), _dev(frame), _dev(off)
    gpu.check(gpu.lib.ff_mpeg_dequant_idct_batch_cuda(kind, C.byref(t), d_r.ptr, d_b.ptr, d_f.ptr, d_o.ptr if use_off else None, stride, n,
                                                      0 if use_off else tpr, clear, None))
    device.sync()
    got = d_f.download(np.uint8, frame.shape)
    assert np.array_equal(got, want), (kind, np.argwhere(got != want)[:4].tolist())
    after = d_b.download(np.int16, blocks.shape)
    if clear:
        assert not after[idx].any() and np.array_equal(after[~keep], blocks[~keep])
    else:
        assert np.array_equal(after, blocks)


def test_bad_arguments_fail_loudly(gpu):
    t = tables.FFMpegDequantTables()                        # all-zero permutation is not a permutation
    assert gpu.lib.ff_mpeg_dequant_batch_cuda(2, C.byref(t), None, None, 4, None) == -1
    assert "permutation" in gpu.last_error()
    gpu.lib.avb200_clear_error()
