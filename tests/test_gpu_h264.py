"""GPU parity for the batched H.264 entry points (residual add, motion compensation, weighted prediction, deblocking
wavefront) against the CPU oracle applied macroblock by macroblock in the reference's order (tests/h264_util.py) --
small pictures, a ragged one, and the full config-3 geometry (1920x1088, 64 slices)."""
import numpy as np
import pytest

from libav_b200 import synth
import h264_util as hu

pytestmark = pytest.mark.gpu
SIZES = [(3, 2), (7, 5), (20, 12)]


def _dev(a):
    from libav_b200 import device
    return device.DevBuf.from_numpy(a)


@pytest.mark.parametrize("mb_w,mb_h", SIZES + [(120, 68)])
def test_residual_batch(gpu, checker, mb_w, mb_h):
    from libav_b200 import device
    y, cb, cr = synth.h264_picture(mb_w, mb_h, seed=5)
    rec, coeffs, nnzc = synth.h264_residual_work(mb_w, mb_h, seed=mb_w)
    wy, wcb, wcr, wco = y.copy(), cb.copy(), cr.copy(), coeffs.copy()
    hu.oracle_residual(checker, rec, wco, nnzc, wy, wcb, wcr)
    d = [_dev(x) for x in (rec, coeffs, nnzc, y, cb, cr)]
    gpu.check(gpu.lib.ff_h264_idct_add_mb_batch_cuda(d[0].ptr, rec.shape[0], d[1].ptr, 768, d[2].ptr, d[3].ptr, d[4].ptr, d[5].ptr,
                                                     y.strides[0], cb.strides[0], None))
    device.sync()
    assert np.array_equal(d[3].download(np.uint8, y.shape), wy)
    assert np.array_equal(d[4].download(np.uint8, cb.shape), wcb)
    assert np.array_equal(d[5].download(np.uint8, cr.shape), wcr)
    assert np.array_equal(d[1].download(np.int16, coeffs.shape), wco)      # consumed coefficients are zeroed identically


@pytest.mark.parametrize("bipred", [False, True])
@pytest.mark.parametrize("mb_w,mb_h", SIZES + [(120, 68)])
def test_mc_batch(gpu, checker, mb_w, mb_h, bipred):
    from libav_b200 import device
    refs = [synth.h264_picture(mb_w, mb_h, seed=11), synth.h264_picture(mb_w, mb_h, seed=12)]
    rec = synth.h264_mc_work(mb_w, mb_h, seed=mb_h, max_mv=64 if mb_w > 4 else 24, avg_second=bipred)
    y, cb, cr = synth.h264_picture(mb_w, mb_h, seed=13)
    wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
    hu.oracle_mc(checker, rec, refs, wy, wcb, wcr)
    dref = [[_dev(p) for p in r] for r in refs]
    planes = np.array([[p.ptr for p in r] for r in dref], dtype=np.uint64)
    d_planes, d_rec = _dev(planes), _dev(rec)
    dy, dcb, dcr = _dev(y), _dev(cb), _dev(cr)
    gpu.check(gpu.lib.ff_h264_mc_batch_cuda(d_rec.ptr, rec.shape[0], d_planes.ptr, dy.ptr, dcb.ptr, dcr.ptr, y.strides[0], cb.strides[0],
                                            16 * mb_w, 16 * mb_h, None))
    device.sync()
    assert np.array_equal(dy.download(np.uint8, y.shape), wy)
    assert np.array_equal(dcb.download(np.uint8, cb.shape), wcb)
    assert np.array_equal(dcr.download(np.uint8, cr.shape), wcr)


@pytest.mark.parametrize("mb_w,mb_h,slices", [(3, 2, 1), (7, 5, 1), (7, 5, 4), (20, 12, 1), (120, 68, 64), (120, 68, 1)])
def test_deblock_wavefront(gpu, checker, mb_w, mb_h, slices):
    from libav_b200 import device
    rng = np.random.default_rng(mb_w * 100 + slices)

    def smooth(shape):              # smooth content + noise so that a large share of the edges really gets filtered
        base = rng.integers(40, 200, size=(shape[0] // 8 + 1, shape[1] // 8 + 1))
        up = np.kron(base, np.ones((8, 8), dtype=np.int64))[:shape[0], :shape[1]]
        return np.clip(up + rng.integers(-6, 7, size=shape), 0, 255).astype(np.uint8)
    y, cb, cr = smooth((16 * mb_h, 16 * mb_w)), smooth((8 * mb_h, 8 * mb_w)), smooth((8 * mb_h, 8 * mb_w))
    rec = synth.h264_deblock_work(mb_w, mb_h, seed=slices, slices=slices)
    wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
    hu.oracle_deblock(checker, rec, mb_w, mb_h, wy, wcb, wcr)
    assert not np.array_equal(wy, y)                                    # the test does exercise the filters
    d_rec, dy, dcb, dcr = _dev(rec), _dev(y), _dev(cb), _dev(cr)
    prog = device.DevBuf(8 * mb_h)
    gpu.check(gpu.lib.ff_h264_deblock_picture_cuda(d_rec.ptr, mb_w, mb_h, dy.ptr, dcb.ptr, dcr.ptr, y.strides[0], cb.strides[0], prog.ptr, None))
    device.sync()
    assert np.array_equal(dy.download(np.uint8, y.shape), wy)
    assert np.array_equal(dcb.download(np.uint8, cb.shape), wcb)
    assert np.array_equal(dcr.download(np.uint8, cr.shape), wcr)


@pytest.mark.parametrize("bi", [0, 1])
def test_weight_batch(gpu, checker, bi):
    from libav_b200 import device
    rng = np.random.default_rng(bi)
    plane = rng.integers(0, 256, size=(128, 256), dtype=np.uint8)
    src = rng.integers(0, 256, size=(128, 256), dtype=np.uint8)
    recs = []
    for by in range(0, 128, 16):
        for bx in range(0, 256, 16):
            w = int(rng.choice([16, 8, 4, 2]))
            recs.append((by * 256 + bx, w, int(rng.choice([2, 4, 8, 16])), int(rng.integers(0, 8)), 0, int(rng.integers(-128, 128)),
                         int(rng.integers(-128, 128)), int(rng.integers(-128, 128)), 0))
    rec = np.array(recs, dtype=synth.WEIGHT_DT)
    want = plane.copy()
    hu.oracle_weight(checker, rec, want, src if bi else None)
    d_rec, d_pl, d_src = _dev(rec), _dev(plane), _dev(src)
    gpu.check(gpu.lib.ff_h264_weight_batch_cuda(d_rec.ptr, rec.shape[0], d_pl.ptr, d_src.ptr if bi else None, 256, None))
    device.sync()
    assert np.array_equal(d_pl.download(np.uint8, plane.shape), want)


def test_deblock_batch_of_stacked_pictures(gpu, checker):
    """Three independent pictures stacked vertically in one launch: each must equal its own serial result, and nothing
    may leak across the picture seams."""
    from libav_b200 import device
    mb_w, mb_h, P = 9, 4, 3
    rng = np.random.default_rng(9)
    ys, cbs, crs, recs, want = [], [], [], [], []
    for k in range(P):
        base = rng.integers(60, 190)
        y = np.clip(base + rng.integers(-8, 9, size=(16 * mb_h, 16 * mb_w)), 0, 255).astype(np.uint8)
        cb = np.clip(base + rng.integers(-8, 9, size=(8 * mb_h, 8 * mb_w)), 0, 255).astype(np.uint8)
        cr = np.clip(base + rng.integers(-8, 9, size=(8 * mb_h, 8 * mb_w)), 0, 255).astype(np.uint8)
        rec = synth.h264_deblock_work(mb_w, mb_h, seed=20 + k, slices=1 + k)
        wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
        hu.oracle_deblock(checker, rec, mb_w, mb_h, wy, wcb, wcr)
        ys.append(y); cbs.append(cb); crs.append(cr); recs.append(rec); want.append((wy, wcb, wcr))
    Y, CB, CR, R = np.concatenate(ys), np.concatenate(cbs), np.concatenate(crs), np.concatenate(recs)
    d_rec, dy, dcb, dcr = _dev(R), _dev(Y), _dev(CB), _dev(CR)
    prog = device.DevBuf(8 * mb_h * P)
    gpu.check(gpu.lib.ff_h264_deblock_batch_cuda(d_rec.ptr, mb_w, mb_h, P, dy.ptr, dcb.ptr, dcr.ptr, Y.strides[0], CB.strides[0], prog.ptr, None))
    device.sync()
    assert np.array_equal(dy.download(np.uint8, Y.shape), np.concatenate([w[0] for w in want]))
    assert np.array_equal(dcb.download(np.uint8, CB.shape), np.concatenate([w[1] for w in want]))
    assert np.array_equal(dcr.download(np.uint8, CR.shape), np.concatenate([w[2] for w in want]))


def test_mc_batch_of_stacked_pictures(gpu, checker):
    """Two pictures stacked vertically: vectors that leave a picture must be clamped to THAT picture's reference."""
    from libav_b200 import device
    mb_w, mb_h, P = 5, 3, 2
    H = 16 * mb_h
    refs = [[synth.h264_picture(mb_w, mb_h, seed=30 + 2 * k + r) for r in range(2)] for k in range(P)]
    recs, wants, outs = [], [], []
    for k in range(P):
        rec = synth.h264_mc_work(mb_w, mb_h, seed=40 + k, max_mv=80)
        y, cb, cr = synth.h264_picture(mb_w, mb_h, seed=50 + k)
        wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
        hu.oracle_mc(checker, rec, refs[k], wy, wcb, wcr, pad=64)
        r2 = rec.copy(); r2["y"] = r2["y"] + k * H
        recs.append(r2); wants.append((wy, wcb, wcr)); outs.append((y, cb, cr))
    rec = np.concatenate(recs)
    stack = lambda idx, r: np.concatenate([refs[k][r][idx] for k in range(P)])
    dref = [[_dev(stack(i, r)) for i in range(3)] for r in range(2)]
    planes = np.array([[p.ptr for p in r] for r in dref], dtype=np.uint64)
    Y, CB, CR = (np.concatenate([o[i] for o in outs]) for i in range(3))
    d_planes, d_rec, dy, dcb, dcr = _dev(planes), _dev(rec), _dev(Y), _dev(CB), _dev(CR)
    gpu.check(gpu.lib.ff_h264_mc_batch_cuda(d_rec.ptr, rec.shape[0], d_planes.ptr, dy.ptr, dcb.ptr, dcr.ptr, Y.strides[0], CB.strides[0],
                                            16 * mb_w, H, None))
    device.sync()
    assert np.array_equal(dy.download(np.uint8, Y.shape), np.concatenate([w[0] for w in wants]))
    assert np.array_equal(dcb.download(np.uint8, CB.shape), np.concatenate([w[1] for w in wants]))
    assert np.array_equal(dcr.download(np.uint8, CR.shape), np.concatenate([w[2] for w in wants]))


def test_dc_dequant_batch(gpu, checker):
    """luma / chroma DC transforms + dequantisation over a coefficient arena, against the oracle's two slot functions"""
    from libav_b200 import device
    from oracle.loader import ptr
    r = np.random.RandomState(77)
    n = 5000
    coeffs = r.randint(-2000, 2001, (n, 768)).astype(np.int16)
    luma_dc = r.randint(-2000, 2001, (n, 16)).astype(np.int16)
    rec = np.zeros(n, np.dtype([("luma_qmul", "<u4"), ("chroma_qmul", "<u4", (2,))]))
    rec["luma_qmul"] = r.randint(0, 4, n) * r.randint(16, 4000, n)          # a quarter of the macroblocks have no luma DC block
    rec["chroma_qmul"] = (r.randint(0, 3, (n, 2)) > 0) * r.randint(16, 4000, (n, 2))
    want = coeffs.copy()
    for m in range(n):
        if rec["luma_qmul"][m]:
            checker.h264_luma_dc_dequant_idct(ptr(want[m]), ptr(luma_dc[m].copy()), int(rec["luma_qmul"][m]))
        for p in range(2):
            if rec["chroma_qmul"][m, p]:
                blk = np.ascontiguousarray(want[m, 256 * (p + 1):256 * (p + 2)])
                checker.h264_chroma_dc_dequant_idct(ptr(blk), int(rec["chroma_qmul"][m, p]))
                want[m, 256 * (p + 1):256 * (p + 2)] = blk
    d_c, d_l, d_r = _dev(coeffs), _dev(luma_dc), _dev(rec)
    gpu.check(gpu.lib.ff_h264_dc_dequant_batch_cuda(d_r.ptr, n, d_c.ptr, 768, d_l.ptr, None))
    device.sync()
    got = d_c.download(np.int16, coeffs.shape)
    assert np.array_equal(got, want), np.argwhere((got != want).any(axis=1))[:4].ravel().tolist()
