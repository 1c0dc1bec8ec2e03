"""GPU parity for the batched intra reconstruction wavefront: ff_h264_intra_mb_batch_cuda against the
oracle's predictors + transforms applied macroblock by macroblock in hl_decode_mb()'s order -- all-intra pictures, mixed
pictures whose inter macroblocks only serve as neighbours, a stacked batch, the config-3 geometry."""
import numpy as np
import pytest

from libav_b200 import synth
import h264_util as hu

pytestmark = pytest.mark.gpu


def _dev(a):
    from libav_b200 import device
    return device.DevBuf.from_numpy(np.ascontiguousarray(a))


def run_gpu(gpu, rec, coeffs, nnzc, mb_w, mb_h, n_pic, y, cb, cr):
    from libav_b200 import device
    d = [_dev(v) for v in (rec, coeffs, nnzc, y, cb, cr)]
    prog = device.DevBuf(4 * mb_h * n_pic)
    gpu.check(gpu.lib.ff_h264_intra_mb_batch_cuda(d[0].ptr, mb_w, mb_h, n_pic, d[1].ptr, 768, d[2].ptr, d[3].ptr, d[4].ptr, d[5].ptr,
                                                  y.strides[0], cb.strides[0], prog.ptr, None))
    device.sync()
    return (d[3].download(np.uint8, y.shape), d[4].download(np.uint8, cb.shape), d[5].download(np.uint8, cr.shape),
            d[1].download(np.int16, coeffs.shape))


@pytest.mark.parametrize("mb_w,mb_h,p_intra", [(1, 1, 1.0), (3, 2, 1.0), (7, 5, 1.0), (9, 6, 0.4), (20, 12, 0.7)])
def test_intra_picture(gpu, checker, mb_w, mb_h, p_intra):
    y, cb, cr = synth.h264_picture(mb_w, mb_h, seed=21)
    rec, coeffs, nnzc = synth.h264_intra_work(mb_w, mb_h, seed=mb_w + 3, p_intra=p_intra)
    wy, wcb, wcr, wco = y.copy(), cb.copy(), cr.copy(), coeffs.copy()
    hu.oracle_intra(checker, rec, wco, nnzc, mb_w, mb_h, wy, wcb, wcr)
    gy, gcb, gcr, gco = run_gpu(gpu, rec, coeffs, nnzc, mb_w, mb_h, 1, y, cb, cr)
    assert np.array_equal(gy, wy), np.argwhere(gy != wy)[:5].tolist()
    assert np.array_equal(gcb, wcb) and np.array_equal(gcr, wcr)
    assert np.array_equal(gco, wco)                       # consumed coefficients are zeroed identically
    assert (wy != y).mean() > 0.3 * p_intra


def test_stacked_pictures_do_not_see_each_other(gpu, checker):
    mb_w, mb_h, P = 6, 4, 3
    ys, cbs, crs, recs, cos, nzs, want = [], [], [], [], [], [], []
    for k in range(P):
        y, cb, cr = synth.h264_picture(mb_w, mb_h, seed=30 + k)
        rec, co, nz = synth.h264_intra_work(mb_w, mb_h, seed=40 + k, p_intra=0.8)
        wy, wcb, wcr, wco = y.copy(), cb.copy(), cr.copy(), co.copy()
        hu.oracle_intra(checker, rec, wco, nz, mb_w, mb_h, wy, wcb, wcr)
        ys.append(y); cbs.append(cb); crs.append(cr); recs.append(rec); cos.append(co); nzs.append(nz); want.append((wy, wcb, wcr))
    Y, CB, CR = np.concatenate(ys), np.concatenate(cbs), np.concatenate(crs)
    gy, gcb, gcr, _ = run_gpu(gpu, np.concatenate(recs), np.concatenate(cos), np.concatenate(nzs), mb_w, mb_h, P, Y, CB, CR)
    assert np.array_equal(gy, np.concatenate([w[0] for w in want]))
    assert np.array_equal(gcb, np.concatenate([w[1] for w in want])) and np.array_equal(gcr, np.concatenate([w[2] for w in want]))


def test_full_hd_intra_picture(gpu, checker):
    mb_w, mb_h = 120, 68
    y, cb, cr = synth.h264_picture(mb_w, mb_h, seed=2)
    rec, coeffs, nnzc = synth.h264_intra_work(mb_w, mb_h, seed=9)
    wy, wcb, wcr, wco = y.copy(), cb.copy(), cr.copy(), coeffs.copy()
    hu.oracle_intra(checker, rec, wco, nnzc, mb_w, mb_h, wy, wcb, wcr)
    gy, gcb, gcr, gco = run_gpu(gpu, rec, coeffs, nnzc, mb_w, mb_h, 1, y, cb, cr)
    assert np.array_equal(gy, wy) and np.array_equal(gcb, wcb) and np.array_equal(gcr, wcr) and np.array_equal(gco, wco)
