"""GPU parity for the batched H.264 calls on 9 / 10-bit and 4:2:2 pictures (libav_b200/csrc/h264_hbd_batch.cu) against the CPU oracle's
BIT_DEPTH > 8 instances applied macroblock by macroblock in the reference's order (tests/h264_hbd_util.py).  Sorted last: written after the
rest of the suite had been green on a B200."""
import numpy as np
import pytest

from libav_b200 import synth
import h264_hbd_util as hh

pytestmark = pytest.mark.gpu
SIZES = [(3, 2), (7, 5), (20, 12)]


def _dev(a):
    from libav_b200 import device
    return device.DevBuf.from_numpy(a)


@pytest.mark.parametrize("c422", [0, 1])
@pytest.mark.parametrize("bits", [9, 10])
@pytest.mark.parametrize("mb_w,mb_h", SIZES)
def test_residual_batch_hbd(gpu, checker, mb_w, mb_h, bits, c422):
    from libav_b200 import device
    y, cb, cr = hh.picture(mb_w, mb_h, bits, c422, seed=5)
    rec, coeffs, nnzc = hh.residual_work(mb_w, mb_h, bits, c422, y, cb, seed=mb_w + bits)
    wy, wcb, wcr, wco = y.copy(), cb.copy(), cr.copy(), coeffs.copy()
    hh.oracle_residual(checker, bits, c422, rec, wco, nnzc, wy, wcb, wcr)
    assert not np.array_equal(wy, y) and not np.array_equal(wcb, cb)
    d = [_dev(x) for x in (rec, coeffs, nnzc, y, cb, cr)]
    gpu.check(gpu.lib.ff_h264_idct_add_mb_batch_hbd_cuda(bits, 1 + c422, d[0].ptr, rec.shape[0], d[1].ptr, 768, d[2].ptr, d[3].ptr, d[4].ptr, d[5].ptr,
                                                         y.strides[0], cb.strides[0], None))
    device.sync()
    assert np.array_equal(d[3].download(np.uint16, y.shape), wy)
    assert np.array_equal(d[4].download(np.uint16, cb.shape), wcb)
    assert np.array_equal(d[5].download(np.uint16, cr.shape), wcr)
    assert np.array_equal(d[1].download(np.int32, coeffs.shape), wco)      # consumed coefficients are zeroed identically


@pytest.mark.parametrize("staged", [0, 2])
@pytest.mark.parametrize("c422", [0, 1])
@pytest.mark.parametrize("bits", [9, 10])
@pytest.mark.parametrize("mb_w,mb_h", SIZES)
def test_mc_batch_hbd(gpu, checker, mb_w, mb_h, bits, c422, staged):
    """staged 0: the default kernel (patches staged in shared memory); 2: the clamped-global-load form of the same arithmetic (tests/hostsim/ runs it)"""
    from libav_b200 import device
    gpu.lib.avb200_set_tuning(b"mc_hbd_staged", staged)
    refs = [hh.picture(mb_w, mb_h, bits, c422, seed=11), hh.picture(mb_w, mb_h, bits, c422, seed=12)]
    rec = synth.h264_mc_work(mb_w, mb_h, seed=mb_h + bits, max_mv=64 if mb_w > 4 else 24, avg_second=True)
    y, cb, cr = hh.picture(mb_w, mb_h, bits, c422, seed=13)
    wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
    hh.oracle_mc(checker, bits, c422, rec, refs, wy, wcb, wcr)
    dref = [[_dev(p) for p in r] for r in refs]
    planes = np.array([[p.ptr for p in r] for r in dref], dtype=np.uint64)
    d_planes, d_rec = _dev(planes), _dev(rec)
    dy, dcb, dcr = _dev(y), _dev(cb), _dev(cr)
    gpu.check(gpu.lib.ff_h264_mc_batch_hbd_cuda(bits, 1 + c422, d_rec.ptr, rec.shape[0], d_planes.ptr, dy.ptr, dcb.ptr, dcr.ptr, y.strides[0], cb.strides[0],
                                                16 * mb_w, 16 * mb_h, None))
    device.sync()
    gpu.lib.avb200_set_tuning(b"mc_hbd_staged", 0)
    assert np.array_equal(dy.download(np.uint16, y.shape), wy)
    assert np.array_equal(dcb.download(np.uint16, cb.shape), wcb)
    assert np.array_equal(dcr.download(np.uint16, cr.shape), wcr)


@pytest.mark.parametrize("bits", [9, 10])
@pytest.mark.parametrize("mb_w,mb_h,slices,P", [(3, 2, 1, 1), (7, 5, 4, 2), (20, 12, 1, 3), (40, 36, 8, 2)])
def test_deblock_batch_hbd(gpu, checker, mb_w, mb_h, slices, P, bits):
    """P stacked pictures, rows of one picture as a wavefront inside one CTA: every picture must equal its own serial result"""
    from libav_b200 import device
    ys, cbs, crs, recs, want = [], [], [], [], []
    for k in range(P):
        y, cb, cr = hh.smooth_picture(mb_w, mb_h, bits, seed=100 * mb_w + k)
        rec = synth.h264_deblock_work(mb_w, mb_h, seed=slices + k, slices=slices)
        wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
        hh.oracle_deblock(checker, bits, rec, mb_w, mb_h, wy, wcb, wcr)
        assert not np.array_equal(wy, y) and not np.array_equal(wcb, cb)
        ys.append(y); cbs.append(cb); crs.append(cr); recs.append(rec); want.append((wy, wcb, wcr))
    Y, CB, CR, R = np.concatenate(ys), np.concatenate(cbs), np.concatenate(crs), np.concatenate(recs)
    d_rec, dy, dcb, dcr = _dev(R), _dev(Y), _dev(CB), _dev(CR)
    gpu.check(gpu.lib.ff_h264_deblock_batch_hbd_cuda(bits, d_rec.ptr, mb_w, mb_h, P, dy.ptr, dcb.ptr, dcr.ptr, Y.strides[0], CB.strides[0], None))
    device.sync()
    gy, gcb, gcr = dy.download(np.uint16, Y.shape), dcb.download(np.uint16, CB.shape), dcr.download(np.uint16, CR.shape)
    for k in range(P):
        assert np.array_equal(gy[16 * mb_h * k:16 * mb_h * (k + 1)], want[k][0]), k
        assert np.array_equal(gcb[8 * mb_h * k:8 * mb_h * (k + 1)], want[k][1]), k
        assert np.array_equal(gcr[8 * mb_h * k:8 * mb_h * (k + 1)], want[k][2]), k


def test_hbd_refusals(gpu):
    lib = gpu.lib
    assert lib.ff_h264_idct_add_mb_batch_hbd_cuda(8, 1, None, 0, None, 768, None, None, None, None, 64, 32, None) == -1
    assert lib.ff_h264_mc_batch_hbd_cuda(10, 3, None, 0, None, None, None, None, 64, 32, 16, 16, None) == -1
    assert lib.ff_h264_deblock_batch_hbd_cuda(10, None, 1, 1, 1, None, None, None, 64, 32, None) == -1
    lib.avb200_clear_error()


@pytest.mark.parametrize("bits", [8, 9, 10])
@pytest.mark.parametrize("mb_w,mb_h,slices,P", [(3, 2, 1, 1), (7, 5, 4, 2), (20, 12, 1, 3), (40, 36, 8, 2)])
def test_deblock_batch_422(gpu, checker, mb_w, mb_h, slices, P, bits):
    """chroma_format_idc 2: two 16-line vertical and four horizontal chroma edges per macroblock (8 / 9 / 10 bit), stacked pictures"""
    from libav_b200 import device
    dt = np.uint8 if bits == 8 else np.uint16
    ys, cbs, crs, recs, exts, want = [], [], [], [], [], []
    for k in range(P):
        y, cb, cr = hh.smooth_picture422(mb_w, mb_h, bits, seed=100 * mb_w + k)
        rec, ext = hh.deblock422_work(mb_w, mb_h, seed=slices + k, slices=slices)
        wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
        hh.oracle_deblock422(checker, bits, rec, ext, mb_w, mb_h, wy, wcb, wcr)
        assert not np.array_equal(wy, y) and not np.array_equal(wcb, cb)
        ys.append(y); cbs.append(cb); crs.append(cr); recs.append(rec); exts.append(ext); want.append((wy, wcb, wcr))
    Y, CB, CR, R, X = np.concatenate(ys), np.concatenate(cbs), np.concatenate(crs), np.concatenate(recs), np.concatenate(exts)
    d_rec, d_ext, dy, dcb, dcr = _dev(R), _dev(X), _dev(Y), _dev(CB), _dev(CR)
    gpu.check(gpu.lib.ff_h264_deblock_batch_422_cuda(bits, d_rec.ptr, d_ext.ptr, mb_w, mb_h, P, dy.ptr, dcb.ptr, dcr.ptr, Y.strides[0], CB.strides[0], None))
    device.sync()
    gy, gcb, gcr = dy.download(dt, Y.shape), dcb.download(dt, CB.shape), dcr.download(dt, CR.shape)
    for k in range(P):
        assert np.array_equal(gy[16 * mb_h * k:16 * mb_h * (k + 1)], want[k][0]), k
        assert np.array_equal(gcb[16 * mb_h * k:16 * mb_h * (k + 1)], want[k][1]), k
        assert np.array_equal(gcr[16 * mb_h * k:16 * mb_h * (k + 1)], want[k][2]), k


def test_chroma422_decisions(gpu, checker):
    """FFH264DeblockInfo.chroma422: the decisions of a 4:2:2 picture (h264_loopfilter.c:633,693-700) against the reference's own h264_loopfilter.c"""
    import ctypes as C
    from libav_b200 import device, tables
    from test_oracle_h264lf_cpu import C422_CASES, run422
    from test_gpu_h264lf import gpu_params
    for case in C422_CASES:
        for (mw, mh) in ((11, 9), (2, 5), (40, 17)):
            d = synth.h264_deblock_info(mw, mh, **case)
            want, want_ext = run422(checker, d)
            ext = device.DevBuf(52 * mw * mh); ext.fill(0xCD)
            got, _ = gpu_params(gpu, [d], chroma422=ext.ptr)
            bad = np.argwhere((got[:, :102] != want[:, :102]).any(axis=1))
            assert not len(bad), (case, mw, mh, bad[:4].ravel().tolist(), got[bad[0, 0]].tolist(), want[bad[0, 0]].tolist())
            gext = ext.download(np.uint8, (mw * mh, 52))
            bad = np.argwhere((gext[:, :50] != want_ext[:, :50]).any(axis=1))
            assert not len(bad), ("ext", case, mw, mh, bad[:4].ravel().tolist(), gext[bad[0, 0]].tolist(), want_ext[bad[0, 0]].tolist())


@pytest.mark.parametrize("bits", [9, 10])
def test_weight_and_dc_batch_hbd(gpu, checker, bits):
    from libav_b200 import device
    lib = gpu.lib

    def run_weight(b, rec, plane, src):
        d_rec, d_pl = _dev(rec), _dev(plane)
        d_src = _dev(src) if src is not None else None
        gpu.check(lib.ff_h264_weight_batch_hbd_cuda(b, d_rec.ptr, rec.shape[0], d_pl.ptr, d_src.ptr if d_src else None, plane.strides[0], None))
        device.sync()
        return d_pl.download(np.uint16, plane.shape)

    def run_dc(c422, recs, coeffs, luma_dc):
        d_rec, d_co, d_dc = _dev(recs), _dev(coeffs), _dev(luma_dc)
        gpu.check(lib.ff_h264_dc_dequant_batch_hbd_cuda(1 + c422, d_rec.ptr, recs.shape[0], d_co.ptr, 768, d_dc.ptr, None))
        device.sync()
        return d_co.download(np.int32, coeffs.shape)
    hh.weight_dc_cases(run_weight, run_dc, checker, bits)


@pytest.mark.parametrize("bits", [9, 10])
@pytest.mark.parametrize("mb_w,mb_h,P,p_intra", [(3, 2, 1, 1.0), (7, 5, 2, 1.0), (20, 12, 2, 0.6), (40, 30, 3, 1.0)])
def test_intra_batch_hbd(gpu, checker, mb_w, mb_h, P, p_intra, bits):
    """ff_h264_intra_mb_batch_hbd_cuda: all-intra and mixed pictures, stacked; samples and consumed coefficients against the oracle chain"""
    from libav_b200 import device
    ys, cbs, crs, recs, cos, nzs, want = [], [], [], [], [], [], []
    for k in range(P):
        y, cb, cr = hh.picture(mb_w, mb_h, bits, 0, seed=50 + k)
        rec, coeffs, nnzc = hh.intra_work(mb_w, mb_h, bits, seed=mb_w + k, p_intra=p_intra)
        wy, wcb, wcr, wco = y.copy(), cb.copy(), cr.copy(), coeffs.copy()
        hh.oracle_intra(checker, bits, rec, wco, nnzc, mb_w, mb_h, wy, wcb, wcr)
        assert not np.array_equal(wy, y)
        ys.append(y); cbs.append(cb); crs.append(cr); recs.append(rec); cos.append(coeffs); nzs.append(nnzc); want.append((wy, wcb, wcr, wco))
    Y, CB, CR = np.concatenate(ys), np.concatenate(cbs), np.concatenate(crs)
    R, CO, NZ = np.concatenate(recs), np.concatenate(cos), np.concatenate(nzs)
    d_rec, d_co, d_nz, dy, dcb, dcr = _dev(R), _dev(CO), _dev(NZ), _dev(Y), _dev(CB), _dev(CR)
    gpu.check(gpu.lib.ff_h264_intra_mb_batch_hbd_cuda(bits, d_rec.ptr, mb_w, mb_h, P, d_co.ptr, 768, d_nz.ptr, dy.ptr, dcb.ptr, dcr.ptr, Y.strides[0], CB.strides[0], None))
    device.sync()
    gy, gcb, gcr, gco = dy.download(np.uint16, Y.shape), dcb.download(np.uint16, CB.shape), dcr.download(np.uint16, CR.shape), d_co.download(np.int32, CO.shape)
    n = mb_w * mb_h
    for k in range(P):
        assert np.array_equal(gy[16 * mb_h * k:16 * mb_h * (k + 1)], want[k][0]), (k, np.argwhere(gy[16 * mb_h * k:16 * mb_h * (k + 1)] != want[k][0])[:4].tolist())
        assert np.array_equal(gcb[8 * mb_h * k:8 * mb_h * (k + 1)], want[k][1]) and np.array_equal(gcr[8 * mb_h * k:8 * mb_h * (k + 1)], want[k][2]), k
        assert np.array_equal(gco[n * k:n * (k + 1)], want[k][3]), k
