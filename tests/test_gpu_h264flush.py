"""The decoder back-end's flush point: ff_h264_flush_pictures_cuda runs everything recorded for a batch of
pictures -- inter prediction, weighted prediction, DC transforms, residual, intra reconstruction, deblocking decisions, loop
filter -- in hl_decode_mb()'s / loop_filter()'s order with one call.  Checked against the CPU checker's functions chained in
the same order, macroblock by macroblock, on the same records."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth, tables
from oracle.loader import ptr
import h264_util as hu
from test_oracle_h264lf_cpu import run as oracle_decisions

pytestmark = pytest.mark.gpu

DC_DT = np.dtype([("luma_qmul", "<u4"), ("chroma_qmul", "<u4", (2,))])


def _dev(a):
    from libav_b200 import device
    return device.DevBuf.from_numpy(np.ascontiguousarray(a))


def picture_work(mb_w, mb_h, seed, p_intra):
    """one picture: records of every stage over one coefficient arena (intra macroblocks take the intra generator's
    coefficients, inter ones the residual generator's)"""
    n = mb_w * mb_h
    r = np.random.RandomState(seed)
    intra, co_i, nz_i = synth.h264_intra_work(mb_w, mb_h, seed=seed + 1, p_intra=p_intra)
    res, co_r, nz_r = synth.h264_residual_work(mb_w, mb_h, seed=seed + 2)
    is_intra = intra["kind"] != 0
    res["luma_mode"][is_intra] = 3; res["chroma"][is_intra] = 0
    coeffs = np.where(is_intra[:, None], co_i, co_r); nnzc = np.where(is_intra[:, None], nz_i, nz_r)
    mc = synth.h264_mc_work(mb_w, mb_h, seed=seed + 3, max_mv=40, avg_second=True)
    dc = np.zeros(n, DC_DT)
    dc["luma_qmul"] = (r.randint(0, 4, n) == 0) * r.randint(16, 4000, n)
    dc["chroma_qmul"] = (r.randint(0, 3, (n, 2)) == 0) * r.randint(16, 4000, (n, 2))
    luma_dc = r.randint(-2000, 2001, (n, 16)).astype(np.int16)
    wrec = np.zeros(max(n // 3, 1), synth.WEIGHT_DT)               # explicit weighting of a third of the macroblocks' luma
    mbs = r.choice(n, wrec.shape[0], replace=False)
    wrec["off"] = (mbs // mb_w) * 16 * (16 * mb_w) + (mbs % mb_w) * 16
    wrec["w"] = 16; wrec["h"] = 16; wrec["log2_denom"] = r.randint(0, 7, wrec.shape[0])
    wrec["weight"] = r.randint(-60, 100, wrec.shape[0]); wrec["offset"] = r.randint(-20, 21, wrec.shape[0])
    # (one set of slice headers per flush: every picture of the batch gets the same side information, its pixels differ)
    info = synth.h264_deblock_info(mb_w, mb_h, seed=mb_w + 4, n_slices=3, bipred=True, t8x8=1, cabac=1)
    return dict(intra=intra, res=res, coeffs=np.ascontiguousarray(coeffs), nnzc=np.ascontiguousarray(nnzc), mc=mc, dc=dc, luma_dc=luma_dc, weight=wrec,
                info=info)


def oracle_chain(o, w, refs, mb_w, mb_h, y, cb, cr):
    hu.oracle_mc(o, w["mc"], refs, y, cb, cr)
    hu.oracle_weight(o, w["weight"], y)
    co = w["coeffs"].copy()
    for m in range(co.shape[0]):
        if w["dc"]["luma_qmul"][m]:
            o.h264_luma_dc_dequant_idct(ptr(co[m]), ptr(w["luma_dc"][m].copy()), int(w["dc"]["luma_qmul"][m]))
        for p in range(2):
            if w["dc"]["chroma_qmul"][m, p]:
                blk = np.ascontiguousarray(co[m, 256 * (p + 1):256 * (p + 2)])
                o.h264_chroma_dc_dequant_idct(ptr(blk), int(w["dc"]["chroma_qmul"][m, p]))
                co[m, 256 * (p + 1):256 * (p + 2)] = blk
    hu.oracle_residual(o, w["res"], co, w["nnzc"], y, cb, cr)
    hu.oracle_intra(o, w["intra"], co, w["nnzc"], mb_w, mb_h, y, cb, cr)
    rec = oracle_decisions(o, w["info"]).view(synth.DEBLOCK_DT).reshape(-1)
    hu.oracle_deblock(o, rec, mb_w, mb_h, y, cb, cr)
    return co


@pytest.mark.parametrize("mb_w,mb_h,P,p_intra", [(7, 5, 1, 0.3), (20, 12, 3, 0.25), (45, 30, 2, 0.1)])
def test_flush_equals_the_chained_checker(gpu, checker, mb_w, mb_h, P, p_intra):
    from libav_b200 import device
    refs = [synth.h264_picture(mb_w, mb_h, seed=11), synth.h264_picture(mb_w, mb_h, seed=12)]
    works = [picture_work(mb_w, mb_h, 100 * k + mb_w, p_intra) for k in range(P)]
    pics = [synth.h264_picture(mb_w, mb_h, seed=50 + k) for k in range(P)]
    want, want_co = [], []
    for w, (y, cb, cr) in zip(works, pics):
        wy, wcb, wcr = y.copy(), cb.copy(), cr.copy()
        want_co.append(oracle_chain(checker, w, refs, mb_w, mb_h, wy, wcb, wcr))
        want.append((wy, wcb, wcr))
    # the batch: pictures stacked vertically, per-macroblock arrays concatenated, record positions shifted to their picture
    Y, CB, CR = (np.concatenate([p[i] for p in pics]) for i in range(3))
    ls, uvls = Y.strides[0], CB.strides[0]
    cat = lambda k: np.concatenate([w[k] for w in works])
    mc, res, wrec = [], [], []
    for k, w in enumerate(works):
        m = w["mc"].copy(); m["y"] += 16 * mb_h * k; mc.append(m)
        r = w["res"].copy(); r["luma_off"] += 16 * mb_h * k * ls; r["chroma_off"] += 8 * mb_h * k * uvls; res.append(r)
        t = w["weight"].copy(); t["off"] += 16 * mb_h * k * ls; wrec.append(t)
    mc, res, wrec = np.concatenate(mc), np.concatenate(res), np.concatenate(wrec)
    mc = np.concatenate([mc[mc["avg"] == 0], mc[mc["avg"] != 0]])            # put records first
    stack_refs = [[np.concatenate([p] * P) for p in r] for r in refs]         # every picture references the same two frames
    dref = [[_dev(p) for p in r] for r in stack_refs]
    d_planes = _dev(np.array([[p.ptr for p in r] for r in dref], dtype=np.uint64))
    infos = [w["info"] for w in works]
    catinfo = lambda k: np.concatenate([d[k] for d in infos], axis=0)
    keep = {k: _dev(catinfo(k)) for k in ("mb_type", "qscale", "nnz", "cbp", "slice_table", "mv0", "mv1", "ref0", "ref1")}
    keep["sp"] = _dev(infos[0]["slice_params"]); keep["cq"] = _dev(infos[0]["chroma_qp_table"])
    info = tables.FFH264DeblockInfo(mb_w, mb_h, P, keep["mb_type"].ptr, keep["qscale"].ptr, keep["nnz"].ptr, keep["cbp"].ptr, keep["slice_table"].ptr,
                                    (C.c_void_p * 2)(keep["mv0"].ptr, keep["mv1"].ptr), (C.c_void_p * 2)(keep["ref0"].ptr, keep["ref1"].ptr),
                                    keep["sp"].ptr, infos[0]["n_slices"], keep["cq"].ptr, infos[0]["cabac"], infos[0]["t8x8"])
    n = mb_w * mb_h * P
    d = dict(y=_dev(Y), cb=_dev(CB), cr=_dev(CR), mc=_dev(mc), res=_dev(res), w=_dev(wrec), co=_dev(cat("coeffs")), nz=_dev(cat("nnzc")),
             dc=_dev(cat("dc")), ldc=_dev(cat("luma_dc")), intra=_dev(cat("intra")), rec=device.DevBuf(n * 104), prog=device.DevBuf(4 * 2 * mb_h * P))
    work = tables.FFH264PictureWork()
    work.mb_w, work.mb_h, work.n_pictures = mb_w, mb_h, P
    work.luma, work.cb, work.cr, work.linesize, work.uvlinesize = d["y"].ptr, d["cb"].ptr, d["cr"].ptr, ls, uvls
    work.mc, work.n_mc, work.refs = d["mc"].ptr, mc.shape[0], d_planes.ptr
    work.weight[0], work.n_weight[0] = d["w"].ptr, wrec.shape[0]
    work.coeffs, work.coeff_stride, work.nnzc = d["co"].ptr, 768, d["nz"].ptr
    work.dc, work.luma_dc, work.residual, work.intra = d["dc"].ptr, d["ldc"].ptr, d["res"].ptr, d["intra"].ptr
    work.deblock_info, work.deblock_records, work.progress = C.pointer(info), d["rec"].ptr, d["prog"].ptr
    gpu.check(gpu.lib.ff_h264_flush_pictures_cuda(C.byref(work), None))
    device.sync()
    gy, gcb, gcr = d["y"].download(np.uint8, Y.shape), d["cb"].download(np.uint8, CB.shape), d["cr"].download(np.uint8, CR.shape)
    assert np.array_equal(gy, np.concatenate([w[0] for w in want])), np.argwhere(gy != np.concatenate([w[0] for w in want]))[:5].tolist()
    assert np.array_equal(gcb, np.concatenate([w[1] for w in want])) and np.array_equal(gcr, np.concatenate([w[2] for w in want]))
    assert np.array_equal(d["co"].download(np.int16, (n, 768)), np.concatenate(want_co))


def test_bad_work_is_refused(gpu):
    work = tables.FFH264PictureWork()
    assert gpu.lib.ff_h264_flush_pictures_cuda(C.byref(work), None) == -1
    gpu.lib.avb200_clear_error()
