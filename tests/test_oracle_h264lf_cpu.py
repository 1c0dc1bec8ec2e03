"""Deblocking DECISIONS: the plain-C restatement (oracle/port/orc_h264lf.c) against the reference's
own h264_loopfilter.c (compiled unmodified into oracle/_ref, slots replaced by recorders) on random but self-consistent
decoder side information -- P and B pictures, several slices, CABAC and CAVLC with the 8x8 transform, both slice-edge
modes, low-qp pictures that take the threshold shortcut, unequal cb/cr qp offsets."""
import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr

CASES = [
    dict(seed=1),
    dict(seed=2, bipred=True),
    dict(seed=3, t8x8=1, cabac=1),
    dict(seed=4, t8x8=1, cabac=0),
    dict(seed=5, t8x8=1, cabac=0, bipred=True, mode=2, n_slices=5),
    dict(seed=6, mode=2, n_slices=7, p_intra=0.4),
    dict(seed=7, qp_lo=0, qp_hi=20),                      # most macroblocks fall under qp_thresh
    dict(seed=8, cb_off=-3, cr_off=5, bipred=True),       # chroma_qp_diff: cb and cr edges get different thresholds
    dict(seed=9, p_intra=0.0, qp_lo=30, qp_hi=40),
    dict(seed=10, mode=0),
]


def run(o, d):
    out = np.zeros((d["mb_w"] * d["mb_h"], 104), np.uint8)
    rc = o.h264_deblock_params(d["mb_w"], d["mb_h"], ptr(d["mb_type"]), ptr(d["qscale"]), ptr(d["nnz"]), ptr(d["cbp"]),
                               ptr(d["slice_table"]), ptr(d["mv0"]), ptr(d["mv1"]), ptr(d["ref0"]), ptr(d["ref1"]),
                               ptr(d["slice_params"]), d["n_slices"], ptr(d["chroma_qp_table"]), d["cabac"], d["t8x8"], ptr(out))
    assert rc == 0
    return out


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s%s" % kv for kv in c.items()))
def test_port_matches_reference_decisions(orc, refo, case):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for (mw, mh) in ((11, 9), (1, 1), (2, 5), (20, 3)):
        d = synth.h264_deblock_info(mw, mh, **case)
        a, b = run(refo, d), run(orc, d)
        bad = np.argwhere((a != b).any(axis=1))
        assert not len(bad), (mw, mh, bad[:4].ravel().tolist(), a[bad[0, 0]].tolist(), b[bad[0, 0]].tolist())


FIELD_CASES = [dict(seed=21), dict(seed=22, bipred=True, t8x8=1, cabac=0, n_slices=4), dict(seed=23, p_intra=0.5, mode=2, n_slices=5)]


def as_field(d):
    """the same side information as one field of a PAFF frame: every macroblock type carries MB_TYPE_INTERLACED (mpegutils.h:58)"""
    d = dict(d)
    d["mb_type"] = d["mb_type"] | np.uint32(0x80) * (d["mb_type"] != 0)
    return d


@pytest.mark.parametrize("case", FIELD_CASES, ids=lambda c: "-".join("%s%s" % kv for kv in c.items()))
def test_field_pictures(orc, refo, case):
    """picture_structure != PICT_FRAME: vertical vector limit 2, bS 3 on horizontal intra macroblock edges (h264_loopfilter.c:551-557,723)"""
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    differs = 0
    for (mw, mh) in ((11, 9), (2, 5), (20, 3)):
        frame = synth.h264_deblock_info(mw, mh, **case)
        d = as_field(frame)
        try:
            refo.h264_deblock_picture_structure(1); orc.h264_deblock_picture_structure(1)
            a, b = run(refo, d), run(orc, d)
        finally:
            refo.h264_deblock_picture_structure(0); orc.h264_deblock_picture_structure(0)
        bad = np.argwhere((a != b).any(axis=1))
        assert not len(bad), (mw, mh, bad[:4].ravel().tolist(), a[bad[0, 0]].tolist(), b[bad[0, 0]].tolist())
        differs += int((a != run(refo, frame)).any())
    assert differs                       # the field rules do change decisions


C422_CASES = [dict(seed=31), dict(seed=32, bipred=True, t8x8=1, cabac=1, n_slices=4), dict(seed=33, t8x8=1, cabac=0, mode=2, n_slices=5),
              dict(seed=34, p_intra=0.5, cb_off=-4, cr_off=6), dict(seed=35, qp_lo=0, qp_hi=24, t8x8=1)]


def run422(o, d):
    """chroma_format_idc 2: (104-byte records, 52-byte records of the four horizontal chroma edges)"""
    ext = np.full((d["mb_w"] * d["mb_h"], 52), 0xEE, np.uint8)
    try:
        o.h264_deblock_chroma422(ptr(ext))
        rec = run(o, d)
    finally:
        o.h264_deblock_chroma422(None)
    return rec, ext


@pytest.mark.parametrize("case", C422_CASES, ids=lambda c: "-".join("%s%s" % kv for kv in c.items()))
def test_chroma422(orc, refo, case):
    """4:2:2: a horizontal chroma edge per luma edge, also inside 8x8-transform macroblocks (h264_loopfilter.c:633,693-700)"""
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    inner = 0
    for (mw, mh) in ((11, 9), (2, 5), (20, 3)):
        d = synth.h264_deblock_info(mw, mh, **case)
        (a, ea), (b, eb) = run422(refo, d), run422(orc, d)
        bad = np.argwhere((a != b).any(axis=1))
        assert not len(bad), (mw, mh, bad[:4].ravel().tolist(), a[bad[0, 0]].tolist(), b[bad[0, 0]].tolist())
        bad = np.argwhere((ea != eb).any(axis=1))
        assert not len(bad), ("ext", mw, mh, bad[:4].ravel().tolist(), ea[bad[0, 0]].tolist(), eb[bad[0, 0]].tolist())
        a420 = run(refo, d)
        assert np.array_equal(a[:, :50], a420[:, :50])                  # luma decisions do not depend on the chroma format
        inner += int(ea[:, [1, 3, 5, 7]].any())                          # alpha of the edges that exist only in 4:2:2
    assert inner


def test_decisions_are_not_trivial(orc):
    d = synth.h264_deblock_info(12, 8, seed=1)
    rec = run(orc, d).view(synth.DEBLOCK_DT).reshape(-1)
    assert (rec["alpha"] != 0).mean() > 0.3 and (rec["alpha"] == 0).mean() > 0.05
    assert rec["intra"].any() and (rec["tc0"] > 0).any() and (rec["tc0"] < 0).any()
