"""MPEG-1/2/4 / H.263 inverse quantisers: the restatement (oracle/port/orc_mpv.c) against the reference's own static
dct_unquantize_*_c reached through ff_mpv_common_init() (oracle/_ref), over random blocks, quantiser scales, coded
lengths, both scans, AIC / AC-prediction flags and full-range levels (int16 wrap of the stored result included)."""
import numpy as np
import pytest

from oracle.loader import ptr

KINDS = {0: "mpeg1_intra", 1: "mpeg1_inter", 2: "mpeg2_intra", 3: "mpeg2_intra_bitexact", 4: "mpeg2_inter", 5: "h263_intra", 6: "h263_inter"}


def random_case(r, kind, wild):
    blk = np.zeros(64, np.int16)
    nz = r.randint(1, 65)
    pos = r.permutation(64)[:nz]
    blk[pos] = r.randint(-32768, 32768, nz) if wild else r.randint(-2047, 2048, nz)
    return dict(block=blk, n=int(r.randint(0, 6)), qscale=int(r.randint(1, 32)), last=int(r.randint(-1 if kind in (1, 4) else 0, 64)),
                ydc=int(r.choice([1, 2, 4, 8, 13, 46])), cdc=int(r.choice([1, 2, 4, 8, 11, 25])),
                intra=r.randint(1, 256, 64).astype(np.uint16), inter=r.randint(1, 256, 64).astype(np.uint16),
                alt=int(r.randint(0, 2)), aic=int(r.randint(0, 2)), acp=int(r.randint(0, 2)))


def apply(o, kind, c):
    b = c["block"].copy()
    if kind in (5, 6) and c["last"] < 0:
        c = dict(c, last=0)                    # the reference asserts block_last_index >= 0 for the H.263 quantisers
    o.mpeg_dequant(kind, ptr(b), c["n"], c["qscale"], c["last"], c["ydc"], c["cdc"], ptr(c["intra"]), ptr(c["inter"]), c["alt"], c["aic"], c["acp"])
    return b


@pytest.mark.parametrize("kind", list(KINDS), ids=list(KINDS.values()))
def test_port_matches_reference(orc, refo, kind):
    if refo is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = np.random.RandomState(100 + kind)
    for it in range(1500):
        c = random_case(r, kind, wild=it % 3 == 0)
        a, b = apply(refo, kind, c), apply(orc, kind, c)
        assert np.array_equal(a, b), (kind, it, c, a.tolist(), b.tolist())


def test_scan_tables(orc, refo):
    for alt in (0, 1):
        p, e = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
        orc.mpeg_scantables(alt, ptr(p), ptr(e))
        assert sorted(p.tolist()) == list(range(64)) and e[63] == 63 and (np.diff(e.astype(int)) >= 0).all()
        if refo is not None:
            p2, e2 = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
            refo.mpeg_scantables(alt, ptr(p2), ptr(e2))
            assert np.array_equal(p, p2) and np.array_equal(e, e2)


def test_synthetic_zigzag_equals_the_scan_table(orc):
    from libav_b200 import synth
    p, e = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
    orc.mpeg_scantables(0, ptr(p), ptr(e))
    sp, se = synth.zigzag_scan_tables()
    assert np.array_equal(sp, p) and np.array_equal(se, e)
