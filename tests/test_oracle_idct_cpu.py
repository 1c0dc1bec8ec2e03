"""CPU: pin the oracle port of the 8x8 IDCT family (oracle/port/orc_idct.c) against
 (a) the reference's own known-answer self test `libavcodec/tests/dct -i` (fate-idct8x8,
     tests/fate/libavcodec.mak:6-9): error statistics reproduced to 8 decimals
     (tests/golden/fate_idct8x8.txt, produced by oracle/refbuild/run_dct_selftest.sh), and
 (b) the unmodified reference compiled into oracle/_ref, byte for byte, when it is built."""
import os

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fate_idct8x8.txt")
NB_ITS = 20000


def _golden():
    out = {}
    for line in open(GOLD):
        if line.startswith("test"):
            f = line.split()
            out[int(f[0][4:])] = dict(kv.split("=") for kv in f[1:])
    return out


def _idct_all(o, blocks):
    b = np.ascontiguousarray(blocks.copy())
    o.idct_batch(2, ptr(b), None, None, 0, b.shape[0], 1)
    return b


@pytest.mark.parametrize("test", [0, 1, 2])
def test_port_reproduces_fate_idct8x8_statistics(orc, test):
    g = _golden()[test]
    blocks = synth.dct_test_blocks(test, NB_ITS)
    got = _idct_all(orc, blocks).astype(np.int64)
    want = synth.ref_idct(blocks.reshape(-1, 8, 8)).reshape(-1, 64).astype(np.int64)
    err = got - want
    ppe = int(np.abs(err).max())
    omse = float((err * err).sum()) / NB_ITS / 64
    ome = float(err.sum()) / NB_ITS / 64
    assert ppe == int(g["ppe"])
    assert "%0.8f" % omse == g["omse"]
    assert "%0.8f" % ome == g["ome"]
    assert int(np.abs(got).max()) == int(g["maxout"])
    # the spec thresholds the reference test enforces (libavcodec/tests/dct.c:218)
    assert ppe <= 1 and omse <= 0.02 and abs(ome) <= 0.0015


def _cases():
    rng = np.random.default_rng(7)
    ext = rng.integers(-32768, 32768, size=(512, 64)).astype(np.int16)     # wrap-around territory
    dc = np.zeros((256, 64), dtype=np.int16)
    dc[:, 0] = np.arange(-2048, 2048, 16)
    rowdc = np.zeros((256, 64), dtype=np.int16)                              # DC-only rows next to dense rows
    rowdc[:, ::8] = rng.integers(-1024, 1024, size=(256, 8))
    rowdc[::2, 9:16] = rng.integers(-300, 300, size=(128, 7))
    return {
        "dct0": synth.dct_test_blocks(0, 4096), "dct1": synth.dct_test_blocks(1, 4096),
        "dct2": synth.dct_test_blocks(2, 4096), "dense": synth.dense_blocks(4096),
        "extreme": ext, "dc": dc, "rowdc": rowdc, "zero": np.zeros((8, 64), dtype=np.int16),
    }


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_port_matches_compiled_reference(orc, refo, name):
    blocks = _cases()[name]
    n = blocks.shape[0]
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, size=(n * 8, 8), dtype=np.uint8)
    off = (np.arange(n, dtype=np.uint32) * 64)
    for mode in (0, 1, 2):
        res = []
        for o in (orc, refo):
            b = blocks.copy()
            f = base.copy()
            o.idct_batch(mode, ptr(b), ptr(f), ptr(off), 8, n, 1)
            res.append(b if mode == 2 else f)
        assert np.array_equal(res[0], res[1]), (name, mode)


def test_clamp_helpers_match_reference(orc, refo):
    rng = np.random.default_rng(5)
    for _ in range(64):
        blk = rng.integers(-700, 700, size=64).astype(np.int16)
        for fn in ("put_pixels_clamped", "put_signed_pixels_clamped", "add_pixels_clamped"):
            pix = rng.integers(0, 256, size=(8, 24), dtype=np.uint8)
            a, b = pix.copy(), pix.copy()
            getattr(orc, fn)(ptr(blk), ptr(a), 24)
            getattr(refo, fn)(ptr(blk), ptr(b), 24)
            assert np.array_equal(a, b), fn
            assert np.array_equal(a[:, 8:], pix[:, 8:])      # nothing outside the 8x8 is touched


def test_threaded_batch_equals_serial(orc):
    blocks = synth.dense_blocks(1000)
    off = np.arange(1000, dtype=np.uint32) * 64
    outs = []
    for nt in (1, 3, 8):
        b, f = blocks.copy(), np.zeros(64000, dtype=np.uint8)
        orc.idct_batch(0, ptr(b), ptr(f), ptr(off), 8, 1000, nt)
        outs.append(f)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
