"""Cases for the 10-bit simple IDCT (ff_simple_idct_put_10 / _add_10 / _10), shared by the host simulation and the GPU tests: the three
table entries with host pointers and the batched entry point, against the oracle (outputs AND the blocks as the C functions leave them).
TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from oracle.loader import ptr


def blocks(rng, n):
    """coefficients in the range 10-bit content produces (DC up to +-8000, AC a few thousand): dense, sparse, DC only, DC-only rows"""
    b = np.zeros((n, 64), np.int16)
    for i in range(n):
        kind = i % 4
        if kind == 0:
            b[i] = rng.integers(-600, 600, 64)
        elif kind == 1:
            idx = rng.integers(0, 64, size=rng.integers(1, 8))
            b[i, idx] = rng.integers(-2000, 2000, size=len(idx))
        elif kind == 2:
            b[i, 0] = rng.integers(-8000, 8000)
        else:
            b[i, :8] = rng.integers(-1500, 1500, 8)
            b[i, 0] = rng.integers(-8000, 8000)
            b[i, 8 * rng.integers(1, 8)] = rng.integers(-3000, 3000)
    return b


def slot_cases(table, checker, seed=0):
    """table: an IDCTDSPContext filled for bits_per_raw_sample 10"""
    rng = np.random.default_rng(seed)
    i16p, u8p = C.POINTER(C.c_int16), C.POINTER(C.c_uint8)
    n = 0
    for blk in blocks(rng, 120):
        pix = rng.integers(0, 1024, size=(10, 24)).astype(np.uint16)
        for mode, slot in enumerate((table.idct_put, table.idct_add, table.idct)):
            a, b, pa, pb = blk.copy(), blk.copy(), pix.copy(), pix.copy()
            if mode == 2:
                slot(C.cast(a.ctypes.data, i16p))
            else:
                slot(C.cast(pa.ctypes.data + 48 + 8, u8p), 48, C.cast(a.ctypes.data, i16p))
            checker.simple_idct10(mode, C.c_void_p(pb.ctypes.data + 48 + 8), 48, ptr(b))
            assert np.array_equal(a, b) and np.array_equal(pa, pb), (mode, n)
            n += 1
    assert table.perm_type == 0 and list(table.idct_permutation) == list(range(64))
    return n


def batch_case(run_batch, checker, mode, n=5000, seed=1, pad=4, shift=0):
    """run_batch(mode, blocks (n, 64) int16, frame (rows, 8 * tiles) uint16, dst_off uint32[n]) -> (blocks, frame) after the call.
    pad 4: a pitch that is not a multiple of 16 bytes (thread-per-block kernel); pad 8: 16-byte rows (the staged kernel), `shift` samples
    of offset on every third block put those on its sample-store path"""
    rng = np.random.default_rng(seed + mode)
    blk = blocks(rng, n)
    tiles = 64
    rows = 8 * ((n + tiles - 1) // tiles)
    frame = rng.integers(0, 1024, size=(rows, 8 * tiles + pad)).astype(np.uint16)
    off = np.array([(i // tiles) * 8 * frame.strides[0] + (i % tiles) * 16 + (2 * shift if i % 3 == 1 and i % tiles < tiles - 1 else 0) for i in range(n)], np.uint32)
    if shift:       # shifted blocks overlap their right neighbour: keep one of each pair (the C order would decide otherwise)
        keep = np.array([i % 3 != 2 or i % tiles == 0 for i in range(n)])
        blk, off = np.ascontiguousarray(blk[keep]), np.ascontiguousarray(off[keep])
        n = len(blk)
    got_b, got_f = run_batch(mode, blk.copy(), frame.copy(), off)
    want_b, want_f = blk.copy(), frame.copy()
    for i in range(n):
        checker.simple_idct10(mode, C.c_void_p(want_f.ctypes.data + int(off[i])), frame.strides[0], C.c_void_p(want_b.ctypes.data + 128 * i))
    assert np.array_equal(got_b, want_b) and np.array_equal(got_f, want_f), mode
    assert mode == 2 or not np.array_equal(got_f, frame)


def fdct10_cases(table, run_batch, checker, seed=0):
    """FDCTDSPContext for bits_per_raw_sample 10 (ff_jpeg_fdct_islow_10 / ff_fdct248_islow_10): the two slots, and ff_fdct_batch_cuda(4 | 5);
    run_batch(which, blocks (n, 64) int16) -> blocks after the call"""
    rng = np.random.default_rng(seed)
    i16p = C.POINTER(C.c_int16)
    for it in range(60):
        blk = (rng.integers(-1023, 1024, 64) if it % 3 else rng.integers(0, 1024, 64)).astype(np.int16)
        for slot, which in ((table.fdct, 4), (table.fdct248, 5)):
            a, b = blk.copy(), blk.copy()
            slot(C.cast(a.ctypes.data, i16p))
            checker.fdct(which, ptr(b))
            assert np.array_equal(a, b), (which, it)
    for which in (4, 5):
        blk = rng.integers(-1023, 1024, size=(3000, 64)).astype(np.int16)
        want = blk.copy()
        for i in range(len(want)):
            checker.fdct(which, C.c_void_p(want.ctypes.data + 128 * i))
        assert np.array_equal(run_batch(which, blk.copy()), want), which
