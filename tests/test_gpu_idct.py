"""GPU parity (through the C-ABI): batched simple IDCT family, clamped pixel ops, block clear/fill, the table
slots installed by ff_idctdsp_init_cuda / ff_blockdsp_init_cuda, and the host-buffer end-to-end call -- all
byte-compared with the CPU oracle (the compiled reference when oracle/_ref travelled with the snapshot, else
the plain-C restatement).  Sizes up to BASELINE config 2 (2^20 blocks) are covered by properties that do not
need the oracle at full size: tiling invariance and put/add consistency."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr

pytestmark = pytest.mark.gpu


def _oracle_batch(o, mode, blocks, frame, off, stride):
    b, f = blocks.copy(), frame.copy()
    o.idct_batch(mode, ptr(b), ptr(f), ptr(off), stride, blocks.shape[0], 8)
    return b, f


def _cases():
    rng = np.random.default_rng(11)
    rowdc = np.zeros((1024, 64), dtype=np.int16)
    rowdc[:, ::8] = rng.integers(-1024, 1024, size=(1024, 8))
    rowdc[::2, 9:16] = rng.integers(-300, 300, size=(512, 7))
    return {
        "dct0": synth.dct_test_blocks(0, 20000), "dct1": synth.dct_test_blocks(1, 20000),
        "dct2": synth.dct_test_blocks(2, 20000), "dense": synth.dense_blocks(20000),
        "extreme": rng.integers(-32768, 32768, size=(4096, 64)).astype(np.int16),
        "rowdc": rowdc, "ragged": synth.dense_blocks(1000 + 13, seed=5), "one": synth.dense_blocks(1, seed=9),
    }


@pytest.mark.parametrize("name", list(_cases()))
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_idct_batch_matches_oracle(gpu, checker, name, mode):
    from libav_b200 import device
    blocks = _cases()[name]
    n = blocks.shape[0]
    tpr = 37 if n > 37 else 1
    stride = tpr * 8
    rows = ((n + tpr - 1) // tpr) * 8
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, size=(rows, stride), dtype=np.uint8)
    off = device.tile_offsets(n, tpr, stride)
    want_b, want_f = _oracle_batch(checker, mode, blocks, frame, off, stride)
    for use_off in (False, True):
        got = device.idct_put_tiles(blocks, tpr, mode=mode, frame=frame, use_offsets=use_off)
        assert np.array_equal(got, want_b if mode == 2 else want_f), (name, mode, use_off)


@pytest.mark.parametrize("name", ["dense", "extreme", "ragged", "one"])
def test_cp_async_variant_matches_oracle(gpu, checker, name):
    """idct_put fetches its groups by TMA (cp.async.bulk.tensor, SWIZZLE_128B tensor map) by default; this pins the
    cp.async fetch that the add / in-place / fused-clear modes use, on the put path as well"""
    from libav_b200 import device
    blocks = _cases()[name]
    n = blocks.shape[0]
    tpr = 37 if n > 37 else 1
    stride = tpr * 8
    frame = np.random.default_rng(2).integers(0, 256, size=(((n + tpr - 1) // tpr) * 8, stride), dtype=np.uint8)
    off = device.tile_offsets(n, tpr, stride)
    _, want = _oracle_batch(checker, 0, blocks, frame, off, stride)
    gpu.lib.avb200_set_tuning(b"idct_tma", 2)
    try:
        got = device.idct_put_tiles(blocks, tpr, mode=0, frame=frame, use_offsets=False)
    finally:
        gpu.lib.avb200_set_tuning(b"idct_tma", 0)
    assert np.array_equal(got, want), name


def test_fused_clear_zeroes_coefficients(gpu, checker):
    from libav_b200 import device
    blocks = synth.dense_blocks(5000, seed=3)
    off = device.tile_offsets(5000, 50, 400)
    _, want = _oracle_batch(checker, 0, blocks, np.zeros((800, 400), np.uint8), off, 400)
    got, after = device.idct_put_tiles(blocks, 50, mode=0, clear=True)
    assert np.array_equal(got, want)
    assert not after.any()


def test_empty_batch_is_a_noop(gpu):
    L = gpu
    assert L.lib.ff_simple_idct_batch_cuda(0, None, None, None, 8, 0, 1, 0, None) == 0
    assert L.lib.ff_pixels_clamped_batch_cuda(0, None, None, None, 8, 0, 1, None) == 0
    assert L.last_error() == ""


def test_bad_arguments_fail_loudly(gpu):
    L = gpu
    assert L.lib.ff_simple_idct_batch_cuda(7, None, None, None, 8, 4, 1, 0, None) == -1
    assert "bad mode" in L.last_error()
    L.lib.avb200_clear_error()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pixels_clamped_match_oracle(gpu, checker, mode):
    from libav_b200 import device
    rng = np.random.default_rng(mode)
    n, tpr = 3001, 61
    blocks = rng.integers(-700, 700, size=(n, 64)).astype(np.int16)
    stride, rows = tpr * 8, ((n + tpr - 1) // tpr) * 8
    frame = rng.integers(0, 256, size=(rows, stride), dtype=np.uint8)
    want = frame.copy()
    fn = [checker.put_pixels_clamped, checker.put_signed_pixels_clamped, checker.add_pixels_clamped][mode]
    off = device.tile_offsets(n, tpr, stride)
    for i in range(n):
        fn(ptr(blocks[i]), C.c_void_p(want.ctypes.data + int(off[i])), stride)
    d_b, d_f = device.DevBuf.from_numpy(blocks), device.DevBuf.from_numpy(frame)
    gpu.check(gpu.lib.ff_pixels_clamped_batch_cuda(mode, d_b.ptr, d_f.ptr, None, stride, n, tpr, None))
    device.sync()
    assert np.array_equal(d_f.download(np.uint8, frame.shape), want)


def test_table_slots_drop_in(gpu, checker):
    """ff_idctdsp_init_cuda / ff_blockdsp_init_cuda fill the reference's tables; call the slots like a codec."""
    from libav_b200 import tables
    c = tables.IDCTDSPContext()
    gpu.lib.ff_idctdsp_init_cuda(C.byref(c), tables.FF_IDCT_SIMPLE, 8, 0)
    assert c.perm_type == 0 and list(c.idct_permutation) == list(range(64))
    rng = np.random.default_rng(4)
    blocks = synth.dct_test_blocks(0, 64)
    for i in range(64):
        for slot, oname in ((c.idct_put, "simple_idct_put"), (c.idct_add, "simple_idct_add")):
            pix = rng.integers(0, 256, size=(8, 40), dtype=np.uint8)
            a, b = pix.copy(), pix.copy()
            blk = np.ascontiguousarray(blocks[i].copy())
            slot(C.cast(a.ctypes.data + 16, C.POINTER(C.c_uint8)), 40, blk.ctypes.data_as(C.POINTER(C.c_int16)))
            getattr(checker, oname)(C.c_void_p(b.ctypes.data + 16), 40, ptr(blocks[i].copy()))
            assert np.array_equal(a, b), oname
        blk = np.ascontiguousarray(blocks[i].copy()); want = blocks[i].copy()
        c.idct(blk.ctypes.data_as(C.POINTER(C.c_int16)))
        checker.simple_idct(ptr(want))
        assert np.array_equal(blk, want)
        for slot, oname in ((c.put_pixels_clamped, "put_pixels_clamped"), (c.put_signed_pixels_clamped, "put_signed_pixels_clamped"),
                            (c.add_pixels_clamped, "add_pixels_clamped")):
            pix = rng.integers(0, 256, size=(8, 24), dtype=np.uint8)
            a, b = pix.copy(), pix.copy()
            blk = rng.integers(-600, 600, size=64).astype(np.int16)
            slot(blk.ctypes.data_as(C.POINTER(C.c_int16)), C.cast(a.ctypes.data, C.POINTER(C.c_uint8)), 24)
            getattr(checker, oname)(ptr(blk), ptr(b), 24)
            assert np.array_equal(a, b), oname
    # negative line size (bottom-up picture)
    pix = np.zeros((8, 8), np.uint8); want = np.zeros((8, 8), np.uint8)
    blk = np.ascontiguousarray(blocks[3].copy())
    c.idct_put(C.cast(pix.ctypes.data + 56, C.POINTER(C.c_uint8)), -8, blk.ctypes.data_as(C.POINTER(C.c_int16)))
    checker.simple_idct_put(C.c_void_p(want.ctypes.data + 56), -8, ptr(blocks[3].copy()))
    assert np.array_equal(pix, want)
    b = tables.BlockDSPContext()
    gpu.lib.ff_blockdsp_init_cuda(C.byref(b))
    z = rng.integers(-5, 5, size=6 * 64).astype(np.int16)
    b.clear_blocks(z.ctypes.data_as(C.POINTER(C.c_int16)))
    assert not z.any()
    z = rng.integers(1, 5, size=128).astype(np.int16)
    b.clear_block(z.ctypes.data_as(C.POINTER(C.c_int16)))
    assert not z[:64].any() and z[64:].all()
    pix = np.zeros((16, 32), np.uint8)
    b.fill_block_tab[0](C.cast(pix.ctypes.data, C.POINTER(C.c_uint8)), 77, 32, 16)
    b.fill_block_tab[1](C.cast(pix.ctypes.data + 16, C.POINTER(C.c_uint8)), 99, 32, 8)
    assert (pix[:, :16] == 77).all() and (pix[:8, 16:24] == 99).all() and not pix[8:, 16:].any() and not pix[:, 24:].any()
    assert gpu.last_error() == ""


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_host_buffer_end_to_end(gpu, checker, mode):
    from libav_b200 import device
    n, tpr = 200000, 512
    blocks = synth.tile_large(synth.dense_blocks(4096, seed=8), n)
    stride, rows = tpr * 8, ((n + tpr - 1) // tpr) * 8
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, size=(rows, stride), dtype=np.uint8)
    off = device.tile_offsets(n, tpr, stride)
    want_b, want_f = _oracle_batch(checker, mode, blocks, frame, off, stride)
    for use_off in (False, True):
        b, f = blocks.copy(), frame.copy()
        gpu.check(gpu.lib.ff_simple_idct_batch_host_cuda(mode, ptr(b), ptr(f), f.nbytes, ptr(off) if use_off else None, stride, n, tpr))
        assert np.array_equal(b if mode == 2 else f, want_b if mode == 2 else want_f), (mode, use_off)


def test_full_size_properties(gpu, checker):
    """BASELINE config 2 size (2^20 blocks): the oracle checks a 1/64 sample; the rest is covered by
    (1) tiling invariance -- every copy of the same block yields the same 8x8 tile, and
    (2) put/add consistency -- idct_add onto a zero frame equals idct_put where no clipping below 0 occurs."""
    from libav_b200 import device
    n, tpr = 1 << 20, 1024
    base = synth.dct_test_blocks(0, 16384)
    blocks = synth.tile_large(base, n)
    frame = device.idct_put_tiles(blocks, tpr)
    tiles = frame.reshape(n // tpr, 8, tpr, 8).transpose(0, 2, 1, 3).reshape(n, 64)
    assert np.array_equal(tiles[:16384], tiles[16384:32768]) and np.array_equal(tiles[:16384], tiles[-16384:])
    off = np.arange(16384, dtype=np.uint32) * 64
    _, want = _oracle_batch(checker, 0, base, np.zeros(16384 * 64, np.uint8), off, 8)
    assert np.array_equal(tiles[:16384].reshape(-1), want)
    added = device.idct_put_tiles(blocks, tpr, mode=1, frame=np.zeros_like(frame))
    assert np.array_equal(added, frame)


@pytest.mark.parametrize("gen", ["dense", "dct0", "sparse"])
def test_config2_every_block_against_the_checker(gpu, checker, gen):
    """BASELINE config 2 at its full size, every one of the 2^20 blocks compared with the CPU checker (all host threads), for the
    three distributions SURVEY 8d names: dense uniform, dct.c test 0, dct.c test 1 (sparse)"""
    from libav_b200 import device
    import os
    n, tpr = 1 << 20, 1024
    if gen == "dense":
        blocks = synth.dense_blocks(n, seed=3)
    else:
        blocks = synth.tile_large(synth.dct_test_blocks(0 if gen == "dct0" else 1, 1 << 16), n)
        blocks = np.ascontiguousarray(blocks + (np.arange(n, dtype=np.int16)[:, None] & 1))      # every copy differs from its neighbours
    stride = tpr * 8
    off = device.tile_offsets(n, tpr, stride)
    want = np.zeros(((n // tpr) * 8, stride), np.uint8)
    checker.idct_batch(0, ptr(blocks.copy()), ptr(want), ptr(off), stride, n, os.cpu_count() or 1)
    got = device.idct_put_tiles(blocks, tpr)
    assert np.array_equal(got, want)
