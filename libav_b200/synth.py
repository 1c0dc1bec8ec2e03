"""Synthetic inputs shaped like the reference's own test generators (numpy, host side).

  lfg(seed, n)                 av_lfg_init + av_lfg_get stream      libavutil/lfg.c:30-46, lfg.h:40-43
  dct_test_blocks(test, n)     libavcodec/tests/dct.c:100-126 init_block() for the IDCT tests
  ref_fdct / ref_idct          libavcodec/dctref.c:58-122 (double precision, same summation order)
  dense_blocks(n)              uniform [-256,255] coefficients
  yuv420p_frame(w, h, seed)    Y from the LFG first, then U[i], V[i] alternately, low byte of each draw
                               (the fill the survey probe used; CRC-pinned in tests/golden)
Used by tests/ and bench.py only; nothing here is on the product path.
"""
import hashlib
import struct

import numpy as np


def lfg(seed, n):
    """First n outputs of av_lfg_get() after av_lfg_init(seed) as uint32."""
    st = np.zeros(64, dtype=np.uint32)
    tmp = bytearray(16)
    for i in range(8, 64, 4):
        tmp[0:4] = struct.pack("<I", seed & 0xFFFFFFFF)
        tmp[4] = i
        tmp = bytearray(hashlib.md5(bytes(tmp)).digest())
        st[i:i + 4] = np.frombuffer(bytes(tmp), dtype="<u4")
    out = np.empty(n + 64, dtype=np.uint32)
    out[:64] = st
    k, total = 64, n + 64
    with np.errstate(over="ignore"):
        while k < total:
            m = min(24, total - k)          # x[k] = x[k-24] + x[k-55]; 24 outputs are independent
            out[k:k + m] = out[k - 24:k - 24 + m] + out[k - 55:k - 55 + m]
            k += m
    return out[64:]


def _dct_coefficients():
    c = np.zeros((8, 8), dtype=np.float64)
    for j in range(8):
        c[0, j] = np.sqrt(0.125)
        for i in range(1, 8):
            c[i, j] = 0.5 * np.cos((8 * i) * (j + 0.5) * np.pi / 64.0)
    return c


_COEF = _dct_coefficients()


def ref_fdct(blocks):
    """ff_ref_fdct on an (n, 8, 8) integer array -> int16, sequential-k accumulation like dctref.c:58-87."""
    b = blocks.astype(np.float64)
    n = b.shape[0]
    out = np.zeros((n, 8, 8), dtype=np.float64)
    for k in range(8):                       # out[i][j] = 8 * sum_k coef[i][k] * block[k][j]
        out += _COEF[None, :, k, None] * b[:, None, k, :]
    out *= 8
    res = np.zeros((n, 8, 8), dtype=np.float64)
    for k in range(8):                       # block[i][j] = sum_k out[i][k] * coef[j][k]
        res += out[:, :, k, None] * _COEF[None, None, :, k]
    return np.floor(res + 0.499999999999).astype(np.int64).astype(np.int16)


def ref_idct(blocks):
    """ff_ref_idct (dctref.c:96-122) on (n, 8, 8) -> int16."""
    b = blocks.astype(np.float64)
    n = b.shape[0]
    out = np.zeros((n, 8, 8), dtype=np.float64)
    for k in range(8):                       # out[i][j] = sum_k block[i][k] * coef[k][j]
        out += b[:, :, k, None] * _COEF[None, None, k, :]
    res = np.zeros((n, 8, 8), dtype=np.float64)
    for k in range(8):                       # block[i][j] = sum_k coef[k][i] * out[k][j]
        res += _COEF[None, k, :, None] * out[:, None, k, :]
    return np.floor(res + 0.5).astype(np.int64).astype(np.int16)


def dct_test_blocks(test, n, seed=1):
    """n blocks drawn exactly like NB_ITS calls of init_block(block, test, is_idct=1, prng) -> (n, 64) int16."""
    if test == 0:
        r = lfg(seed, 64 * n).reshape(n, 64)
        pix = (r % 512).astype(np.int64) - 256
        co = ref_fdct(pix.reshape(n, 8, 8)).astype(np.int32) >> 3
        return co.reshape(n, 64).astype(np.int16)
    if test == 1:
        # variable draws per block: j = r%10+1 then j x (value, position).  `block[lfg()%64] = lfg()%512-256`
        # leaves the draw order to the compiler; gcc evaluates the right-hand side first, and that order is
        # the one that reproduces fate-idct8x8's published statistics (tests/golden/fate_idct8x8.txt).
        r = lfg(seed, 21 * n + 64)
        out = np.zeros((n, 64), dtype=np.int16)
        p = 0
        for b in range(n):
            j = int(r[p] % 10) + 1
            p += 1
            for _ in range(j):
                val = int(r[p] % 512) - 256
                pos = int(r[p + 1] % 64)
                p += 2
                out[b, pos] = val
        return out
    if test == 2:
        r = lfg(seed, n)
        out = np.zeros((n, 64), dtype=np.int16)
        out[:, 0] = ((r % 4096).astype(np.int64) - 2048).astype(np.int16)
        out[:, 63] = (out[:, 0] & 1) ^ 1
        return out
    raise ValueError("test must be 0, 1 or 2")


def dense_blocks(n, seed=1):
    """Every coefficient uniform in [-256, 255]."""
    r = lfg(seed, 64 * n)
    return ((r % 512).astype(np.int64) - 256).astype(np.int16).reshape(n, 64)


def tile_large(base, n):
    """Repeat an (m, 64) block set up to n blocks with a cheap per-copy perturbation-free tiling (the
    distribution, not the exact stream, matters for throughput runs at 2^20 blocks)."""
    m = base.shape[0]
    reps = (n + m - 1) // m
    return np.tile(base, (reps, 1))[:n].copy()


def yuv420p_frame(w, h, seed=1):
    """(Y, U, V) uint8 planes, tight strides; chroma planes are ceil(w/2) x ceil(h/2)."""
    cw, ch = (w + 1) // 2, (h + 1) // 2
    r = lfg(seed, w * h + 2 * cw * ch)
    y = (r[:w * h] & 0xFF).astype(np.uint8).reshape(h, w)
    uv = (r[w * h:] & 0xFF).astype(np.uint8)
    u = uv[0::2].reshape(ch, cw).copy()
    v = uv[1::2].reshape(ch, cw).copy()
    return y, u, v


def pad_rows(plane, pad=8):
    """Copy a plane into rows of width+pad bytes whose padding repeats the last pixel (the byte the reference's
    fast-bilinear scaler reads past the end of each row); returns the strided view of the valid region."""
    h, w = plane.shape
    buf = np.empty((h, w + pad), dtype=plane.dtype)
    buf[:, :w] = plane
    buf[:, w:] = plane[:, -1:]
    return buf[:, :w]


def crc32_ieee_be(data):
    """av_crc(av_crc_get_table(AV_CRC_32_IEEE), 0, buf, len) (libavutil/crc.c): MSB-first CRC-32, polynomial
    0x04C11DB7, init 0, no final xor; av_crc keeps its state byte-swapped, so the value it returns is the
    byte-swapped register.  Chunk-parallel numpy evaluation (CRC with zero init is linear in the message)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    tab = _crc_table()
    n = buf.size
    if n == 0:
        return 0
    chunks = max(1, min(4096, n // 64))
    clen = -(-n // chunks)
    pad = chunks * clen - n
    m = np.concatenate([np.zeros(pad, np.uint8), buf]).reshape(chunks, clen)   # leading zeros are neutral
    crc = np.zeros(chunks, dtype=np.uint32)
    for k in range(clen):
        crc = ((crc << np.uint32(8)) ^ tab[((crc >> np.uint32(24)) ^ m[:, k]) & 0xFF]).astype(np.uint32)
    # operator "append clen zero bytes" as four byte tables
    basis = (np.arange(256, dtype=np.uint32)[None, :] << (np.arange(4, dtype=np.uint32)[:, None] * 8)).astype(np.uint32).reshape(-1)
    for _ in range(clen):
        basis = ((basis << np.uint32(8)) ^ tab[(basis >> np.uint32(24)) & 0xFF]).astype(np.uint32)
    basis = basis.reshape(4, 256)
    acc = 0
    for c in crc.tolist():
        acc = int(basis[0][acc & 0xFF] ^ basis[1][(acc >> 8) & 0xFF] ^ basis[2][(acc >> 16) & 0xFF] ^ basis[3][acc >> 24]) ^ c
    return int.from_bytes(acc.to_bytes(4, "big"), "little")


_CRC_TAB = None


def _crc_table():
    global _CRC_TAB
    if _CRC_TAB is None:
        t = []
        for i in range(256):
            c = i << 24
            for _ in range(8):
                c = ((c << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if c & 0x80000000 else (c << 1) & 0xFFFFFFFF
            t.append(c)
        _CRC_TAB = np.array(t, dtype=np.uint32)
    return _CRC_TAB
