"""GPU parity for the libswscale boundary: sws_getContext_cuda / sws_scale_cuda (host pointers) and the batched
device-pointer call, byte-compared with the CPU oracle on identical synthetic frames -- config 1 (640x480) with its
known-answer CRC, the 4K benchmark geometry, true rescales with every integer scaler algorithm, odd sizes, bgr24 and
planar output; plus frame-batch invariance at the full config-5 batch."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr

pytestmark = pytest.mark.gpu
ACC = 0x40000 | 0x80000


def _planes(arrs):
    return (C.c_void_p * 3)(*[a.ctypes.data for a in arrs]), (C.c_int * 3)(*[a.strides[0] for a in arrs])


def oracle_rgb(o, yuv, dw, dh, flags, pad=0):
    dst = np.zeros((dh, dw * 3 + pad), dtype=np.uint8)
    p, s = _planes(yuv)
    assert o.sws_yuv420p_to_rgb24(p, s, yuv[0].shape[1], yuv[0].shape[0], ptr(dst), dst.strides[0], dw, dh, flags) == dh
    return dst


def oracle_yuv(o, yuv, dw, dh, flags):
    out = [np.zeros((dh, dw), np.uint8), np.zeros(((dh + 1) // 2, (dw + 1) // 2), np.uint8), np.zeros(((dh + 1) // 2, (dw + 1) // 2), np.uint8)]
    p, s = _planes(yuv)
    dp, ds = _planes(out)
    assert o.sws_yuv420p_to_yuv420p(p, s, yuv[0].shape[1], yuv[0].shape[0], dp, ds, dw, dh, flags) == dh
    return out


@pytest.mark.parametrize("w,h,crc", [(640, 480, 0xF046D710), (3840, 2160, 0xE1558C0A)])
def test_same_size_rgb24_known_answer(gpu, checker, w, h, crc):
    from libav_b200 import device
    yuv = synth.yuv420p_frame(w, h, 1)
    ctx = device.SwsContext(w, h, w, h)
    assert ctx.fused
    got = ctx.scale(yuv)
    assert synth.crc32_ieee_be(got.tobytes()) == crc
    assert np.array_equal(got, oracle_rgb(checker, yuv, w, h, 4 | ACC))


ALGOS = {"fast_bilinear": 1, "bilinear": 2, "bicubic": 4, "point": 0x10, "area": 0x20, "bicublin": 0x40, "lanczos": 0x200, "gauss": 0x80}
GEOMS = [(352, 288, 640, 480), (640, 480, 352, 288), (96, 96, 64, 128), (100, 37, 333, 211), (1000, 700, 123, 77),
         (64, 48, 64, 48), (66, 50, 33, 25), (641, 479, 641, 479), (1920, 1080, 3840, 2160)]


@pytest.mark.parametrize("variant,pad", [("tile", 6), ("tile", 8), ("two_pass", 6)])
@pytest.mark.parametrize("algo", list(ALGOS))
def test_scaling_matches_oracle(gpu, checker, algo, variant, pad):
    """both general-path implementations (fused per output tile = default; two passes through int16 line planes = the
    fallback for line windows that do not fit in shared memory), byte-store (pad 6) and vector-store (pad 8) rows"""
    from libav_b200 import device
    from libav_b200._lib import lib
    flags = ALGOS[algo] | ACC
    lib.avb200_set_tuning(b"sws_general_variant", 1 if variant == "two_pass" else 0)
    try:
        for (sw, sh, dw, dh) in GEOMS:
            if sw * sh > 1500 * 1000 and algo not in ("bicubic", "bilinear"):
                continue
            yuv = tuple(synth.pad_rows(pl) for pl in synth.yuv420p_frame(sw, sh, 3))
            for fmt in (device.PIX_FMT_RGB24, device.PIX_FMT_BGR24, device.PIX_FMT_YUV420P):
                if fmt == device.PIX_FMT_YUV420P and pad == 8:
                    continue
                ctx = device.SwsContext(sw, sh, dw, dh, fmt, flags)
                got = ctx.scale(yuv, dst_pad=pad) if fmt != device.PIX_FMT_YUV420P else ctx.scale(yuv)
                if fmt == device.PIX_FMT_YUV420P:
                    for a, b in zip(got, oracle_yuv(checker, yuv, dw, dh, flags)):
                        assert np.array_equal(a, b), (algo, sw, sh, dw, dh, "yuv")
                else:
                    want = oracle_rgb(checker, yuv, dw, dh, flags, pad=pad)
                    if fmt == device.PIX_FMT_BGR24:
                        w3 = want[:, :((dw + 1) // 2) * 6].reshape(dh, -1, 3)[:, :, ::-1].reshape(dh, -1)
                        want = np.concatenate([w3, want[:, ((dw + 1) // 2) * 6:]], axis=1)
                    assert np.array_equal(got, want), (algo, sw, sh, dw, dh, fmt)
                ctx.close()
    finally:
        lib.avb200_set_tuning(b"sws_general_variant", 0)


def test_heavy_downscale_falls_back_to_two_passes(gpu, checker):
    """a 16-row output tile of a 12x vertical reduction needs > 96 KB of source lines: the context must still be bit-exact"""
    from libav_b200 import device
    sw, sh, dw, dh = 256, 4800, 128, 100
    yuv = tuple(synth.pad_rows(pl) for pl in synth.yuv420p_frame(sw, sh, 9))
    ctx = device.SwsContext(sw, sh, dw, dh, device.PIX_FMT_RGB24, 4 | ACC)
    assert np.array_equal(ctx.scale(yuv, dst_pad=8), oracle_rgb(checker, yuv, dw, dh, 4 | ACC, pad=8))


def test_unsupported_requests_fail_loudly(gpu):
    L = gpu
    assert not L.lib.sws_getContext_cuda(640, 480, 9, 640, 480, 2, 4 | ACC, None, None, None)     # monowhite: not a source that is taken over
    assert "sources taken over" in L.last_error()
    L.lib.avb200_clear_error()
    assert not L.lib.sws_getContext_cuda(2, 2, 0, 640, 480, 2, 4 | ACC, None, None, None)
    L.lib.avb200_clear_error()
    assert not L.lib.sws_getContext_cuda(640, 480, 0, 320, 240, 2, 4 | ACC | 0x10000, None, None, None)     # SWS_SRC_V_CHR_DROP
    assert "CHR_DROP" in L.last_error()
    L.lib.avb200_clear_error()


def test_device_batch_equals_single_frames(gpu, checker):
    """config 5 shape: K frames of 4K in one launch; every frame must equal its single-frame result, and frame 0
    the oracle's."""
    from libav_b200 import device
    w, h, K = 3840, 2160, 4
    frames = [synth.yuv420p_frame(w, h, 1 + k) for k in range(K)]
    ysz, csz, osz = w * h, (w // 2) * (h // 2), w * h * 3
    d_y = device.DevBuf(ysz * K); d_u = device.DevBuf(csz * K); d_v = device.DevBuf(csz * K); d_o = device.DevBuf(osz * K)
    d_y.upload(np.concatenate([f[0].reshape(-1) for f in frames]))
    d_u.upload(np.concatenate([f[1].reshape(-1) for f in frames]))
    d_v.upload(np.concatenate([f[2].reshape(-1) for f in frames]))
    ctx = device.SwsContext(w, h, w, h)
    ctx.scale_device([d_y.ptr, d_u.ptr, d_v.ptr], [w, w // 2, w // 2], [d_o.ptr], [w * 3], nframes=K,
                     src_frame=[ysz, csz, csz], dst_frame=[osz])
    device.sync()
    out = d_o.download(np.uint8, (K, h, w * 3))
    assert np.array_equal(out[0], oracle_rgb(checker, frames[0], w, h, 4 | ACC))
    for k in range(1, K):
        assert np.array_equal(out[k], ctx.scale(frames[k])), k


def test_unscaled_table_converter(gpu, checker):
    """same size, no SWS_ACCURATE_RND, even height -> the reference's unscaled yuv2rgb SwsFunc (nearest chroma)"""
    from libav_b200 import device
    for (w, h) in ((64, 48), (66, 50), (641, 480), (1920, 1080)):
        yuv = synth.yuv420p_frame(w, h, 5)
        for fmt in (device.PIX_FMT_RGB24, device.PIX_FMT_BGR24):
            ctx = device.SwsContext(w, h, w, h, fmt, 4)
            assert not ctx.fused
            got = ctx.scale(yuv, dst_pad=6)
            want = oracle_rgb(checker, yuv, w, h, 4, pad=6)
            if fmt == device.PIX_FMT_BGR24:
                n = (w // 2) * 6
                want = np.concatenate([want[:, :n].reshape(h, -1, 3)[:, :, ::-1].reshape(h, -1), want[:, n:]], axis=1)
            assert np.array_equal(got, want), (w, h, fmt)


@pytest.mark.parametrize("variant", ["tile", "two_pass"])
def test_full_chroma_interpolation(gpu, checker, variant):
    """SWS_FULL_CHR_H_INT: yuv2rgb24_full_X_c through both general-path implementations, rgb24 and bgr24, odd sizes"""
    from libav_b200 import device
    from libav_b200._lib import lib
    FULL = 0x2000
    lib.avb200_set_tuning(b"sws_general_variant", 1 if variant == "two_pass" else 0)
    try:
        for (sw, sh, dw, dh) in [(64, 48, 64, 48), (352, 288, 640, 480), (640, 480, 352, 288), (100, 37, 333, 211), (641, 479, 641, 479), (1920, 1080, 1280, 720)]:
            yuv = tuple(synth.pad_rows(pl) for pl in synth.yuv420p_frame(sw, sh, 3))
            for flags in (4 | ACC, 2 | ACC, 0x10 | ACC, 1 | ACC):
                if sw * sh > 1500 * 1000 and flags != (4 | ACC):
                    continue
                want = oracle_rgb(checker, yuv, dw, dh, flags | FULL, pad=6)
                for fmt in (device.PIX_FMT_RGB24, device.PIX_FMT_BGR24):
                    ctx = device.SwsContext(sw, sh, dw, dh, fmt, flags | FULL)
                    assert not ctx.fused
                    got = ctx.scale(yuv, dst_pad=6)
                    w2 = want if fmt == device.PIX_FMT_RGB24 else np.concatenate([want[:, :dw * 3].reshape(dh, -1, 3)[:, :, ::-1].reshape(dh, -1), want[:, dw * 3:]], axis=1)
                    assert np.array_equal(got, w2), (variant, sw, sh, dw, dh, hex(flags), fmt, np.argwhere(got != w2)[:3].tolist())
                    ctx.close()
    finally:
        lib.avb200_set_tuning(b"sws_general_variant", 0)
