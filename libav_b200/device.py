"""Small harness-side helpers over the C-ABI: device buffers owned through avb200_malloc/free and
numpy <-> device copies.  Used by tests/, bench.py and smoke(); all compute goes through `_lib.lib`."""
import ctypes as C

import numpy as np

from . import _lib as L

lib = L.lib


class DevBuf:
    """A device allocation with numpy upload/download."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib.avb200_malloc(max(self.nbytes, 1))
        if not self.ptr:
            L.check(-1, "avb200_malloc(%d)" % self.nbytes)

    @classmethod
    def from_numpy(cls, a, stream=None):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        b.upload(a, stream)
        return b

    def upload(self, a, stream=None):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        L.check(lib.avb200_memcpy_h2d(self.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes, stream), "h2d")
        L.check(lib.avb200_stream_sync(stream), "sync")

    def download(self, dtype, shape, stream=None):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        L.check(lib.avb200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, stream), "d2h")
        L.check(lib.avb200_stream_sync(stream), "sync")
        return out

    def fill(self, value=0):
        L.check(lib.avb200_memset(self.ptr, value, self.nbytes, None), "memset")

    def free(self):
        if self.ptr:
            lib.avb200_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sync():
    L.check(lib.avb200_device_sync(), "device_sync")
    err = L.last_error()
    if err:
        lib.avb200_clear_error()
        raise L.AVB200Error(err)


def tile_offsets(n, tiles_per_row, stride):
    i = np.arange(n, dtype=np.uint64)
    return ((i // tiles_per_row) * 8 * stride + (i % tiles_per_row) * 8).astype(np.uint32)


def idct_put_tiles(blocks, tiles_per_row, mode=0, frame=None, use_offsets=False, clear=False):
    """Run ff_simple_idct_batch_cuda on (n, 64) int16 host blocks; returns the frame as (rows, stride) uint8
    (mode 0/1) or the transformed blocks (mode 2)."""
    n = blocks.shape[0]
    d_blocks = DevBuf.from_numpy(blocks)
    if mode == 2:
        L.check(lib.ff_simple_idct_batch_cuda(2, d_blocks.ptr, None, None, 0, n, 0, 0, None), "idct")
        sync()
        return d_blocks.download(np.int16, (n, 64))
    stride = tiles_per_row * 8
    rows = ((n + tiles_per_row - 1) // tiles_per_row) * 8
    if frame is None:
        frame = np.zeros((rows, stride), dtype=np.uint8)
    d_frame = DevBuf.from_numpy(frame)
    d_off = DevBuf.from_numpy(tile_offsets(n, tiles_per_row, stride)) if use_offsets else None
    L.check(lib.ff_simple_idct_batch_cuda(mode, d_blocks.ptr, d_frame.ptr, d_off.ptr if d_off else None, stride, n,
                                          tiles_per_row, int(clear), None), "idct")
    sync()
    out = d_frame.download(np.uint8, frame.shape)
    if clear:
        return out, d_blocks.download(np.int16, (n, 64))
    return out


# ---------------------------------------------------------------------------------------------------
# libswscale boundary
# ---------------------------------------------------------------------------------------------------
PIX_FMT_YUV420P, PIX_FMT_RGB24, PIX_FMT_BGR24, PIX_FMT_NV12, PIX_FMT_NV21 = 0, 2, 3, 23, 24
PIX_FMT_YUYV422, PIX_FMT_UYVY422 = 1, 15
PIX_FMT_ARGB, PIX_FMT_RGBA, PIX_FMT_ABGR, PIX_FMT_BGRA = 25, 26, 27, 28
RGB32_FORMATS = (25, 26, 27, 28)
# planar yuv formats (libavutil/pixfmt.h values): (log2 chroma width, log2 chroma height, bits per sample; > 8 = little-endian uint16)
PLANAR_FORMATS = {0: (1, 1, 8), 4: (1, 0, 8), 5: (0, 0, 8), 6: (2, 2, 8), 7: (2, 0, 8), 31: (0, 1, 8),
                  62: (1, 1, 9), 64: (1, 1, 10), 72: (1, 0, 9), 66: (1, 0, 10), 68: (0, 0, 9), 70: (0, 0, 10),
                  47: (1, 1, 16), 49: (1, 0, 16), 51: (0, 0, 16)}
PLANAR_FORMATS.update({f - 1: PLANAR_FORMATS[f] for f in (62, 64, 72, 66, 68, 70)})     # big-endian twins
PLANAR_FORMATS.update({f + 1: PLANAR_FORMATS[f] for f in (47, 49, 51)})
PLANAR_FORMATS.update({12: (1, 1, 8), 13: (1, 0, 8), 14: (0, 0, 8), 32: (0, 1, 8)})     # full-range (yuvj) twins of 420p 422p 444p 440p
PLANAR_BE = {61, 63, 71, 65, 67, 69, 48, 50, 52}
SWS_FAST_BILINEAR, SWS_BILINEAR, SWS_BICUBIC, SWS_X, SWS_POINT, SWS_AREA = 1, 2, 4, 8, 0x10, 0x20
SWS_BICUBLIN, SWS_GAUSS, SWS_SINC, SWS_LANCZOS, SWS_SPLINE = 0x40, 0x80, 0x100, 0x200, 0x400
SWS_ACCURATE_RND, SWS_BITEXACT = 0x40000, 0x80000


class SwsContext:
    """sws_getContext_cuda / sws_scale_cuda / sws_freeContext_cuda with numpy planes (HOST pointers), plus the
    device-pointer batch call."""

    def __init__(self, src_w, src_h, dst_w, dst_h, dst_fmt=PIX_FMT_RGB24, flags=SWS_BICUBIC | SWS_ACCURATE_RND | SWS_BITEXACT,
                 src_fmt=PIX_FMT_YUV420P):
        self.src_w, self.src_h, self.dst_w, self.dst_h, self.dst_fmt = src_w, src_h, dst_w, dst_h, dst_fmt
        self.ctx = lib.sws_getContext_cuda(src_w, src_h, src_fmt, dst_w, dst_h, dst_fmt, flags, None, None, None)
        if not self.ctx:
            L.check(-1, "sws_getContext_cuda")

    @property
    def fused(self):
        return bool(lib.sws_is_fused_cuda(self.ctx))

    def scale(self, yuv, dst_pad=0, fill=0):
        """Host-pointer drop-in call: (Y, U, V) -- (Y, UV) for nv12 / nv21, (packed,) for yuyv422 / uyvy422 / rgb24 / bgr24 --
        uint8 arrays (any row stride) -> rgb (h, 3w [+pad]) or 3 planes, pre-filled with `fill`."""
        src = (C.c_void_p * 4)(*([a.ctypes.data for a in yuv] + [None] * (4 - len(yuv))))
        sst = (C.c_int * 4)(*([a.strides[0] for a in yuv] + [0] * (4 - len(yuv))))
        if self.dst_fmt in (PIX_FMT_NV12, PIX_FMT_NV21):
            out = [np.full((self.dst_h, self.dst_w), fill, np.uint8), np.full(((self.dst_h + 1) // 2, 2 * ((self.dst_w + 1) // 2)), fill, np.uint8)]
        elif self.dst_fmt in PLANAR_FORMATS:
            hs, vs, bits = PLANAR_FORMATS[self.dst_fmt]
            dt = np.uint8 if bits == 8 else np.dtype(">u2" if self.dst_fmt in PLANAR_BE else "<u2")
            cw, ch = -((-self.dst_w) >> hs), -((-self.dst_h) >> vs)
            out = [np.full((self.dst_h, self.dst_w), fill, dt), np.full((ch, cw), fill, dt), np.full((ch, cw), fill, dt)]
        else:
            rgb16 = 36 <= self.dst_fmt <= 43 or 54 <= self.dst_fmt <= 57          # rgb565 / 555 / bgr565 / 555, rgb444 / bgr444 (LE and BE)
            bpp = 6 if self.dst_fmt in (34, 35, 59, 60) else 4 if self.dst_fmt in RGB32_FORMATS else 2 if self.dst_fmt in (PIX_FMT_YUYV422, PIX_FMT_UYVY422) or rgb16 else 1 if self.dst_fmt == 8 else 3      # 8 = gray8
            out = [np.full((self.dst_h, self.dst_w * bpp + dst_pad), fill, np.uint8)]
        dst = (C.c_void_p * 4)(*([a.ctypes.data for a in out] + [None] * (4 - len(out))))
        dstr = (C.c_int * 4)(*([a.strides[0] for a in out] + [0] * (4 - len(out))))
        r = lib.sws_scale_cuda(self.ctx, src, sst, 0, self.src_h, dst, dstr)
        if r != self.dst_h:
            L.check(-1, "sws_scale_cuda")
        return out[0] if len(out) == 1 else out

    def scale_device(self, d_src, src_strides, d_dst, dst_strides, nframes=1, src_frame=None, dst_frame=None, stream=None):
        src = (C.c_void_p * 3)(*(list(d_src) + [None] * (3 - len(d_src))))
        sst = (C.c_int * 3)(*(list(src_strides) + [0] * (3 - len(src_strides))))
        dst = (C.c_void_p * 3)(*(list(d_dst) + [None] * (3 - len(d_dst))))
        dstr = (C.c_int * 3)(*(list(dst_strides) + [0] * (3 - len(dst_strides))))
        sf = (C.c_size_t * 3)(*(list(src_frame) + [0] * (3 - len(src_frame)))) if src_frame else None
        df = (C.c_size_t * 3)(*(list(dst_frame) + [0] * (3 - len(dst_frame)))) if dst_frame else None
        r = lib.sws_scale_frames_cuda(self.ctx, src, sst, sf, dst, dstr, df, nframes, stream)
        if r < 0:
            L.check(-1, "sws_scale_frames_cuda")
        return r

    def close(self):
        if self.ctx:
            lib.sws_freeContext_cuda(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
