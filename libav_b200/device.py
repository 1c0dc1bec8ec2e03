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
