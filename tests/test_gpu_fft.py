"""GPU: float FFT / MDCT filterbank against the CPU oracle at the north star's tolerance (1e-6 relative to the largest
output magnitude), sizes 2^4 .. 2^12 like the reference's FATE targets (tests/fate/fft.mak), batched; plus the FFTContext
slots installed by ff_fft_init_cuda / ff_mdct_init_cuda on a live reference context (revtab order, context tables)."""
import ctypes as C

import numpy as np
import pytest

from oracle.loader import ptr

pytestmark = pytest.mark.gpu
TOL = 1e-6


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("nbits", range(1, 13))
@pytest.mark.parametrize("inverse", [0, 1])
def test_fft_batch(gpu, checker, nbits, inverse):
    from libav_b200 import device
    rng = np.random.default_rng(nbits * 2 + inverse)
    T = 37
    z = rng.uniform(-1, 1, size=(T, 2 << nbits)).astype(np.float32)
    want = z.copy()
    for t in range(T):
        if nbits >= 2:
            checker.fft(nbits, inverse, ptr(want[t]))
        else:
            a, b = want[t, 0] + 1j * want[t, 1], want[t, 2] + 1j * want[t, 3]
            want[t] = [(a + b).real, (a + b).imag, (a - b).real, (a - b).imag]
    d = device.DevBuf.from_numpy(z)
    gpu.check(gpu.lib.ff_fft_batch_cuda(nbits, inverse, d.ptr, T, None))
    device.sync()
    got = d.download(np.float32, z.shape)
    assert rel_err(got, want) < TOL * max(1, nbits / 4), rel_err(got, want)


@pytest.mark.parametrize("nbits", range(4, 13))
@pytest.mark.parametrize("scale", [1.0, -1.0, 1.0 / 32768])
def test_mdct_batch(gpu, checker, nbits, scale):
    from libav_b200 import device
    rng = np.random.default_rng(nbits)
    n, T = 1 << nbits, 19
    for op, fn, nin, nout in ((0, "imdct_half", n // 2, n // 2), (1, "imdct_calc", n // 2, n), (2, "mdct_calc", n, n // 2)):
        x = rng.uniform(-1, 1, size=(T, nin)).astype(np.float32)
        want = np.zeros((T, nout), np.float32)
        for t in range(T):
            getattr(checker, fn)(nbits, scale, ptr(want[t]), ptr(x[t].copy()))
        d_in, d_out = device.DevBuf.from_numpy(x), device.DevBuf(4 * T * nout)
        gpu.check(gpu.lib.ff_mdct_batch_cuda(op, nbits, scale, d_out.ptr, d_in.ptr, T, None))
        device.sync()
        got = d_out.download(np.float32, want.shape)
        assert rel_err(got, want) < TOL * max(1, nbits / 4), (fn, rel_err(got, want))


def test_fftcontext_slots(gpu, refo):
    """the reference's own ff_fft_init / ff_mdct_init context, then the CUDA hook, then the calls a codec makes"""
    from libav_b200 import tables
    lib = refo.lib
    lib.ref_fft_ctx_new.restype = C.c_void_p
    lib.ref_fft_ctx_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    lib.ref_fft_ctx_free.argtypes = [C.c_void_p, C.c_int]
    assert lib.ref_sizeof_fftcontext() == C.sizeof(tables.FFTContext)
    rng = np.random.default_rng(1)
    for nbits in (4, 7, 10):
        for inverse in (0, 1):
            p = lib.ref_fft_ctx_new(nbits, inverse, 0, 1.0)
            ctx = C.cast(p, C.POINTER(tables.FFTContext))
            z = rng.uniform(-1, 1, size=2 << nbits).astype(np.float32)
            want, got = z.copy(), z.copy()
            ctx.contents.fft_permute(ctx, ptr(want)); ctx.contents.fft_calc(ctx, ptr(want))          # reference C slots
            gpu.lib.ff_fft_init_cuda(ctx)
            ctx.contents.fft_permute(ctx, ptr(got)); ctx.contents.fft_calc(ctx, ptr(got))             # permute = C, calc = CUDA
            assert rel_err(got, want) < TOL * max(1, nbits / 4)
            lib.ref_fft_ctx_free(p, 0)
    for nbits in (5, 9, 11):
        n = 1 << nbits
        for inverse, fields in ((1, ("imdct_half", "imdct_calc")), (0, ("mdct_calc",))):
            p = lib.ref_fft_ctx_new(nbits, inverse, 1, -1.0 / 7)
            ctx = C.cast(p, C.POINTER(tables.FFTContext))
            x = rng.uniform(-1, 1, size=n).astype(np.float32)
            want = {}
            for f in fields:
                out = np.zeros(n, np.float32)
                getattr(ctx.contents, f)(ctx, out.ctypes.data_as(C.POINTER(C.c_float)), x.ctypes.data_as(C.POINTER(C.c_float)))
                want[f] = out
            gpu.lib.ff_mdct_init_cuda(ctx)
            for f in fields:
                out = np.zeros(n, np.float32)
                getattr(ctx.contents, f)(ctx, out.ctypes.data_as(C.POINTER(C.c_float)), x.ctypes.data_as(C.POINTER(C.c_float)))
                assert rel_err(out, want[f]) < TOL * max(1, nbits / 4), (f, nbits)
            lib.ref_fft_ctx_free(p, 1)
    assert gpu.last_error() == ""
