"""ff_set_cmp() -- the reference's own selection of compare functions (libavcodec/me_cmp.c:365-417, unmodified in oracle/_ref) -- over the
MECmpContext that the PRODUCT library's ff_me_cmp_init_cuda() filled: it must hand out the product's slots for the kinds that are taken over
and nothing for the encoder-state kinds.  Pointer logic only, so the CPU suite runs it on the real product library; the GPU test calls
what was selected (tests/test_zz_gpu_late_slots.py)."""
import ctypes as C

FF_CMP = dict(SAD=0, SSE=1, SATD=2, DCT=3, PSNR=4, BIT=5, RD=6, ZERO=7, VSAD=8, VSSE=9, NSSE=10, DCTMAX=13, DCT264=14)


def select(refo, table, kind):
    out = (C.c_void_p * 6)()
    refo.me_cmp_select(C.byref(table), FF_CMP[kind], out)
    return [out[i] for i in range(6)]


def addr(f):
    return C.cast(f, C.c_void_p).value


def test_reference_selection_hands_out_the_product_slots(built, refo):
    import libav_b200._lib as L
    from libav_b200 import tables
    t = tables.MECmpContext()
    L.lib.ff_me_cmp_init_cuda(C.byref(t))
    for kind, field in (("SAD", t.sad), ("SSE", t.sse), ("SATD", t.hadamard8_diff), ("VSAD", t.vsad), ("VSSE", t.vsse), ("NSSE", t.nsse)):
        got = select(refo, t, kind)
        assert got == [addr(field[i]) for i in range(6)] and got[0], kind
    assert select(refo, t, "SAD")[1] and select(refo, t, "SATD")[4] and select(refo, t, "SATD")[5]      # 8-wide and the intra pair
    for kind in ("DCT", "PSNR", "BIT", "RD", "DCTMAX", "DCT264"):          # encoder-state kinds: whatever the caller had installed (nothing here)
        assert select(refo, t, kind) == [None] * 6, kind
    assert all(select(refo, t, "ZERO"))                                     # the reference's own zero_cmp
