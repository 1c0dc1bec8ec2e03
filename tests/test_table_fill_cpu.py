"""Which entries do the hooks fill?  Every ff_*_init_cuda() of the PRODUCT library is run on a zeroed table (pointer logic only: no GPU needed)
and the set of non-NULL entries is compared, word for word, with what the reference's own ff_*_init() fills for the same arguments
(oracle/_ref ref_table_fill).  Where the product takes a table over it must fill exactly the reference's entries -- no hole a codec would
call through, no entry the reference leaves NULL -- except the ones documented as left to the caller."""
import ctypes as C

import numpy as np
import pytest

WORD = C.sizeof(C.c_void_p)


def words(table, nbytes=None):
    n = (nbytes or C.sizeof(table)) // WORD
    return [1 if v else 0 for v in (C.c_void_p * n).from_buffer(table)]


def ref_words(refo, kind, a=0, b=0):
    out = np.zeros(4096, np.uint8)
    n = refo.table_fill(kind, a, b, out.ctypes.data, len(out)) if kind != 3 else refo.pred_table_fill(a, b, out.ctypes.data, len(out))
    assert n > 0
    return out[:n].tolist()


def test_hooks_fill_what_the_reference_fills(built, refo):
    import libav_b200._lib as L
    from libav_b200 import tables
    lib = L.lib
    n = 0
    for bits in (8, 9, 10):
        for idc in (1, 2):
            t = tables.H264DSPContext(); lib.ff_h264dsp_init_cuda(C.byref(t), bits, idc)
            assert words(t) == ref_words(refo, 0, bits, idc), ("H264DSPContext", bits, idc); n += 1
            t = tables.H264PredContext(); lib.ff_h264_pred_init_cuda(C.byref(t), 27, bits, idc)
            assert words(t) == ref_words(refo, 3, bits, idc), ("H264PredContext", bits, idc); n += 1
        t = tables.H264QpelContext(); lib.ff_h264qpel_init_cuda(C.byref(t), bits)
        assert words(t) == ref_words(refo, 1, bits), ("H264QpelContext", bits); n += 1
        t = tables.H264ChromaContext(); lib.ff_h264chroma_init_cuda(C.byref(t), bits)
        assert words(t) == ref_words(refo, 2, bits), ("H264ChromaContext", bits); n += 1
    t = tables.HpelDSPContext(); lib.ff_hpeldsp_init_cuda(C.byref(t), 0)
    assert words(t) == ref_words(refo, 4, 0); n += 1
    t = tables.QpelDSPContext(); lib.ff_qpeldsp_init_cuda(C.byref(t))
    assert words(t) == ref_words(refo, 6); n += 1
    t = tables.BlockDSPContext(); lib.ff_blockdsp_init_cuda(C.byref(t))
    assert words(t) == ref_words(refo, 7); n += 1
    for bits, algo in ((8, 0), (8, 1), (8, 2), (10, 0)):                  # FF_DCT_AUTO, FASTINT, INT at 8 bit; 10 bit
        t = tables.FDCTDSPContext(); lib.ff_fdctdsp_init_cuda(C.byref(t), algo, bits, int(bits > 8))
        assert words(t) == ref_words(refo, 8, bits, algo), ("FDCTDSPContext", bits, algo); n += 1
    t = tables.PixblockDSPContext(); lib.ff_pixblockdsp_init_cuda(C.byref(t), 0)
    assert words(t) == ref_words(refo, 9, 8); n += 1
    for bits, algo in ((8, 0), (8, 2), (10, 0)):                         # FF_IDCT_AUTO, FF_IDCT_SIMPLE at 8 bit; 10 bit
        t = tables.IDCTDSPContext(); lib.ff_idctdsp_init_cuda(C.byref(t), algo, bits, int(bits > 8))
        assert words(t, 6 * WORD) == ref_words(refo, 10, bits, algo), ("IDCTDSPContext", bits, algo); n += 1
    # MECmpContext: everything the reference fills except the encoder-state kinds (dct_sad, quant_psnr, bit, rd, dct_max, dct264_sad: SURVEY 8 a12)
    t = tables.MECmpContext(); lib.ff_me_cmp_init_cuda(C.byref(t))
    mine, ref = words(t), ref_words(refo, 5, 8)
    left = {name for name in ("dct_sad", "quant_psnr", "bit", "rd", "dct_max", "dct264_sad")}
    for name, _ in tables.MECmpContext._fields_:
        f = getattr(tables.MECmpContext, name)
        sl = slice(f.offset // WORD, (f.offset + f.size) // WORD)
        if name in left:
            assert not any(mine[sl]) and any(ref[sl]), name
        else:
            assert mine[sl] == ref[sl], name
    n += 1
    assert n == 30


def test_what_is_not_taken_over_is_left_alone(built):
    import libav_b200._lib as L
    from libav_b200 import tables
    lib = L.lib
    for make, call in ((tables.H264DSPContext, lambda t: lib.ff_h264dsp_init_cuda(t, 12, 1)), (tables.H264QpelContext, lambda t: lib.ff_h264qpel_init_cuda(t, 12)),
                       (tables.H264ChromaContext, lambda t: lib.ff_h264chroma_init_cuda(t, 14)), (tables.H264PredContext, lambda t: lib.ff_h264_pred_init_cuda(t, 139, 8, 1)),
                       (tables.H264PredContext, lambda t: lib.ff_h264_pred_init_cuda(t, 27, 12, 1)), (tables.IDCTDSPContext, lambda t: lib.ff_idctdsp_init_cuda(t, 1, 8, 0)),
                       (tables.FDCTDSPContext, lambda t: lib.ff_fdctdsp_init_cuda(t, 3, 8, 0)), (tables.PixblockDSPContext, lambda t: lib.ff_pixblockdsp_init_cuda(t, 1))):
        t = make()
        call(C.byref(t))
        w = words(t)
        if make is tables.IDCTDSPContext:
            assert not any(w[3:6])                          # FF_IDCT_INT: the transform entries stay, the pixel clamps are taken
        else:
            assert not any(w), make.__name__
