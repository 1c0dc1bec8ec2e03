"""GPU: slots added while no GPU box was at hand (developed against tests/hostsim/, see tests/test_hostsim_slots_cpu.py): the
chroma_format_idc 2 entries of H264DSPContext and libswscale's per-line SwsContext slots.  Same cases as the host simulation, now
through the product library on a B200.  (The file name sorts last on purpose: `pytest -x` reaches these after every older test.)"""
import ctypes as C

import numpy as np
import pytest

import slot_cases
import sws_line_cases as L

pytestmark = pytest.mark.gpu


def test_h264dsp_slots_422(gpu, checker):
    slot_cases.h264dsp_422_cases(gpu.lib, gpu.last_error, checker)


def make_ctx_factory(gpu):
    def make_ctx(dst_fmt, flags, src_fmt=0):
        ctx = gpu.lib.sws_getContext_cuda(64, 48, src_fmt, 96, 80, dst_fmt, flags, None, None, None)
        assert ctx, (dst_fmt, flags, gpu.last_error())
        return ctx, gpu.lib.sws_freeContext_cuda
    return make_ctx


def test_sws_line_slots(gpu, checker):
    calls = L.SlotCalls(gpu.lib, make_ctx_factory(gpu), gpu.last_error)
    try:
        assert L.compare(calls, L.OracleCalls(checker), seed=5) > 150
    finally:
        calls.close()
    assert gpu.last_error() == ""


def test_sws_line_slots_colourspace(gpu, checker):
    """sws_setColorspaceDetails_cuda on the registered context is seen by the packed slots"""
    tab = (C.c_int * 4)(*L.FCC)

    def make_ctx(dst_fmt, flags, src_fmt=0):
        ctx, free = make_ctx_factory(gpu)(dst_fmt, flags, src_fmt)
        if dst_fmt in L.PACKED_FMTS and dst_fmt not in (1, 15):
            assert gpu.lib.sws_setColorspaceDetails_cuda(ctx, tab, 0, tab, 0, 3000, 70000, 60000) == 0
        return ctx, free
    calls = L.SlotCalls(gpu.lib, make_ctx, gpu.last_error)
    checker.sws_set_colorspace(tab, 0, 3000, 70000, 60000)
    try:
        assert L.compare(calls, L.OracleCalls(checker), seed=6) > 150
    finally:
        checker.sws_set_colorspace(None, 0, 0, 0, 0)
        calls.close()
    assert gpu.last_error() == ""


def test_sws_line_slot_refusals(gpu):
    from libav_b200 import tables
    t = tables.SwsLineSlotsCUDA()
    assert gpu.lib.ff_sws_init_swscale_cuda(None, None, C.byref(t)) == -1
    gpu.lib.avb200_clear_error()
    ctx = gpu.lib.sws_getContext_cuda(64, 48, 0, 96, 80, 2, 4, None, None, None)
    assert gpu.lib.ff_sws_init_swscale_cuda(C.c_void_p(0x7000), ctx, C.byref(t)) == 0
    gpu.lib.sws_freeContext_cuda(ctx)                       # the registration ends with the context
    import numpy as np
    lum, cu, out = np.zeros(64, np.int16), np.zeros(64, np.int16), np.zeros(256, np.uint8)
    t.yuv2packed1(C.c_void_p(0x7000), L.vp(lum), L.ptrs([cu, cu]), L.ptrs([cu, cu]), None, L.vp(out), 16, 0, 0)
    assert "not registered" in gpu.last_error()
    gpu.lib.avb200_clear_error()


def test_sws_range_conversion_frames(gpu, checker):
    """full-range (yuvj) <-> limited-range planar yuv through sws_scale_cuda: the two-pass path with the range kernel between the passes"""
    import numpy as np
    from libav_b200 import device
    import test_sws_range_cpu as R
    from test_sws_planar_dst import run
    n = 0
    for (sf, df, w, h, dw, dh, flags) in list(R.cases()) + [(12, 0, 1280, 720, 1920, 1080, 4 | R.ACC), (0, 12, 1920, 1080, 1280, 720, 4 | R.ACC)]:
        pl = R.planes(sf, w, h, 21)
        rc, want = run(checker, sf, pl, w, h, df, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, df, flags, src_fmt=sf)
        got = ctx.scale(pl, fill=7)
        for a, b in zip(got, want):
            assert np.array_equal(a, b[:, :a.shape[1]]), (sf, df, w, h, dw, dh, hex(flags))
        ctx.close()
        n += 1
    assert n > 400 and gpu.last_error() == ""


@pytest.mark.parametrize("bits", [9, 10])
def test_h264_high_bit_depth_slots(gpu, checker, bits):
    import hbd_cases
    assert hbd_cases.compare(hbd_cases.TableCallee(gpu.lib), checker, bits, seed=3) > 800
    assert gpu.last_error() == ""


def test_simple_idct10(gpu, checker):
    """ff_idctdsp_init_cuda(c, idct_algo, 10, 1): the 10-bit simple IDCT entries, and ff_simple_idct10_batch_cuda on device buffers"""
    import numpy as np
    import idct10_cases
    from libav_b200 import tables
    lib = gpu.lib
    t = tables.IDCTDSPContext()
    lib.ff_idctdsp_init_cuda(C.byref(t), 0, 10, 1)
    assert idct10_cases.slot_cases(t, checker) == 360
    assert t.put_pixels_clamped and t.add_pixels_clamped

    def run_batch(mode, blk, frame, off):
        bufs = []
        for a in (blk, frame, off):
            d = lib.avb200_malloc(a.nbytes)
            assert d and lib.avb200_memcpy_h2d(d, a.ctypes.data, a.nbytes, None) == 0
            bufs.append(d)
        assert lib.ff_simple_idct10_batch_cuda(mode, bufs[0], bufs[1], bufs[2], frame.strides[0], len(blk), None) == 0
        assert lib.avb200_memcpy_d2h(blk.ctypes.data, bufs[0], blk.nbytes, None) == 0 and lib.avb200_memcpy_d2h(frame.ctypes.data, bufs[1], frame.nbytes, None) == 0
        assert lib.avb200_device_sync() == 0
        for d in bufs:
            lib.avb200_free(d)
        return blk, frame
    for mode in range(3):
        idct10_cases.batch_case(run_batch, checker, mode)                       # pitch 1032: the thread-per-block kernel
        idct10_cases.batch_case(run_batch, checker, mode, n=5003, pad=8)        # 16-byte rows: the staged kernel, ragged last group
        idct10_cases.batch_case(run_batch, checker, mode, n=4000, pad=8, shift=1, seed=7)   # ... with blocks off the 16-byte grid
        idct10_cases.batch_case(run_batch, checker, mode, n=100000, pad=8, seed=3)          # more groups than resident warps
    assert gpu.last_error() == ""


def test_fdct10(gpu, checker):
    import idct10_cases
    from libav_b200 import tables
    lib = gpu.lib
    t = tables.FDCTDSPContext()
    lib.ff_fdctdsp_init_cuda(C.byref(t), 0, 10, 1)

    def run_batch(which, blk):
        d = lib.avb200_malloc(blk.nbytes)
        assert d and lib.avb200_memcpy_h2d(d, blk.ctypes.data, blk.nbytes, None) == 0
        assert lib.ff_fdct_batch_cuda(which, d, len(blk), None) == 0
        assert lib.avb200_memcpy_d2h(blk.ctypes.data, d, blk.nbytes, None) == 0 and lib.avb200_device_sync() == 0
        lib.avb200_free(d)
        return blk
    idct10_cases.fdct10_cases(t, run_batch, checker)
    assert gpu.last_error() == ""


@pytest.mark.parametrize("bits", [8, 9, 10])
def test_h264_pred_422(gpu, checker, bits):
    """ff_h264_pred_init_cuda(h, AV_CODEC_ID_H264, bits, 2): the chroma entries are the 8 x 16 functions; the luma entries stay what they are at idc 1"""
    import hbd_cases
    from libav_b200 import tables
    fill = lambda h, b: gpu.lib.ff_h264_pred_init_cuda(h, 27, b, 2)
    assert hbd_cases.pred422_compare(hbd_cases.Pred422Callee(fill), checker, bits, seed=3) > 150
    h1, h2 = tables.H264PredContext(), tables.H264PredContext()
    gpu.lib.ff_h264_pred_init_cuda(C.byref(h1), 27, bits, 1); gpu.lib.ff_h264_pred_init_cuda(C.byref(h2), 27, bits, 2)
    same = lambda a, b: C.cast(a, C.c_void_p).value == C.cast(b, C.c_void_p).value
    assert same(h1.pred4x4[3], h2.pred4x4[3]) and same(h1.pred16x16[3], h2.pred16x16[3]) and not same(h1.pred8x8[0], h2.pred8x8[0])
    assert gpu.last_error() == ""


def test_startcode_slot(gpu):
    import numpy as np
    from libav_b200 import tables
    for bits in (8, 10):
        c = tables.H264DSPContext()
        gpu.lib.ff_h264dsp_init_cuda(C.byref(c), bits, 1)
        slot_cases.startcode_cases(np.random.default_rng(bits), c.startcode_find_candidate)
    assert gpu.last_error() == ""


def test_sws_scale_negative_strides(gpu):
    """bottom-up pictures (negative strides on either side) give the bytes of the same picture stored top-down; what the converters leave
    untouched (row padding) stays untouched"""
    import numpy as np
    from test_sws_planar_dst import source
    lib = gpu.lib
    for (sf, df, w, h, dw, dh, flags) in [(0, 2, 64, 48, 96, 80, 4), (0, 2, 66, 50, 66, 50, 4 | 0x40000), (0, 0, 101, 37, 64, 48, 4), (4, 5, 64, 48, 64, 48, 2),
                                           (1, 2, 64, 48, 96, 80, 4), (0, 28, 64, 48, 33, 25, 4), (0, 23, 64, 48, 96, 80, 4)]:
        ctx = lib.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, None, None, None)
        assert ctx, (sf, df, gpu.last_error())
        pl = source(sf, w, h, 31)
        if df in (2, 28):
            outs = [np.full((dh, dw * (4 if df == 28 else 3) + 24), 9, np.uint8)]
        elif df == 23:
            outs = [np.full((dh, dw + 8), 9, np.uint8), np.full(((dh + 1) // 2, 2 * ((dw + 1) // 2) + 8), 9, np.uint8)]
        else:
            cw, ch = (dw, dh) if df == 5 else ((dw + 1) // 2, (dh + 1) // 2)
            outs = [np.full((dh, dw + 8), 9, np.uint8), np.full((ch, cw + 8), 9, np.uint8), np.full((ch, cw + 8), 9, np.uint8)]

        def call(src, dst):
            sp = (C.c_void_p * 4)(*([a.ctypes.data for a in src] + [None] * (4 - len(src))))
            ss = (C.c_int * 4)(*([a.strides[0] for a in src] + [0] * (4 - len(src))))
            dp = (C.c_void_p * 4)(*([a.ctypes.data for a in dst] + [None] * (4 - len(dst))))
            ds = (C.c_int * 4)(*([a.strides[0] for a in dst] + [0] * (4 - len(dst))))
            assert lib.sws_scale_cuda(ctx, sp, ss, 0, h, dp, ds) == dh, gpu.last_error()
        want = [o.copy() for o in outs]
        call(pl, want)
        # the same pictures stored bottom-up: flipped copies viewed through negative strides
        src_up = [np.ascontiguousarray(a[::-1])[::-1] for a in pl]
        got_store = [np.ascontiguousarray(o[::-1]) for o in outs]
        got = [g[::-1] for g in got_store]
        assert all(a.strides[0] < 0 for a in src_up + got)
        call(src_up, got)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (sf, df, w, h, dw, dh)
        # mixed: top-down source, bottom-up destination
        got_store = [np.ascontiguousarray(o[::-1]) for o in outs]
        got = [g[::-1] for g in got_store]
        call(pl, got)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), ("mixed", sf, df)
        lib.sws_freeContext_cuda(ctx)
    assert gpu.last_error() == ""


def test_sws_slots_under_the_reference_scheduler(gpu, refo):
    """the product's per-line slots installed in real reference SwsContexts and driven by the reference's unmodified swscale()
    (tests/sws_dropin_cases.py); needs the compiled reference, which travels to the GPU box as oracle/_ref"""
    import sws_dropin_cases as D

    def make_ctx(sf, w, h, df, dw, dh, flags):
        ctx = gpu.lib.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, None, None, None)
        assert ctx, (sf, df, hex(flags), gpu.last_error())
        return ctx
    assert D.check(refo, gpu.lib, make_ctx, gpu.lib.sws_freeContext_cuda, geoms=D.GEOMS[:3]) == 46
    assert gpu.last_error() == ""


def test_loop_filter_slots_under_the_reference_driver(gpu, refo):
    """the reference's own deblocking driver (h264_loopfilter.c, unmodified, in oracle/_ref) filtering pictures through the product's
    H264DSPContext slots: tests/h264_dropin_cases.py"""
    import h264_dropin_cases as D
    assert D.check(refo, gpu.lib, sizes=((6, 4), (1, 1))) == 12
    assert gpu.last_error() == ""


def test_me_cmp_functions_selected_by_the_reference(gpu, refo, checker):
    """ff_set_cmp() (the reference's, in oracle/_ref) picks compare functions out of the product's MECmpContext; call what it picked"""
    import numpy as np
    from libav_b200 import tables
    from test_me_cmp_select_cpu import select
    t = tables.MECmpContext()
    gpu.lib.ff_me_cmp_init_cuda(C.byref(t))
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, size=(24, 48), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    for kind, okind in (("SAD", 1), ("SSE", 2), ("SATD", 3), ("NSSE", 6)):
        picked = select(refo, t, kind)
        for sidx in (0, 1):
            f = tables.me_cmp_func(picked[sidx])
            h = 16 if sidx == 0 else 8
            got = f(None, slot_cases.P(a, 48 * 2 + 16), slot_cases.P(b, 48 * 3 + 5), 48, h)
            assert got == checker.me_cmp(okind, sidx, 0, C.c_void_p(a.ctypes.data + 48 * 2 + 16), C.c_void_p(b.ctypes.data + 48 * 3 + 5), 48, h), (kind, sidx)
    assert gpu.last_error() == ""


def test_sws_filter_frames(gpu, checker):
    """sws_getContext_cuda with a SwsFilter pair (the reference's struct layout): frames against the checker for the vector sets of
    tests/test_sws_filter_cpu.py (their banks are pinned to the reference on the CPU)"""
    import numpy as np
    import test_sws_filter_cpu as F
    from libav_b200 import synth
    lib = gpu.lib
    for k, s in enumerate(F.SETS):
        sf, keep1 = F.make_filter(s.get("src"))
        df, keep2 = F.make_filter(s.get("dst"))
        keep = F.set_oracle(checker, s)
        try:
            for (sw, sh, dw, dh) in F.GEOMS:
                yuv = synth.yuv420p_frame(sw, sh, 3 + k)
                for flags in (4 | F.ACC, 4, 2):
                    for dfmt in (2, 0):
                        if dfmt == 2:
                            rc, want = F.to_rgb(checker, yuv, dw, dh, flags, pad=3)
                            want = [want]
                            got = [np.zeros((dh, dw * 3 + 3), np.uint8)]
                        else:
                            rc, want = F.to_yuv(checker, yuv, dw, dh, flags)
                            got = [np.zeros_like(w) for w in want]
                        assert rc == dh
                        ctx = lib.sws_getContext_cuda(sw, sh, 0, dw, dh, dfmt, flags, C.byref(sf), C.byref(df), None)
                        assert ctx, (k, gpu.last_error())
                        sp = (C.c_void_p * 4)(*([a.ctypes.data for a in yuv] + [None]))
                        ss = (C.c_int * 4)(*([a.strides[0] for a in yuv] + [0]))
                        dp = (C.c_void_p * 4)(*([a.ctypes.data for a in got] + [None] * (4 - len(got))))
                        ds = (C.c_int * 4)(*([a.strides[0] for a in got] + [0] * (4 - len(got))))
                        assert lib.sws_scale_cuda(ctx, sp, ss, 0, sh, dp, ds) == dh, gpu.last_error()
                        lib.sws_freeContext_cuda(ctx)
                        for a, b in zip(got, want):
                            assert np.array_equal(a, b), (k, sw, sh, dw, dh, hex(flags), dfmt)
        finally:
            F.set_oracle(checker, None)
        del keep, keep1, keep2
    assert gpu.last_error() == ""


def test_sws_range_conversion_with_a_semi_planar_or_packed_side(gpu, checker):
    import numpy as np
    from libav_b200 import device
    import test_sws_range_cpu as R
    n = 0
    for (sf, df, w, h, dw, dh, flags) in R.mixed_cases():
        pl = R.mixed_planes(sf, w, h, 13)
        rc, want = R.mixed_run(checker, sf, pl, w, h, df, dw, dh, flags)
        assert rc == dh
        ctx = device.SwsContext(w, h, dw, dh, df, flags, src_fmt=sf)
        got = ctx.scale(pl, fill=7)
        got = got if isinstance(got, list) else [got]
        for a, b in zip(got, want):
            assert np.array_equal(a, b[:, :a.shape[1]]), (sf, df, w, h, dw, dh, hex(flags))
        ctx.close()
        n += 1
    assert n > 230 and gpu.last_error() == ""


def test_sws_high_bit_depth_sources(gpu, checker):
    """9 / 10 / 16-bit planar sources through sws_scale_cuda (tests/test_sws_hbd_sources_cpu.py cases; host-simulated on the CPU)"""
    import numpy as np
    import test_sws_hbd_sources_cpu as H
    lib = gpu.lib
    n = 0
    for (sf, df, w, h, dw, dh, flags) in H.cases((64, 63, 62, 47, 66, 70, 69)):
        pl = H.planes(sf, w, h, 5)
        rc, want = H.run(checker, sf, pl, w, h, df, dw, dh, flags)
        assert rc == dh
        ctx = lib.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, None, None, None)
        assert ctx, (sf, df, gpu.last_error())
        got = H.outputs(df, dw, dh)
        sp = (C.c_void_p * 4)(*([a.ctypes.data for a in pl] + [None]))
        ss = (C.c_int * 4)(*([a.strides[0] for a in pl] + [0]))
        dp = (C.c_void_p * 4)(*([a.ctypes.data for a in got] + [None] * (4 - len(got))))
        ds = (C.c_int * 4)(*([a.strides[0] for a in got] + [0] * (4 - len(got))))
        assert lib.sws_scale_cuda(ctx, sp, ss, 0, h, dp, ds) == dh, gpu.last_error()
        lib.sws_freeContext_cuda(ctx)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (sf, df, w, h, dw, dh, hex(flags))
        n += 1
    assert n > 1000 and gpu.last_error() == ""


def test_sws_scale_slices(gpu, refo):
    """sws_scale() slice by slice (tests/test_hostsim_sws_slices_cpu.py has the CPU twin): per call the return value and the whole destination
    picture against the compiled reference given the same slices -- through the tile kernels here"""
    import test_hostsim_sws_slices_cpu as S
    ACC = 0x40000 | 0x80000
    n = 0
    for sf, df in ((0, 2), (0, 0), (4, 0), (5, 28), (0, 23), (23, 0), (2, 0), (0, 1), (0, 63)):
        for (w, h, dw, dh) in ((64, 48, 96, 80), (66, 50, 33, 25), (64, 48, 64, 48)):
            for flags in (4 | ACC, 2):
                if (w, h) == (dw, dh) and not (flags & ACC == ACC and sf in (0, 4, 5) and df in (2, 28)):
                    continue
                for plan in S.plans(h, 1 << S.VSUB.get(sf, 0)):
                    S.compare(gpu.lib, refo, sf, df, w, h, dw, dh, flags, plan)
                    n += 1
    for sf, df, flags, align in ((0, 0, 4, 2), (0, 2, 4, 2), (0, 23, 4, 2), (23, 0, 4, 2), (2, 3, 4, 1), (3, 0, 4, 2), (1, 0, 4, 2), (4, 1, 4, 1), (0, 62, 4, 2)):
        for plan in S.plans(48, align):
            S.compare(gpu.lib, refo, sf, df, 64, 48, 64, 48, flags, plan)
            n += 1
    assert n > 100 and gpu.last_error() == ""


def test_me_cmp_quant_metrics_batch(gpu, checker, orc):
    """MECmpContext.quant_psnr / bit / rd (libav_b200/csrc/me_cmp_enc.cu): ff_me_cmp_enc_batch_cuda on device buffers, every encoder state x
    metric x block size against the compiled reference's me_cmp.c:621-782 (with its own ff_dct_quantize_c / dct_unquantize_*_c)"""
    import numpy as np
    import enc_cases
    lib = gpu.lib

    def run_batch(kind, sidx, handle, cur, ref, recs, h):
        out, last = np.zeros(len(recs), np.int32), np.zeros(len(recs), np.int32)
        bufs = []
        for a in (cur, ref, recs, out, last):
            d = lib.avb200_malloc(a.nbytes)
            assert d and lib.avb200_memcpy_h2d(d, a.ctypes.data, a.nbytes, None) == 0
            bufs.append(d)
        assert lib.ff_me_cmp_enc_batch_cuda(kind, sidx, handle, bufs[0], bufs[1], cur.strides[0], h, bufs[2], len(recs), bufs[3], bufs[4], None) == 0, gpu.last_error()
        assert lib.avb200_memcpy_d2h(out.ctypes.data, bufs[3], out.nbytes, None) == 0 and lib.avb200_memcpy_d2h(last.ctypes.data, bufs[4], last.nbytes, None) == 0
        assert lib.avb200_device_sync() == 0
        for d in bufs:
            lib.avb200_free(d)
        return out, last
    assert enc_cases.batch_cases(lib, run_batch, checker, orc, n=200) > 80000
    assert lib.ff_me_cmp_enc_batch_cuda(13, 0, None, None, None, 0, 8, None, 0, None, None, None) == -1
    lib.avb200_clear_error()


def test_me_cmp_enc_tables_at_a_reused_address(gpu, checker, orc):
    """the device copies of the VLC length tables are found by host address: tables announced again at an address that held others before
    (freed and reused memory; here rewritten in place) must be uploaded again -- by ff_me_cmp_enc_init_cuda and ff_me_cmp_enc_state_cuda"""
    import ctypes as C
    import enc_cases as E
    from libav_b200 import tables
    T = E.Tables(4)
    lib = gpu.lib
    u8p = C.POINTER(C.c_uint8)
    for round_ in range(2):
        if round_:
            T2 = E.Tables(99)
            for name in ("intra_len", "intra_last", "inter_len", "inter_last", "luma_dc"):
                getattr(T, name)[:] = getattr(T2, name)                      # same addresses, other tables
        enc = E.FakeEncoder(T)
        table = tables.MECmpContext()
        key = C.c_void_p(enc.key.ctypes.data)
        label, st = E.states(T)[5]
        enc.load(st, checker)
        assert lib.ff_me_cmp_enc_init_cuda(C.byref(table), key, C.byref(enc.view)) == 0
        cur, rf = E.block_pairs(6, seed=35)
        want, wside = E.oracle_scores(checker, 15, 1, st, cur, rf, 6, 8, h263_guard=orc)
        n = 0
        for i in range(6):
            if want[i] is None:
                continue
            enc.ints[6], enc.ints[8] = st.mb_intra, -2
            a = C.cast(cur.ctypes.data + 16 * i * cur.strides[0] + 8, u8p)
            b = C.cast(rf.ctypes.data + 16 * i * rf.strides[0] + 8, u8p)
            assert table.bit[1](key, a, b, cur.strides[0], 8) == want[i], (round_, i)
            n += 1
        assert n >= 3
        lib.ff_me_cmp_enc_uninit_cuda(key)
        # the batch state announces its tables too
        p, _, _, _ = E.product_state(st, checker)
        vlc = E.vlc_tables(T)
        handle = lib.ff_me_cmp_enc_state_cuda(C.byref(p), C.byref(vlc))
        assert handle
        recs = np.array([[16 * i * cur.strides[0] + 8, 16 * i * rf.strides[0] + 8] for i in range(6)], np.uint32)
        got, last = np.zeros(6, np.int32), np.zeros(6, np.int32)
        bufs = []
        for a_ in (cur, rf, recs, got, last):
            d = lib.avb200_malloc(a_.nbytes)
            assert d and lib.avb200_memcpy_h2d(d, a_.ctypes.data, a_.nbytes, None) == 0
            bufs.append(d)
        assert lib.ff_me_cmp_enc_batch_cuda(15, 1, handle, bufs[0], bufs[1], cur.strides[0], 8, bufs[2], 6, bufs[3], bufs[4], None) == 0, gpu.last_error()
        assert lib.avb200_memcpy_d2h(got.ctypes.data, bufs[3], got.nbytes, None) == 0 and lib.avb200_device_sync() == 0
        for d in bufs:
            lib.avb200_free(d)
        for i in range(6):
            assert want[i] is None or int(got[i]) == want[i], (round_, "batch", i)
        lib.ff_me_cmp_enc_state_free_cuda(handle)


def test_me_cmp_quant_metrics_slots(gpu, checker, orc):
    """ff_me_cmp_enc_init_cuda: the six table entries over a live encoder state (return values and the context fields the C functions write)"""
    import enc_cases
    assert enc_cases.slot_cases(gpu.lib, checker, orc) > 800
    assert "not taken over" in gpu.last_error()          # (the refusal slot_cases() ends with; nothing else was recorded)
    gpu.lib.avb200_clear_error()
