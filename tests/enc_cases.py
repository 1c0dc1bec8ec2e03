"""Shared cases for MECmpContext.quant_psnr / bit / rd (libavcodec/me_cmp.c:621-782): encoder states and block pairs.

Used by the CPU oracle test (port against the compiled reference) and by the GPU tests (product against the compiled reference)."""
import ctypes as C
import itertools

import numpy as np

from oracle.loader import OrcEncState

KINDS = {14: "quant_psnr", 15: "bit", 16: "rd"}


class Tables:
    """VLC length tables shaped like a codec's (uni_*_rl_len: [run * 128 + level + 64], the DC table [diff + 256]); the functions only index them."""

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.intra_len = rng.integers(2, 30, 64 * 128, dtype=np.uint8)
        self.intra_last = rng.integers(2, 30, 64 * 128, dtype=np.uint8)
        self.inter_len = rng.integers(2, 30, 64 * 128, dtype=np.uint8)
        self.inter_last = rng.integers(2, 30, 64 * 128, dtype=np.uint8)
        self.luma_dc = rng.integers(1, 24, 512, dtype=np.uint8)


def make_state(tables, fdct_sel=0, dequant=0, qscale=4, mb_intra=0, seed=0, alternate_scan=0, h263_aic=0, ac_pred=0, flat=False):
    rng = np.random.default_rng(1000 + seed)
    st = OrcEncState()
    st.fdct_sel, st.dequant, st.qscale, st.mb_intra = fdct_sel, dequant, qscale, mb_intra
    st.y_dc_scale, st.c_dc_scale = (8, 8) if dequant != 3 else (int(rng.integers(8, 20)), 9)
    st.h263_aic, st.ac_pred, st.alternate_scan = h263_aic, ac_pred, alternate_scan
    st.intra_quant_bias = int(rng.choice([96, 0, 64]))          # 3 << (QUANT_BIAS_SHIFT - 3) is the MPEG default (mpegvideo_enc.c:330-340)
    st.inter_quant_bias = int(rng.choice([0, -64, -32]))
    st.ac_esc_length = int(rng.integers(18, 31))
    im = np.full(64, 16, np.uint16) if flat else rng.integers(8, 84, 64).astype(np.uint16)
    nm = np.full(64, 16, np.uint16) if flat else rng.integers(8, 64, 64).astype(np.uint16)
    im[0] = 8
    for i in range(64):
        st.intra_matrix[i], st.inter_matrix[i] = int(im[i]), int(nm[i])
    st.intra_ac_vlc_length, st.intra_ac_vlc_last_length = tables.intra_len.ctypes.data, tables.intra_last.ctypes.data
    st.inter_ac_vlc_length, st.inter_ac_vlc_last_length = tables.inter_len.ctypes.data, tables.inter_last.ctypes.data
    st.luma_dc_vlc_length = tables.luma_dc.ctypes.data
    return st


def states(tables):
    """(label, state) over transform x quantiser family x intra x qscale (+ scan / aic / ac_pred variants)"""
    out = []
    k = 0
    for fdct_sel, dequant, mb_intra in itertools.product((0, 2), (0, 1, 2, 3), (0, 1)):
        for qscale in (1, 3, 9, 31):
            k += 1
            alt = int(dequant in (1, 2) and qscale == 3)
            aic = int(dequant == 3 and qscale == 9)
            acp = int(dequant == 3 and qscale == 31)
            out.append(("fdct%d_dq%d_intra%d_q%d" % (fdct_sel, dequant, mb_intra, qscale),
                        make_state(tables, fdct_sel, dequant, qscale, mb_intra, seed=k, alternate_scan=alt, h263_aic=aic, ac_pred=acp, flat=(k % 5 == 0))))
    return out


def block_pairs(n, seed, stride=48):
    """n pairs of 16x16 areas (cur, ref) inside two planes: smooth, noisy and high-contrast differences"""
    rng = np.random.default_rng(seed)
    rows = 16 * n
    cur = np.zeros((rows, stride), np.uint8)
    ref = np.zeros((rows, stride), np.uint8)
    for i in range(n):
        base = rng.integers(0, 256, (16, 16))
        mode = i % 4
        if mode == 0:   d = rng.integers(-3, 4, (16, 16))                       # mostly quantised away
        elif mode == 1: d = rng.integers(-40, 41, (16, 16))
        elif mode == 2: d = rng.integers(-255, 256, (16, 16))
        else:           d = (np.add.outer(np.arange(16), np.arange(16)) * rng.integers(1, 9) - 60)   # a ramp: strong low frequencies
        c = np.clip(base + d, 0, 255)
        cur[16 * i:16 * i + 16, 8:24] = c
        ref[16 * i:16 * i + 16, 8:24] = base
    return cur, ref


def oracle_scores(o, kind, sidx, st, cur, ref, n, h, h263_guard=None):
    """scores + (block_last_index[0], mb_intra) after each call; h263_guard = a port to ask whether the reference would hit its own
    assert (dct_unquantize_h263_inter_c on an all-zero block, mpegvideo.c:252): those records are reported as None"""
    scores, sides = [], []
    side = (C.c_int32 * 2)()
    for i in range(n):
        a = cur.ctypes.data + 16 * i * cur.strides[0] + 8
        b = ref.ctypes.data + 16 * i * ref.strides[0] + 8
        if st.h263_aic and st.mb_intra and kind != 14:
            # advanced intra coding quantises the DC with a step of 8: a difference block whose mean exceeds ~31 gives a level outside the
            # 512-entry DC length table and the reference reads past it (me_cmp.c:675, "FIXME: chroma"); undefined there, left out here
            blk = cur[16 * i:16 * i + (16 if not sidx and h == 16 else 8), 8:8 + (8 if sidx else 16)].astype(np.int64) - \
                ref[16 * i:16 * i + (16 if not sidx and h == 16 else 8), 8:8 + (8 if sidx else 16)]
            sums = [abs(int(blk[y:y + 8, x:x + 8].sum())) for y in range(0, blk.shape[0], 8) for x in range(0, blk.shape[1], 8)]
            if max(sums) > 1900:
                scores.append(None); sides.append(None)
                continue
        if h263_guard is not None and st.dequant == 3 and kind == 14:
            lasts = []
            for blk in range(1 if sidx else (4 if h == 16 else 2)):
                off = 8 * (blk & 1) + 8 * (blk >> 1) * cur.strides[0]
                h263_guard.me_cmp_quant(kind, 1, C.byref(st), a + off, b + off, cur.strides[0], 8, side)
                lasts.append(side[0])
            if min(lasts) < 0:
                scores.append(None); sides.append(None)
                continue
        r = o.me_cmp_quant(kind, sidx, C.byref(st), a, b, cur.strides[0], h, side)
        scores.append(r); sides.append((side[0], side[1]))
    return scores, sides


# ---- the product (libav_b200/csrc/me_cmp_enc.cu), on the GPU or host-simulated ---------------------------------------------------------

def product_state(st, checker):
    """FFMECmpEncState for an OrcEncState: the q matrices and scan table come from the checker's ff_convert_matrix / scan tables"""
    from libav_b200 import tables
    qi, qn, scan = np.zeros(64, np.int32), np.zeros(64, np.int32), np.zeros(64, np.uint8)
    checker.enc_qmatrices(C.byref(st), qi.ctypes.data, qn.ctypes.data, scan.ctypes.data)
    p = tables.FFMECmpEncState()
    p.fdct, p.dequant, p.qscale, p.mb_intra, p.y_dc_scale, p.h263_aic = st.fdct_sel, st.dequant, st.qscale, st.mb_intra, st.y_dc_scale, st.h263_aic
    p.intra_quant_bias, p.inter_quant_bias, p.ac_esc_length = st.intra_quant_bias, st.inter_quant_bias, st.ac_esc_length
    for i in range(64):
        p.q_intra_matrix[i], p.q_inter_matrix[i] = int(qi[i]), int(qn[i])
        p.intra_matrix[i], p.inter_matrix[i], p.scantable[i] = st.intra_matrix[i], st.inter_matrix[i], int(scan[i])
    return p, qi, qn, scan


def vlc_tables(T):
    from libav_b200 import tables
    v = tables.FFMECmpVlcTables()
    v.intra_ac_vlc_length, v.intra_ac_vlc_last_length = T.intra_len.ctypes.data, T.intra_last.ctypes.data
    v.inter_ac_vlc_length, v.inter_ac_vlc_last_length = T.inter_len.ctypes.data, T.inter_last.ctypes.data
    v.luma_dc_vlc_length = T.luma_dc.ctypes.data
    return v


def batch_cases(lib, run_batch, checker, guard, n=40):
    """run_batch(kind, sidx, handle, cur, ref, recs (n, 2) uint32, h) -> (scores int32[n], last_index int32[n]); every state x kind x size
    against the checker record by record"""
    T = Tables(3)
    vlc = vlc_tables(T)
    checked = 0
    for kind in sorted(KINDS):
        cur, rf = block_pairs(n, seed=10 + kind)
        recs = np.array([[16 * i * cur.strides[0] + 8, 16 * i * rf.strides[0] + 8] for i in range(n)], np.uint32)
        for label, st in states(T):
            p, _, _, _ = product_state(st, checker)
            handle = lib.ff_me_cmp_enc_state_cuda(C.byref(p), C.byref(vlc) if kind != 14 else None)
            assert handle, label
            for sidx, h in ((1, 8), (0, 16), (0, 8)):
                want, wside = oracle_scores(checker, kind, sidx, st, cur, rf, n, h, h263_guard=guard)
                got, last = run_batch(kind, sidx, handle, cur, rf, recs, h)
                for i in range(n):
                    if want[i] is None:
                        continue
                    assert int(got[i]) == want[i] and int(last[i]) == wside[i][0], (KINDS[kind], label, sidx, h, i, int(got[i]), want[i], int(last[i]), wside[i])
                    checked += 1
            lib.ff_me_cmp_enc_state_free_cuda(handle)
    return checked


class FakeEncoder:
    """The MpegEncContext fields the slots read, as numpy cells a FFMECmpEncView points into (the glue of examples/reference_binding fills the
    same view with offsets into the real struct)."""

    def __init__(self, T):
        from libav_b200 import tables
        self.T = T
        self.ints = np.zeros(16, np.int32)          # 0 qscale 1 y_dc_scale 2 h263_aic 3 intra_bias 4 inter_bias 5 esc 6 mb_intra; 8.. block_last_index
        self.qi = np.zeros((32, 64), np.int32)
        self.qn = np.zeros((32, 64), np.int32)
        self.qi_p = C.c_void_p(self.qi.ctypes.data)
        self.qn_p = C.c_void_p(self.qn.ctypes.data)
        self.im = np.zeros(64, np.uint16)
        self.nm = np.zeros(64, np.uint16)
        self.scan = np.zeros(64, np.uint8)
        self.vlc_p = (C.c_void_p * 5)(T.intra_len.ctypes.data, T.intra_last.ctypes.data, T.inter_len.ctypes.data, T.inter_last.ctypes.data, T.luma_dc.ctypes.data)
        self.key = np.zeros(8, np.uint8)            # stands in for the opaque MpegEncContext *
        v = tables.FFMECmpEncView()
        cell = lambda k: self.ints.ctypes.data + 4 * k
        v.qscale, v.y_dc_scale, v.h263_aic, v.intra_quant_bias, v.inter_quant_bias, v.ac_esc_length, v.mb_intra = [cell(k) for k in range(7)]
        v.block_last_index = cell(8)
        v.q_intra_matrix, v.q_inter_matrix = C.addressof(self.qi_p), C.addressof(self.qn_p)
        v.intra_matrix, v.inter_matrix, v.scantable = self.im.ctypes.data, self.nm.ctypes.data, self.scan.ctypes.data
        base = C.addressof(self.vlc_p)
        v.intra_ac_vlc_length, v.intra_ac_vlc_last_length, v.inter_ac_vlc_length, v.inter_ac_vlc_last_length, v.luma_dc_vlc_length = [base + 8 * k for k in range(5)]
        v.idct_perm_none = v.plain_quantiser = 1
        self.view = v

    def load(self, st, checker):
        p, qi, qn, scan = product_state(st, checker)
        self.ints[:7] = [st.qscale, st.y_dc_scale, st.h263_aic, st.intra_quant_bias, st.inter_quant_bias, st.ac_esc_length, st.mb_intra]
        self.ints[8] = -2
        self.qi[st.qscale], self.qn[st.qscale] = qi, qn
        self.im[:], self.nm[:], self.scan[:] = list(st.intra_matrix), list(st.inter_matrix), scan
        self.view.fdct, self.view.dequant = st.fdct_sel, st.dequant


def slot_cases(lib, checker, guard, n=6):
    """ff_me_cmp_enc_init_cuda: the six table entries driven like the encoder drives them (live context fields change between calls), return
    values and the context fields the C functions write (mb_intra, block_last_index[0]) against the checker"""
    from libav_b200 import tables
    T = Tables(4)
    enc = FakeEncoder(T)
    table = tables.MECmpContext()
    u8p = C.POINTER(C.c_uint8)
    checked = 0
    key = C.c_void_p(enc.key.ctypes.data)
    for label, st in states(T)[::3]:
        enc.load(st, checker)
        assert lib.ff_me_cmp_enc_init_cuda(C.byref(table), key, C.byref(enc.view)) == 0, label
        for kind, slots in ((14, table.quant_psnr), (15, table.bit), (16, table.rd)):
            cur, rf = block_pairs(n, seed=20 + kind)
            for sidx, h in ((1, 8), (0, 16), (0, 8)):
                want, wside = oracle_scores(checker, kind, sidx, st, cur, rf, n, h, h263_guard=guard)
                for i in range(n):
                    if want[i] is None:
                        continue
                    enc.ints[6], enc.ints[8] = st.mb_intra, -2
                    a = C.cast(cur.ctypes.data + 16 * i * cur.strides[0] + 8, u8p)
                    b = C.cast(rf.ctypes.data + 16 * i * rf.strides[0] + 8, u8p)
                    got = slots[sidx](key, a, b, cur.strides[0], h)
                    assert got == want[i] and (int(enc.ints[8]), int(enc.ints[6])) == wside[i], (KINDS[kind], label, sidx, h, i, got, want[i], enc.ints[6:9], wside[i])
                    checked += 1
    assert all(not table.dct_sad[k] for k in range(6))          # nothing else is touched
    lib.ff_me_cmp_enc_uninit_cuda(key)
    # refusals leave the table alone
    t2 = tables.MECmpContext()
    enc.view.idct_perm_none = 0
    assert lib.ff_me_cmp_enc_init_cuda(C.byref(t2), key, C.byref(enc.view)) == -1 and not t2.rd[0]
    enc.view.idct_perm_none = 1
    return checked
