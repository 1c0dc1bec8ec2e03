"""GPU parity: MECmpContext metrics (every slot the C table fills), the exhaustive motion search (config 4 geometry and
tie / border cases), HpelDSPContext (all tabs x sizes x half-pel positions, including the reference's avg 2x2 xy2
quirk) and the four forward DCTs -- all against the CPU oracle."""
import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr
from h264_util import at

pytestmark = pytest.mark.gpu


def _dev(a):
    from libav_b200 import device
    return device.DevBuf.from_numpy(a)


KINDS = [(0, s, d) for s in (0, 1) for d in range(4)] + [(1, 0, 0), (1, 1, 0), (2, 0, 0), (2, 1, 0), (2, 2, 0), (3, 0, 0), (3, 1, 0),
                                                         (4, 0, 0), (5, 0, 0), (6, 0, 0), (6, 1, 0), (7, 0, 0), (7, 1, 0), (8, 0, 0),
                                                         (8, 1, 0), (9, 0, 0), (9, 1, 0)]


@pytest.mark.parametrize("kind,sidx,dxy", KINDS)
def test_me_cmp_batch(gpu, checker, kind, sidx, dxy):
    from libav_b200 import device
    rng = np.random.default_rng(kind * 100 + sidx * 10 + dxy)
    W, H = 256, 96
    cur = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    ref = np.clip(cur.astype(int) + rng.integers(-25, 26, size=cur.shape), 0, 255).astype(np.uint8)
    w = 16 >> sidx
    for h in ([8, 16] if (w == 16 or kind in (3, 7)) and not (kind in (3, 7) and sidx == 1) else [8]):
        n = 500
        rec = np.zeros(n, dtype=synth.MECMP_DT)
        rec["cur_off"] = rng.integers(0, H - 18, size=n) * W + (rng.integers(0, (W - 20) // w, size=n) * w)     # blk1 is width-aligned
        rec["ref_off"] = rng.integers(0, H - 18, size=n) * W + rng.integers(0, W - 20, size=n)
        want = np.array([checker.me_cmp(kind, sidx, dxy, at(cur, r["cur_off"]), at(ref, r["ref_off"]), W, h) for r in rec], dtype=np.int32)
        assert (want >= 0).all()
        d_out, d_cur, d_ref, d_rec = device.DevBuf(4 * n), _dev(cur), _dev(ref), _dev(rec)      # keep the buffers alive
        gpu.check(gpu.lib.ff_me_cmp_batch_cuda(kind, sidx, dxy, d_cur.ptr, d_ref.ptr, W, h, d_rec.ptr, n, d_out.ptr, None))
        device.sync()
        assert np.array_equal(d_out.download(np.int32, (n,)), want), (kind, sidx, dxy, h)


def test_sum_abs_dctelem_and_null_slots(gpu, checker):
    from libav_b200 import device
    rng = np.random.default_rng(0)
    blocks = rng.integers(-2048, 2048, size=(300, 64)).astype(np.int16)
    rec = np.zeros(300, dtype=synth.MECMP_DT)
    rec["cur_off"] = np.arange(300) * 128
    want = np.array([checker.me_cmp(10, 0, 0, ptr(b), None, 0, 0) for b in blocks], dtype=np.int32)
    d_out = device.DevBuf(1200)
    db, d_rec = _dev(blocks), _dev(rec)
    gpu.check(gpu.lib.ff_me_cmp_batch_cuda(10, 0, 0, db.ptr, db.ptr, 0, 0, d_rec.ptr, 300, d_out.ptr, None))
    device.sync()
    assert np.array_equal(d_out.download(np.int32, (300,)), want)
    assert gpu.lib.ff_me_cmp_batch_cuda(4, 1, 0, db.ptr, db.ptr, 16, 8, d_rec.ptr, 1, d_out.ptr, None) == -1   # vsad[1] is NULL in C too
    gpu.lib.avb200_clear_error()


@pytest.mark.parametrize("w,h", [(96, 64), (1920, 1080 // 16 * 16)])
def test_full_search(gpu, checker, w, h):
    from libav_b200 import device
    cur, ref = synth.me_frames(w, h, seed=2)
    flat = np.full((h, w), 77, np.uint8)
    for c, r in ((cur, ref), (flat, flat)):
        n = (w // 16) * (h // 16)
        want = np.zeros(3 * n, np.int32)
        checker.full_search(ptr(c), ptr(r), w, w, h, 16, 0, h // 16, ptr(want), 16)
        d_out = device.DevBuf(12 * n)
        dc, dr = _dev(c), _dev(r)
        # two shards, as a 2-GPU frame split would run them
        half = (h // 16) // 2
        gpu.check(gpu.lib.ff_full_search_cuda(dc.ptr, dr.ptr, w, w, h, 16, 0, half, d_out.ptr, None))
        gpu.check(gpu.lib.ff_full_search_cuda(dc.ptr, dr.ptr, w, w, h, 16, half, h // 16, d_out.ptr, None))
        device.sync()
        assert np.array_equal(d_out.download(np.int32, (3 * n,)), want)
        if w > 1000 and c is cur:
            break


def test_hpel_batch(gpu, checker):
    from libav_b200 import device
    rng = np.random.default_rng(3)
    W, H = 512, 256
    src = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    dst = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    recs = []
    cells = [(by, bx) for by in range(0, H - 32, 32) for bx in range(0, W - 32, 32)]
    k = 0
    for tab in range(4):
        for sidx in range(4):
            if (tab == 2 and sidx > 1) or (tab == 3 and sidx != 0):
                continue
            for dxy in range(4):
                w = 16 >> sidx
                for h in ([8, 16] if w == 16 else [4, 8] if w == 8 else [4] if w == 4 else [2]):
                    by, bx = cells[k % len(cells)]
                    k += 1
                    recs.append((by * W + bx, (by + 3) * W + bx + 5, tab, sidx, dxy, h))
    assert k <= len(cells)
    rec = np.array(recs, dtype=synth.HPEL_DT)
    want = dst.copy()
    for r in rec:
        assert checker.hpel(int(r["tab"]), int(r["sidx"]), int(r["dxy"]), at(want, r["dst_off"]), at(src, r["src_off"]), W, int(r["h"])) == 0
    d_dst, d_rec, d_src = _dev(dst), _dev(rec), _dev(src)
    gpu.check(gpu.lib.ff_hpel_batch_cuda(d_rec.ptr, rec.shape[0], d_dst.ptr, d_src.ptr, W, None))
    device.sync()
    assert np.array_equal(d_dst.download(np.uint8, dst.shape), want)


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_fdct_batch(gpu, checker, which):
    from libav_b200 import device
    rng = np.random.default_rng(which)
    blocks = rng.integers(-256, 256, size=(5000, 64)).astype(np.int16)
    blocks[::10] = rng.integers(-1024, 1024, size=(500, 64))
    want = blocks.copy()
    for b in want:
        checker.fdct(which, ptr(b))
    d = _dev(blocks)
    gpu.check(gpu.lib.ff_fdct_batch_cuda(which, d.ptr, 5000, None))
    device.sync()
    assert np.array_equal(d.download(np.int16, blocks.shape), want)


@pytest.mark.parametrize("kind", [11, 12, 13])
def test_me_cmp_dct_metrics(gpu, checker, kind):
    """dct_sad / dct_max / dct264_sad: the encoder-state metrics that only need the DSP tables, batch call vs checker"""
    from libav_b200 import device
    r = np.random.RandomState(kind)
    w, h = 176, 144
    cur = r.randint(0, 256, (h, w)).astype(np.uint8)
    ref = np.clip(cur.astype(int) + r.randint(-20, 21, cur.shape), 0, 255).astype(np.uint8)
    n = 600
    rec = np.zeros(n, synth.MECMP_DT)
    rec["cur_off"] = r.randint(0, h - 16, n) * w + r.randint(0, w - 16, n)
    rec["ref_off"] = r.randint(0, h - 16, n) * w + r.randint(0, w - 16, n)
    d_c, d_r, d_rec = device.DevBuf.from_numpy(cur), device.DevBuf.from_numpy(ref), device.DevBuf.from_numpy(rec)
    out = device.DevBuf(4 * n)
    for sidx, hh in ((0, 16), (0, 8), (1, 8)):
        for sel in ((0, 2) if kind != 13 else (0,)):
            gpu.check(gpu.lib.ff_me_cmp_batch_cuda(kind, sidx, sel, d_c.ptr, d_r.ptr, w, hh, d_rec.ptr, n, out.ptr, None))
            device.sync()
            got = out.download(np.int32, (n,))
            want = np.array([checker.me_cmp(kind, sidx, sel, cur.ctypes.data + int(q["cur_off"]), ref.ctypes.data + int(q["ref_off"]), w, hh) for q in rec], np.int32)
            assert np.array_equal(got, want), (kind, sidx, hh, sel, np.flatnonzero(got != want)[:5].tolist())
