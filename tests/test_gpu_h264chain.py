"""config 3 as a chain at its full size: 1920x1088 pictures of 64 slices, MC -> residual -> deblock on the GPU through the batched C-ABI
calls (the sequence bench.py times) against the reference's own tables driven over the same records on the host
(oracle/refbuild/refapi_h264pic.c, itself pinned to the per-function checker in tests/test_oracle_h264pic_cpu.py)."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from test_oracle_h264pic_cpu import run_driver

pytestmark = pytest.mark.gpu


def gpu_chain(gpu, w, P, pics):
    from libav_b200 import device
    lib = gpu.lib
    mb_w, mb_h = w["mb_w"], w["mb_h"]
    W, H = 16 * mb_w, 16 * mb_h
    dev = lambda a: device.DevBuf.from_numpy(np.ascontiguousarray(a))
    d_ref = [[dev(np.concatenate([p] * P)) for p in r] for r in w["refs"]]
    planes = np.array([[p.ptr for p in r] for r in d_ref], np.uint64)
    d_planes = dev(planes)
    mcs, ress = [], []
    for k in range(P):
        m = w["mc"].copy(); m["y"] = m["y"] + k * H; mcs.append(m)
        r = w["res"].copy(); r["luma_off"] = r["luma_off"] + k * W * H; r["chroma_off"] = r["chroma_off"] + k * W * H // 4; ress.append(r)
    mc, res = np.concatenate(mcs), np.concatenate(ress)
    d_mc, d_res = dev(mc), dev(res)
    d_nnz, d_dbk = dev(np.concatenate([w["nnzc"]] * P)), dev(np.concatenate([w["dbk"]] * P))
    d_coef = dev(np.concatenate([w["coeffs"]] * P))
    d_y = dev(np.concatenate([p[0] for p in pics])); d_cb = dev(np.concatenate([p[1] for p in pics])); d_cr = dev(np.concatenate([p[2] for p in pics]))
    d_prog = dev(np.zeros(2 * mb_h * P + 64, np.uint32))
    gpu.check(lib.ff_h264_mc_batch_cuda(d_mc.ptr, mc.shape[0], d_planes.ptr, d_y.ptr, d_cb.ptr, d_cr.ptr, W, W // 2, W, H, None), "mc")
    gpu.check(lib.ff_h264_idct_add_mb_batch_cuda(d_res.ptr, res.shape[0], d_coef.ptr, 768, d_nnz.ptr, d_y.ptr, d_cb.ptr, d_cr.ptr, W, W // 2, None), "residual")
    gpu.check(lib.ff_h264_deblock_batch_cuda(d_dbk.ptr, mb_w, mb_h, P, d_y.ptr, d_cb.ptr, d_cr.ptr, W, W // 2, d_prog.ptr, None), "deblock")
    device.sync()
    return (d_y.download(np.uint8, (P * H, W)), d_cb.download(np.uint8, (P * H // 2, W // 2)), d_cr.download(np.uint8, (P * H // 2, W // 2)),
            d_coef.download(np.int16, (P * mb_w * mb_h, 768)))


@pytest.mark.parametrize("mb_w,mb_h,slices,P", [(120, 68, 64, 3), (120, 68, 1, 2), (45, 30, 7, 4)])
def test_config3_chain_matches_the_reference_driver(gpu, refo, mb_w, mb_h, slices, P):
    w = synth.h264_config3_picture(mb_w, mb_h, slices, seed=3)
    pics = [synth.h264_picture(mb_w, mb_h, seed=90 + k) for k in range(P)]
    gy, gcb, gcr, gco = gpu_chain(gpu, w, P, pics)
    y = np.concatenate([p[0] for p in pics]); cb = np.concatenate([p[1] for p in pics]); cr = np.concatenate([p[2] for p in pics])
    co = np.concatenate([w["coeffs"]] * P)
    assert run_driver(refo, w, P, y, cb, cr, co, 8) == 0
    assert np.array_equal(gy, y), np.argwhere(gy != y)[:4].tolist()
    assert np.array_equal(gcb, cb) and np.array_equal(gcr, cr)
    assert np.array_equal(gco, co), "consumed coefficients"
