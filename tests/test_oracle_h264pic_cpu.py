"""The batched config-3 driver of the compiled reference (oracle/refbuild/refapi_h264pic.c: MC -> residual per macroblock, loop filter per
slice, pthreads over slices) against the per-function checker chained from Python in the same order (tests/h264_util.py) -- so that the
thing bench.py times on the host cores is known to compute the pictures the GPU parity tests are held to."""
import ctypes as C

import numpy as np
import pytest

from libav_b200 import synth
from oracle.loader import ptr
import h264_util as hu


def run_driver(o, w, P, y, cb, cr, coeffs, nthreads):
    refs = (C.c_void_p * 6)(*[p.ctypes.data for r in w["refs"] for p in r])
    return o.h264_pictures(ptr(w["mc"]), ptr(w["mc_first"]), ptr(w["res"]), ptr(w["dbk"]), ptr(w["nnzc"]), ptr(coeffs), refs, 2,
                           ptr(y), ptr(cb), ptr(cr), y.strides[0], cb.strides[0], w["mb_w"], w["mb_h"], P, w["slices"], nthreads)


@pytest.mark.parametrize("mb_w,mb_h,slices,threads", [(12, 8, 4, 1), (20, 9, 7, 5), (45, 30, 64, 8)])
def test_driver_equals_the_chained_functions(refo, mb_w, mb_h, slices, threads):
    w = synth.h264_config3_picture(mb_w, mb_h, slices, seed=mb_w)
    P = 3
    pics = [synth.h264_picture(mb_w, mb_h, seed=70 + k) for k in range(P)]
    y = np.concatenate([p[0] for p in pics]); cb = np.concatenate([p[1] for p in pics]); cr = np.concatenate([p[2] for p in pics])
    coeffs = np.concatenate([w["coeffs"]] * P)
    assert run_driver(refo, w, P, y, cb, cr, coeffs, threads) == 0
    H = 16 * mb_h
    for k in range(P):
        wy, wcb, wcr = pics[k][0].copy(), pics[k][1].copy(), pics[k][2].copy()
        co = w["coeffs"].copy()
        hu.oracle_mc(refo, w["mc"], w["refs"], wy, wcb, wcr)
        hu.oracle_residual(refo, w["res"], co, w["nnzc"], wy, wcb, wcr)
        hu.oracle_deblock(refo, w["dbk"], mb_w, mb_h, wy, wcb, wcr)
        assert np.array_equal(y[k * H:(k + 1) * H], wy), k
        assert np.array_equal(cb[k * H // 2:(k + 1) * H // 2], wcb) and np.array_equal(cr[k * H // 2:(k + 1) * H // 2], wcr), k
        assert np.array_equal(coeffs[k * mb_w * mb_h:(k + 1) * mb_w * mb_h], co), "consumed coefficients"
