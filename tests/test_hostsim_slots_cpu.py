"""CPU: the product's H.264 slot path (libav_b200/csrc/slots.cu + h264dsp.cuh: staging, argument packing, the per-call slot
kernel, the table hooks) compiled unchanged as host C++ (tests/hostsim/) and driven through the same cases as the GPU tests
(tests/slot_cases.py), against the oracle.  Not the product and not a fallback: tests/hostsim/libslots_hostsim.so is only ever
loaded here.  The GPU tests remain the parity tests proper; this keeps the slot code checked while it changes without a GPU."""
import ctypes as C
import os
import subprocess

import pytest

import slot_cases

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
LIB = os.path.join(HERE, "libslots_hostsim.so")


@pytest.fixture(scope="module")
def sim(built):
    root = os.path.dirname(os.path.dirname(HERE))
    deps = [os.path.join(HERE, "slots_hostsim.cpp"), os.path.join(HERE, "shim", "cuda_runtime.h")] + \
        [os.path.join(root, "libav_b200", "csrc", f) for f in ("slots.cu", "h264dsp.cuh", "common.cuh", "scratch.h")] + \
        [os.path.join(root, "include", f) for f in ("avdsp_b200.h", "avdsp_b200_tables.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", os.path.join(HERE, "shim"), "-Wno-unknown-pragmas",
                        "-o", LIB, os.path.join(HERE, "slots_hostsim.cpp")], check=True)
    lib = C.CDLL(LIB)
    lib.avb200_last_error.restype = C.c_char_p
    return lib


def last_error_of(lib):
    return lambda: lib.avb200_last_error().decode()


def test_h264_qpel_and_chroma_slots(sim, checker):
    slot_cases.qpel_and_chroma_cases(sim, last_error_of(sim), checker)


def test_h264dsp_slots(sim, checker):
    slot_cases.h264dsp_cases(sim, last_error_of(sim), checker, weights=False)


def test_h264dsp_slots_422(sim, checker):
    slot_cases.h264dsp_422_cases(sim, last_error_of(sim), checker)


def test_batched_kernels_are_not_simulated(sim):
    from libav_b200 import tables
    c = tables.H264DSPContext()
    sim.ff_h264dsp_init_cuda(C.byref(c), 8, 1)
    import numpy as np
    pix = np.zeros((16, 32), np.uint8)
    c.weight_h264_pixels_tab[0](slot_cases.P(pix), 32, 8, 5, 37, -3)
    assert "not simulated" in sim.avb200_last_error().decode()
    sim.avb200_clear_error()
