"""CPU: the per-line SwsContext slots of the product (host simulation of libav_b200/csrc/sws_slots.cu) plugged into a real reference
SwsContext and driven by the reference's unmodified swscale() -- see tests/sws_dropin_cases.py."""
import ctypes as C

import numpy as np
import pytest

from test_hostsim_slots_cpu import sim          # noqa: F401  (the fixture that builds / loads tests/hostsim/libslots_hostsim.so)


def test_slots_under_the_reference_scheduler(sim, refo):
    import libav_b200._lib as prod
    import sws_dropin_cases as D
    sim.hostsim_sws_context.restype = C.c_void_p
    sim.hostsim_sws_free.argtypes = [C.c_void_p]

    def make_ctx(sf, w, h, df, dw, dh, flags):
        view = np.zeros(32, np.int32)
        assert prod.lib.sws_debug_slot_view_cuda(w, h, sf, dw, dh, df, flags, view.ctypes.data) == 26, (sf, df, hex(flags), prod.last_error())
        return sim.hostsim_sws_context(view.ctypes.data)
    assert D.check(refo, sim, make_ctx, sim.hostsim_sws_free) == 84
    assert sim.avb200_last_error().decode() == ""
