"""CPU: the per-line SwsContext slots of the product (host simulation of libav_b200/csrc/sws_slots.cu) plugged into a real reference
SwsContext and driven by the reference's unmodified swscale() -- see tests/sws_dropin_cases.py."""
import ctypes as C

import numpy as np
import pytest

from test_hostsim_slots_cpu import sim          # noqa: F401  (the fixture that builds / loads tests/hostsim/libslots_hostsim.so)


def test_slots_under_the_reference_scheduler(sim, refo):
    import sws_dropin_cases as D

    def make_ctx(sf, w, h, df, dw, dh, flags):
        ctx = sim.sws_getContext_cuda(w, h, sf, dw, dh, df, flags, None, None, None)
        assert ctx, (sf, df, hex(flags), sim.avb200_last_error())
        return ctx
    assert D.check(refo, sim, make_ctx, sim.sws_freeContext_cuda) == 84
    assert sim.avb200_last_error().decode() == ""
