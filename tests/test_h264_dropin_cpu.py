"""CPU: the loop-filter slots of the product's H264DSPContext (host simulation of libav_b200/csrc/slots.cu) under the reference's own
deblocking driver -- see tests/h264_dropin_cases.py."""
from test_hostsim_slots_cpu import sim          # noqa: F401


def test_loop_filter_slots_under_the_reference_driver(sim, refo):
    import h264_dropin_cases as D
    assert D.check(refo, sim) == 18
    assert sim.avb200_last_error().decode() == ""
