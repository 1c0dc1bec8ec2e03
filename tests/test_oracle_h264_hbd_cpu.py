"""CPU: pin the port's 9 / 10-bit H.264 DSP restatement (oracle/port/orc_h264_hbd.c) against the tables the compiled reference fills for
those depths (oracle/refbuild/refapi_h264_hbd.c)."""
import pytest

import hbd_cases


@pytest.mark.parametrize("bits", [9, 10])
def test_port_matches_reference(orc, refo, bits):
    assert hbd_cases.compare(orc, refo, bits, seed=1) > 800


@pytest.mark.parametrize("bits", [8, 9, 10])
def test_pred422_port_matches_reference(orc, refo, bits):
    assert hbd_cases.pred422_compare(orc, refo, bits, seed=1) > 150
