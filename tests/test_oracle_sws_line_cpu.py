"""CPU: pin the port's restatement of libswscale's per-line slot functions (oracle/port/orc_sws.c sws_line_*) against the functions the
compiled reference installs in its own SwsContext (oracle/refbuild/refapi.c line_ctx: sws_getContext -> c->hyScale, c->yuv2packedX ...)."""
import sws_line_cases as L


def test_port_matches_reference(orc, refo):
    assert L.compare(L.OracleCalls(orc), L.OracleCalls(refo), seed=1) > 150
    for o in (orc, refo):
        o.sws_set_colorspace((L.C.c_int * 4)(*L.FCC), 0, 3000, 70000, 60000)
    try:
        assert L.compare(L.OracleCalls(orc), L.OracleCalls(refo), seed=2) > 150
    finally:
        for o in (orc, refo):
            o.sws_set_colorspace(None, 0, 0, 0, 0)
