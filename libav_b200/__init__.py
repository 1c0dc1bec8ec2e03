"""libav_b200 -- B200-native (sm_100a) back-end for libav's per-block DSP tables and libswscale's scaler.

The product is the C-ABI shared library libav_b200/libavdsp_b200.so (see include/avdsp_b200.h); this
Python package is only the harness-side binding: `_lib` (ctypes prototypes), `tables` (ctypes mirrors of the
reference's function-pointer tables filled by ff_*_init_cuda) and `synth` (the reference tests' synthetic
input generators, used by tests/ and bench.py).
"""
__all__ = ["build", "synth"]
