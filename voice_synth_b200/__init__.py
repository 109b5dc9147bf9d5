"""voice_synth_b200 -- thin Python harness over libvoicesynth_cuda (the C ABI in include/voicesynth.h).

The product is the shared library and the C tools in host/; this package only exists so that tests
and bench.py can drive the same entry points through ctypes.  There is no CPU fallback: importing
`voice_synth_b200.api` fails loudly when the library has not been built.
"""
from .api import (Context, FilterParams, FlowParams, VsError, flow_nsamples, lib_path,  # noqa: F401
                  VS_F_JITTER, VS_F_NOISE, VS_F_SHIMMER)
