"""pysdr_b200 — B200-native receive-DSP path of aa2il/pySDR behind the reference's own L1 operator surface.

    import pysdr_b200.sig_proc as dsp        # drop-in for `import sig_proc as dsp` (reference receiver.py:45)

Everything numeric runs in libpysdr_b200.so (hand-written sm_100a CUDA, C ABI in include/pysdr_b200.h).
There is no CPU fallback: without the built library or without a CUDA device the package raises.
"""
from ._lib import PysdrError, load  # noqa: F401

__version__ = "0.1.0"
