"""fungp_b200 -- B200-native (sm_100a) implementation of funGp's Gaussian-process hot path.

The product is ``lib/libfungp_b200.so`` (hand-written CUDA behind the plain-C ABI of include/fungp_b200.h);
this package is the host-side mirror of the reference's R interface for that path.  No CPU fallback exists.
"""
from .core import Context, Problem, FgpError, NotPosDefError, KER, DIST, BASIS  # noqa: F401
