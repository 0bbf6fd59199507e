"""canonicalvoting_b200 -- B200-native (sm_100a) implementation of the CanonicalVoting hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed plumbing); all
compute is hand-written CUDA behind the C ABI declared in include/cvb200.h
(canonicalvoting_b200/_C/libcvb200.so).  There is no CPU fallback: importing the
package never loads the library, but the first op call fails loudly when the library
is missing or no CUDA device is present.

Reference-facing modules (same names / signatures as the reference):
    hv_cuda.forward / hv_cuda.backward      houghvoting/src/hv_cuda.cpp:74-77
    hough_voting.vote / HoughVoting         train_joint.py:22-56 (HVFunction, HoughVoting)
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
