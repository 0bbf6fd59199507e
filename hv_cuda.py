"""`import hv_cuda` -- module name of the reference's native extension
(houghvoting/setup.py:7, houghvoting/src/hv_cuda.cpp:74-77).  With this repository on
sys.path the reference scripts' `import hv_cuda` binds to the B200 implementation."""
from canonicalvoting_b200.hv_cuda import backward, forward  # noqa: F401
