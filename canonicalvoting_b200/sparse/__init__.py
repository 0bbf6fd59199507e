"""canonicalvoting_b200.sparse -- the subset of MinkowskiEngine the reference uses,
re-designed for B200: device hash map + output-stationary neighbour tables (csrc/sparse_coords.cu) and
implicit-GEMM sparse convolution (csrc/sparse_conv*.cu).  The top-level package `MinkowskiEngine/` re-exports
this module under the reference's names so that `import MinkowskiEngine as ME` (train_joint.py:9,
utils/minkunet.py:28) binds to it."""
from . import utils  # noqa: F401
from .coords import CoordinateManager  # noqa: F401
from .functional import get_forward_mode, set_forward_mode  # noqa: F401
from .modules import (BasicBlock, Bottleneck, MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiConvolutionTranspose,  # noqa: F401
                      MinkowskiReLU, SparseTensor, cat)
