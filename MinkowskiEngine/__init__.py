"""`import MinkowskiEngine as ME` -- the package name the reference imports (train_joint.py:9,
utils/minkunet.py:28, utils/resnet.py:28).  With this repository on sys.path the name binds to the
B200-native sparse-voxel stack (canonicalvoting_b200.sparse), which implements exactly the subset of
MinkowskiEngine 0.5.x the reference's hot path uses."""
from canonicalvoting_b200.sparse import (BasicBlock, CoordinateManager, MinkowskiBatchNorm, MinkowskiConvolution,  # noqa: F401
                                         MinkowskiConvolutionTranspose, MinkowskiReLU, SparseTensor, cat)
from . import modules, utils  # noqa: F401

__version__ = "0.5.3+cvb200"
