"""MinkowskiEngine.modules.resnet_block (imported at utils/minkunet.py:30, utils/resnet.py:29)."""
from canonicalvoting_b200.sparse.modules import BasicBlock, Bottleneck  # noqa: F401
