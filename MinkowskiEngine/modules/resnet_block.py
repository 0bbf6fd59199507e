"""MinkowskiEngine.modules.resnet_block (imported at utils/minkunet.py:30, utils/resnet.py:29)."""
from canonicalvoting_b200.sparse.modules import BasicBlock  # noqa: F401


class Bottleneck:  # only MinkUNet50/101 use it; the reference's scripts never instantiate those
    expansion = 4

    def __init__(self, *a, **k):
        raise NotImplementedError("Bottleneck blocks (MinkUNet50/101) are outside the reference's hot path")
