"""MinkowskiEngine.utils subset (batched_coordinates, sparse_quantize, kaiming_normal_)."""
from canonicalvoting_b200.sparse.utils import batched_coordinates, kaiming_normal_, sparse_quantize  # noqa: F401
