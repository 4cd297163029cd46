import _bootstrap  # noqa: F401
from pcfm.functional import Gather, furthest_point_sample, gather, logits_mask

__all__ = ["gather", "furthest_point_sample", "logits_mask"]
