import _bootstrap  # noqa: F401
from pcfm.functional import Grouping, grouping

__all__ = ["grouping"]
