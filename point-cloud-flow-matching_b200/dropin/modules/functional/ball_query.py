import _bootstrap  # noqa: F401
from pcfm.functional import ball_query

__all__ = ["ball_query"]
