import _bootstrap  # noqa: F401
from pcfm.nn import BallQuery

__all__ = ["BallQuery"]
