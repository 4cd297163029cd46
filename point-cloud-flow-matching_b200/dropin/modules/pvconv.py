import _bootstrap  # noqa: F401
from pcfm.nn import PVConv

__all__ = ["PVConv"]
