import _bootstrap  # noqa: F401
from pcfm.nn import KLLoss

__all__ = ["KLLoss"]
