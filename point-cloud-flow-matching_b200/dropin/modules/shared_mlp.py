import _bootstrap  # noqa: F401
from pcfm.nn import SharedMLP

__all__ = ["SharedMLP"]
