import _bootstrap  # noqa: F401
from pcfm.nn import SE3d

__all__ = ["SE3d"]
