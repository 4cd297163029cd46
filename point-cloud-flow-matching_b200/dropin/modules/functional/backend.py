import _bootstrap  # noqa: F401
from pcfm.backend import pvcnn as _backend

__all__ = ["_backend"]
