import _bootstrap  # noqa: F401
from pcfm.backend import chamfer_5d as chamfer_5D
from pcfm.metrics import chamfer_5DDist, chamfer_5DFunction
