import _bootstrap  # noqa: F401
from pcfm.backend import chamfer_2d as chamfer_2D
from pcfm.metrics import chamfer_2DDist, chamfer_2DFunction
