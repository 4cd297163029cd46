import _bootstrap  # noqa: F401
from pcfm.backend import chamfer_6d as chamfer_6D
from pcfm.metrics import chamfer_6DDist, chamfer_6DFunction
