import _bootstrap  # noqa: F401
from pcfm.backend import chamfer as chamfer_3D
from pcfm.metrics import (chamfer_3DDist, chamfer_3DDist_nograd, chamfer_3DFunction, chamfer_3DFunction_noGrad)
