"""`import chamfer_3D` (README.md:116-134): module-level forward / backward with caller-allocated outputs."""
import _bootstrap  # noqa: F401
from pcfm.backend import chamfer as _impl

forward = _impl.forward
backward = _impl.backward
