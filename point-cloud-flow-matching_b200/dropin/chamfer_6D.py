"""`import chamfer_6D` (chamfer6D/chamfer_cuda.cpp:30-33): module-level forward / backward, caller-allocated outputs."""
import _bootstrap  # noqa: F401
from pcfm.backend import chamfer_6d as _impl

forward = _impl.forward
backward = _impl.backward
