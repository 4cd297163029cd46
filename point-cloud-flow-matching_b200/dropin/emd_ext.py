"""`import emd_ext` (README.md:91-93)."""
import _bootstrap  # noqa: F401
from pcfm.backend import emd as _impl

approxmatch_forward = _impl.approxmatch_forward
matchcost_forward = _impl.matchcost_forward
matchcost_backward = _impl.matchcost_backward
