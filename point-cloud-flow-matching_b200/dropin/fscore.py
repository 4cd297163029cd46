"""`from fscore import fscore` (ChamferDistancePytorch/fscore.py:3-16)."""
import _bootstrap  # noqa: F401
from pcfm.metrics import fscore  # noqa: F401
