import _bootstrap  # noqa: F401
from pcfm.functional import NeighborInterpolation, nearest_neighbor_interpolate

__all__ = ["nearest_neighbor_interpolate"]
