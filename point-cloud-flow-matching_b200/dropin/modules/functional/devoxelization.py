import _bootstrap  # noqa: F401
from pcfm.functional import TrilinearDevoxelization, trilinear_devoxelize

__all__ = ["trilinear_devoxelize"]
