"""pcfm — host-side mirror of the reference's operator API over libpcfm_b200.so (sm_100a CUDA, C ABI).

Layout
  pcfm._lib        ctypes binding of include/pcfm_b200.h (fails loudly if the library is missing)
  pcfm.backend     drop-in for the reference's three pybind modules:
                     backend.pvcnn   == `_pvcnn_backend`  (third_party/pvcnn/modules/functional/src/bindings.cpp:10-37)
                     backend.chamfer == `chamfer_3D`      (ChamferDistancePytorch/chamfer3D/chamfer_cuda.cpp:30-33)
                     backend.emd     == `emd_ext`         (PyTorchEMD/cuda/emd.cpp:23-27)
  pcfm.evaluation  block-sharded all-pairs CD / EMD and MMD / COV / 1-NNA
The reference-named import trees (`modules.*`, `chamfer_3D`, `third_party.PyTorchEMD.*`) live in ../dropin/.
There is NO CPU fallback anywhere in this package.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
