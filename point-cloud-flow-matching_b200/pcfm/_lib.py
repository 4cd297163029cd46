"""ctypes binding of libpcfm_b200.so.  Every symbol declared in include/pcfm_b200.h is bound here."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCFM_B200_LIB: developer override used for A/B builds of the same library (tests/perf_ops.py)
LIB_PATH = os.environ.get("PCFM_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libpcfm_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
        "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU / PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH)

_p = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pcfm_b200.h one to one
SIGNATURES = {
    "pcfm_abi_version": (_i, []),
    "pcfm_build_info": (ctypes.c_char_p, []),
    "pcfm_launch_count": (ctypes.c_int64, []),
    "pcfm_reset_launch_count": (None, []),
    "pcfm_set_option": (_i, [ctypes.c_char_p, _i]),
    "pcfm_avg_voxelize_forward_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_avg_voxelize_forward": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "pcfm_avg_voxelize_backward_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_avg_voxelize_backward": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_trilinear_devoxelize_forward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "pcfm_trilinear_devoxelize_backward_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_trilinear_devoxelize_backward": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_trilinear_devoxelize_backward_coords_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_trilinear_devoxelize_backward_coords": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_avg_voxelize_forward_dual": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p, _sz, _p]),
    "pcfm_avg_voxelize_forward_planned": (_i, [_p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_trilinear_devoxelize_backward_planned": (_i, [_p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_trilinear_devoxelize_forward_cells_ok": (_i, [_i, _i, _i, _i]),
    "pcfm_trilinear_devoxelize_forward_cells": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _i, _p]),
    "pcfm_ball_query": (_i, [_p, _p, _i, _i, _i, _f, _i, _p, _p]),
    "pcfm_grouping_forward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "pcfm_grouping_backward_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "pcfm_grouping_backward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_gather_features_forward": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "pcfm_gather_features_backward_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_gather_features_backward": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_stream_capture_id": (ctypes.c_ulonglong, [_p]),
    "pcfm_voxel_coords": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "pcfm_furthest_point_sampling_workspace": (_sz, [_i, _i, _i]),
    "pcfm_furthest_point_sampling": (_i, [_p, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_three_nn_interpolate_forward": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "pcfm_three_nn_interpolate_backward_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_three_nn_interpolate_backward": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_chamfer3d_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "pcfm_chamfer3d_forward_workspace": (_sz, [_i, _i, _i]),
    "pcfm_chamfer3d_forward_ws": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "pcfm_chamfer3d_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "pcfm_emd_approxmatch_workspace": (_sz, [_i, _i, _i]),
    "pcfm_emd_approxmatch_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_emd_matchcost_workspace": (_sz, [_i, _i, _i]),
    "pcfm_emd_matchcost_forward": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "pcfm_emd_matchcost_backward": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "pcfm_chamfer_nd_forward": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "pcfm_chamfer_nd_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "pcfm_chamfer3d_allpairs": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "pcfm_emd_allpairs_workspace": (_sz, [_i, _i, _i, _i]),
    "pcfm_emd_allpairs": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not match the header
    _fn.restype = _res
    _fn.argtypes = _args

if lib.pcfm_abi_version() != 1:
    raise ImportError(f"{LIB_PATH}: ABI version {lib.pcfm_abi_version()} != 1")


class PcfmError(RuntimeError):
    pass


def check(rc, what):
    if rc == 0:
        return
    if rc == -1:
        raise PcfmError(f"{what}: bad argument (PCFM_E_BADARG)")
    if rc == -2:
        raise PcfmError(f"{what}: workspace too small (PCFM_E_WORKSPACE)")
    raise PcfmError(f"{what}: CUDA error {rc}")


def launch_count():
    return int(lib.pcfm_launch_count())


def reset_launch_count():
    lib.pcfm_reset_launch_count()


options = {}           # what set_option() was last called with
option_listeners = []  # callables run after every set_option (backend.plan_memo drops its plans)


def set_option(key, value):
    check(lib.pcfm_set_option(key.encode(), int(value)), f"pcfm_set_option({key})")
    options[key] = int(value)
    for fn in option_listeners:
        fn()


# PCFM_B200_OPTIONS="pdl=0,plan=1": developer override, pcfm_set_option pairs applied at import (whole test-suite runs
# under a non-default launch / plan-build mode: tests/gpu_scripts/option_matrix.sh)
for _pair in filter(None, os.environ.get("PCFM_B200_OPTIONS", "").split(",")):
    _k, _v = _pair.split("=")
    set_option(_k.strip(), int(_v))


def build_info():
    return lib.pcfm_build_info().decode()


def stream_capture_id(stream_ptr):
    """Identity of the CUDA-graph capture recording on the stream (0 = not capturing)."""
    return int(lib.pcfm_stream_capture_id(stream_ptr))
