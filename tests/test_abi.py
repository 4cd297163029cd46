"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/pcfm_b200.h
declares, the ctypes table matches, and the reference-named import trees resolve (no compute calls: no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import PKG, ROOT

HEADER = os.path.join(ROOT, "include", "pcfm_b200.h")
LIB = os.path.join(PKG, "lib", "libpcfm_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcfm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        subprocess.run([sys.executable, os.path.join(PKG, "build.py")], check=True)
    return LIB


def test_header_symbols_exported(built):
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(built)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in pcfm_b200.h but not exported: {missing}"
    exported = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    extra = sorted(set(re.findall(r" T (pcfm_[a-z0-9_]+)", exported)) - set(names))
    assert not extra, f"exported but not declared in the header: {extra}"


def test_ctypes_table_matches_header(built):
    from pcfm import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib.pcfm_abi_version() == 1
    assert b"sm_100a" in _lib.lib.pcfm_build_info()


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_workspace_queries_are_host_only(built):
    from pcfm import _lib
    l = _lib.lib
    assert l.pcfm_avg_voxelize_forward_workspace(8, 128, 20000, 32) > 8 * 20000 * 4
    assert l.pcfm_trilinear_devoxelize_backward_workspace(8, 128, 20000, 32) > 8 * 8 * 20000 * 4
    assert l.pcfm_emd_approxmatch_workspace(2, 1024, 1024) >= 2 * 2 * 2048 * 4
    # argument validation happens before any CUDA call
    assert l.pcfm_avg_voxelize_forward(None, None, 1, 1, 1, 0, None, None, None, None, 0, None) == -1
    assert l.pcfm_chamfer3d_allpairs(None, None, 1, 1, 0, 0, None, None) == -1
    # plan-reuse entries: no plan workspace / no coordinates -> PCFM_E_BADARG, empty problems -> 0, all without CUDA
    assert l.pcfm_avg_voxelize_forward_planned(None, 1, 1, 1, 8, None, None, 0, None) == -1
    assert l.pcfm_trilinear_devoxelize_backward_planned(None, 1, 1, 1, 8, None, None, 0, None) == -1
    assert l.pcfm_avg_voxelize_forward_planned(None, 0, 4, 100, 8, None, None, 0, None) == 0
    assert l.pcfm_trilinear_devoxelize_backward_planned(None, 2, 0, 100, 8, None, None, 0, None) == 0
    assert l.pcfm_avg_voxelize_forward_dual(None, None, None, 1, 1, 1, 8, None, None, None, None, 0, None, 0, None) == -1
    assert l.pcfm_avg_voxelize_forward_dual(None, None, None, 1, 1, 1, 300, None, None, None, None, 0, None, 0, None) == -1
    assert l.pcfm_set_option(b"devox_cl", 0) == 0
    assert l.pcfm_set_option(b"nonsense", 1) == -1
    assert l.pcfm_set_option(b"stage", 0) == 0


def test_dropin_import_trees(built):
    sys.path.insert(0, os.path.join(PKG, "dropin"))
    try:
        from modules.pvconv import PVConv          # models.py:14
        from modules.shared_mlp import SharedMLP   # models.py:15
        import modules.functional as F
        from modules.functional.backend import _backend
        import chamfer_3D
        import emd_ext
        from chamfer3D.dist_chamfer_3D import chamfer_3DDist, chamfer_3DFunction
        from third_party.PyTorchEMD.emd import earth_mover_distance
        from third_party.PyTorchEMD.backend import emd_cuda_dynamic
        from third_party.PyTorchEMD.emd_nograd import earth_mover_distance_nograd
        from fscore import fscore                  # ChamferDistancePytorch/fscore.py
    finally:
        sys.path.pop(0)
    for name in ["avg_voxelize_forward", "avg_voxelize_backward", "trilinear_devoxelize_forward",
                 "trilinear_devoxelize_backward", "ball_query", "grouping_forward", "grouping_backward",
                 "gather_features_forward", "gather_features_backward", "furthest_point_sampling",
                 "three_nearest_neighbors_interpolate_forward", "three_nearest_neighbors_interpolate_backward"]:
        assert callable(getattr(_backend, name)), name   # bindings.cpp:10-37
    assert callable(chamfer_3D.forward) and callable(chamfer_3D.backward)
    for name in ["approxmatch_forward", "matchcost_forward", "matchcost_backward"]:
        assert callable(getattr(emd_ext, name)) and callable(getattr(emd_cuda_dynamic, name))
    # parameter names must match the reference so its checkpoints load (pvconv.py:20-31, se.py:9-14)
    keys = set(PVConv(8, 16, 3, 8, with_se=True).state_dict())
    for k in ["voxel_layers.0.weight", "voxel_layers.1.running_mean", "voxel_layers.3.bias", "voxel_layers.6.fc.0.weight",
              "voxel_layers.6.fc.2.weight", "point_features.layers.0.weight", "point_features.layers.1.weight"]:
        assert k in keys, k
    assert set(SharedMLP(4, [8, 16]).state_dict()) >= {"layers.0.weight", "layers.3.weight", "layers.4.running_var"}


def test_argument_checks_raise_like_the_reference(built):
    """utils.hpp:7-18: non-CUDA / wrong dtype inputs raise RuntimeError before any kernel is launched."""
    import torch
    from pcfm.backend import pvcnn
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pvcnn.avg_voxelize_forward(torch.zeros(1, 2, 4), torch.zeros(1, 3, 4, dtype=torch.int32), 4)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pvcnn.ball_query(torch.zeros(1, 3, 4), torch.zeros(1, 3, 8), 0.1, 2)


def test_product_does_not_import_oracle():
    """The product path must never reach into oracle/ (or /root/reference)."""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "/root/reference" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_fscore_matches_the_reference_definition(built):
    """ChamferDistancePytorch/fscore.py:3-16: precision = share of squared NN distances below the threshold,
    F = harmonic mean, 0 (not NaN) where neither cloud has a hit."""
    import torch
    from pcfm.metrics import fscore
    d1 = torch.tensor([[0.0, 0.5e-3, 2e-3, 1.0], [1.0, 1.0, 1.0, 1.0]])
    d2 = torch.tensor([[0.0, 1e-4], [0.5, 2.0]])
    f, p1, p2 = fscore(d1, d2)
    assert torch.allclose(p1, torch.tensor([0.5, 0.0])) and torch.allclose(p2, torch.tensor([1.0, 0.0]))
    assert torch.allclose(f, torch.tensor([2 * 0.5 * 1.0 / 1.5, 0.0]))
    f2, _, _ = fscore(d1, d2, threshold=10.0)
    assert torch.equal(f2, torch.ones(2))
