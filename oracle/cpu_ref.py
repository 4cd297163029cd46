"""numpy front-end of oracle/pcfm_oracle.c (plain-C restatement of the reference kernels).  TEST INFRASTRUCTURE.

Pinned against: golden vectors of the compiled reference CUDA kernels run on a B200 (tests/golden/), the
reference's own known answers (ChamferDistancePytorch/unit_test.py:14-35, PyTorchEMD/test_emd_loss.py:7-43,
README.md:116-134) — see tests/test_oracle_cpu.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pcfm_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libpcfm_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
               "-fvisibility=hidden", "-o", _OUT, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def avg_voxelize_forward(feat, coords, r):
    feat, coords = _f(feat), _i(coords)
    b, c, n = feat.shape
    s = r ** 3
    out = np.empty((b, c, s), np.float32)
    ind = np.empty((b, n), np.int32)
    cnt = np.empty((b, s), np.int32)
    lib().oracle_avg_voxelize_forward(_p(feat), _p(coords), b, c, n, r, _p(out), _p(ind), _p(cnt))
    return out, ind, cnt


def avg_voxelize_backward(grad_y, ind, cnt):
    grad_y, ind, cnt = _f(grad_y), _i(ind), _i(cnt)
    b, c, s = grad_y.shape
    n = ind.shape[1]
    gx = np.empty((b, c, n), np.float32)
    lib().oracle_avg_voxelize_backward(_p(grad_y), _p(ind), _p(cnt), b, c, n, s, _p(gx))
    return gx


def trilinear_devoxelize_forward(r, coords, feat):
    coords, feat = _f(coords), _f(feat)
    b, c = feat.shape[:2]
    n = coords.shape[2]
    outs = np.empty((b, c, n), np.float32)
    inds = np.empty((b, 8, n), np.int32)
    wgts = np.empty((b, 8, n), np.float32)
    lib().oracle_trilinear_devoxelize_forward(_p(coords), _p(feat), b, c, n, r, 1, _p(outs), _p(inds), _p(wgts))
    return outs, inds, wgts


def trilinear_devoxelize_backward(grad_y, inds, wgts, r):
    grad_y, inds, wgts = _f(grad_y), _i(inds), _f(wgts)
    b, c, n = grad_y.shape
    gx = np.empty((b, c, r ** 3), np.float32)
    lib().oracle_trilinear_devoxelize_backward(_p(grad_y), _p(inds), _p(wgts), b, c, n, r ** 3, _p(gx))
    return gx


def ball_query(centers, points, radius, u):
    centers, points = _f(centers), _f(points)
    b, _, m = centers.shape
    n = points.shape[2]
    out = np.empty((b, m, u), np.int32)
    lib().oracle_ball_query(_p(centers), _p(points), b, n, m, ctypes.c_float(radius), u, _p(out))
    return out


def grouping_forward(feat, idx):
    feat, idx = _f(feat), _i(idx)
    b, c, n = feat.shape
    m, u = idx.shape[1:]
    out = np.empty((b, c, m, u), np.float32)
    lib().oracle_grouping_forward(_p(feat), _p(idx), b, c, n, m, u, _p(out))
    return out


def grouping_backward(grad_y, idx, n):
    grad_y, idx = _f(grad_y), _i(idx)
    b, c, m, u = grad_y.shape
    gx = np.empty((b, c, n), np.float32)
    lib().oracle_grouping_backward(_p(grad_y), _p(idx), b, c, n, m, u, _p(gx))
    return gx


def gather_forward(feat, idx):
    feat, idx = _f(feat), _i(idx)
    b, c, n = feat.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().oracle_gather_forward(_p(feat), _p(idx), b, c, n, m, _p(out))
    return out


def gather_backward(grad_y, idx, n):
    grad_y, idx = _f(grad_y), _i(idx)
    b, c, m = grad_y.shape
    gx = np.empty((b, c, n), np.float32)
    lib().oracle_gather_backward(_p(grad_y), _p(idx), b, c, n, m, _p(gx))
    return gx


def furthest_point_sampling(coords, m):
    coords = _f(coords)
    b, _, n = coords.shape
    out = np.zeros((b, m), np.int32)
    lib().oracle_furthest_point_sampling(_p(coords), b, n, m, _p(out))
    return out


def three_nn_interpolate_forward(points, centers, feat):
    points, centers, feat = _f(points), _f(centers), _f(feat)
    b, c, m = feat.shape
    n = points.shape[2]
    w = np.empty((b, 3, n), np.float32)
    idx = np.empty((b, 3, n), np.int32)
    out = np.empty((b, c, n), np.float32)
    lib().oracle_three_nn(_p(points), _p(centers), b, n, m, _p(w), _p(idx))
    lib().oracle_three_nn_interpolate(_p(feat), _p(idx), _p(w), b, c, m, n, _p(out))
    return out, idx, w


def three_nn_interpolate_backward(grad_y, idx, w, m):
    grad_y, idx, w = _f(grad_y), _i(idx), _f(w)
    b, c, n = grad_y.shape
    gx = np.empty((b, c, m), np.float32)
    lib().oracle_three_nn_interpolate_backward(_p(grad_y), _p(idx), _p(w), b, c, n, m, _p(gx))
    return gx


def chamfer_forward(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((b, n), np.float32)
    d2 = np.zeros((b, m), np.float32)
    i1 = np.zeros((b, n), np.int32)
    i2 = np.zeros((b, m), np.int32)
    lib().oracle_chamfer_forward(_p(xyz1), _p(xyz2), b, n, m, _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, gd1, gd2, idx1, idx2):
    xyz1, xyz2, gd1, gd2, idx1, idx2 = _f(xyz1), _f(xyz2), _f(gd1), _f(gd2), _i(idx1), _i(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros((b, n, 3), np.float32)
    g2 = np.zeros((b, m, 3), np.float32)
    lib().oracle_chamfer_backward(_p(xyz1), _p(xyz2), _p(gd1), _p(gd2), _p(idx1), _p(idx2), b, n, m, _p(g1), _p(g2))
    return g1, g2


def chamfer_nd_forward(xyz1, xyz2):
    """chamfer{2,5,6}D forward (chamfer6D.cu:11-183); the dimension is the last axis of the inputs."""
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, dim = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((b, n), np.float32)
    d2 = np.zeros((b, m), np.float32)
    i1 = np.zeros((b, n), np.int32)
    i2 = np.zeros((b, m), np.int32)
    lib().oracle_chamfer_nd_forward(_p(xyz1), _p(xyz2), b, n, m, dim, _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_nd_backward(xyz1, xyz2, gd1, gd2, idx1, idx2):
    xyz1, xyz2, gd1, gd2, idx1, idx2 = _f(xyz1), _f(xyz2), _f(gd1), _f(gd2), _i(idx1), _i(idx2)
    b, n, dim = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros((b, n, dim), np.float32)
    g2 = np.zeros((b, m, dim), np.float32)
    lib().oracle_chamfer_nd_backward(_p(xyz1), _p(xyz2), _p(gd1), _p(gd2), _p(idx1), _p(idx2), b, n, m, dim,
                                     _p(g1), _p(g2))
    return g1, g2


def emd_approxmatch(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.empty((b, m, n), np.float32)
    lib().oracle_emd_approxmatch(_p(xyz1), _p(xyz2), b, n, m, _p(match))
    return match


def emd_matchcost(xyz1, xyz2, match):
    xyz1, xyz2, match = _f(xyz1), _f(xyz2), _f(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = np.empty((b,), np.float32)
    lib().oracle_emd_matchcost(_p(xyz1), _p(xyz2), _p(match), b, n, m, _p(cost))
    return cost


def emd_matchcost_backward(grad_cost, xyz1, xyz2, match):
    grad_cost, xyz1, xyz2, match = _f(grad_cost), _f(xyz1), _f(xyz2), _f(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    lib().oracle_emd_matchcost_backward(_p(grad_cost), _p(xyz1), _p(xyz2), _p(match), b, n, m, _p(g1), _p(g2))
    return g1, g2


def earth_mover_distance(xyz1, xyz2):
    """emd.py:27-51 with transpose=False: cost / N."""
    match = emd_approxmatch(xyz1, xyz2)
    return emd_matchcost(xyz1, xyz2, match) / float(np.asarray(xyz1).shape[1])


def voxelization_module(features, coords, r, normalize=True, eps=0.0):
    """modules/voxelization.py:16-25 in torch CPU ops (same op order as the module) + the C avg_voxelize."""
    import torch
    c = torch.as_tensor(np.asarray(coords, dtype=np.float32))
    nc = c - c.mean(2, keepdim=True)
    if normalize:
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5
    else:
        nc = (nc + 1) / 2.0
    nc = torch.clamp(nc * r, 0, r - 1)
    vox = torch.round(nc).to(torch.int32)
    out, ind, cnt = avg_voxelize_forward(features, vox.numpy(), r)
    return out, nc.numpy(), vox.numpy(), ind, cnt
