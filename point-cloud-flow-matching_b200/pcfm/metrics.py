"""Chamfer-3D and approximate-EMD operators with the reference's Python signatures.

  chamfer_3DFunction / chamfer_3DDist / *_noGrad   <- ChamferDistancePytorch/chamfer3D/dist_chamfer_3D.py:41-133
  chamfer_{2,5,6}DFunction / chamfer_{2,5,6}DDist   <- ChamferDistancePytorch/chamfer6D/dist_chamfer_6D.py:30-82 (and 2D, 5D)
  EarthMoverDistanceFunction / earth_mover_distance <- PyTorchEMD/emd.py:6-51
  earth_mover_distance_nograd                       <- PyTorchEMD/emd_nograd.py:7-44
  fscore                                            <- ChamferDistancePytorch/fscore.py:3-16
Inputs are cast to float32 under autocast exactly like the reference's custom_fwd(cast_inputs=float32).
"""
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .backend import chamfer as chamfer_3D
from .backend import chamfer_2d, chamfer_5d, chamfer_6d
from .backend import emd as emd_cuda


def _check_xyz(t):
    assert t.size(-1) == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"


def _chamfer_forward(xyz1, xyz2):
    _check_xyz(xyz1)
    _check_xyz(xyz2)
    b, n, _ = xyz1.size()
    m = xyz2.size(1)
    dev = xyz1.device
    dist1 = torch.zeros(b, n, device=dev)
    dist2 = torch.zeros(b, m, device=dev)
    idx1 = torch.zeros(b, n, dtype=torch.int32, device=dev)
    idx2 = torch.zeros(b, m, dtype=torch.int32, device=dev)
    chamfer_3D.forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
    return dist1, dist2, idx1, idx2


class chamfer_3DFunction(Function):
    """(xyz1[B,N,3], xyz2[B,M,3]) -> (dist1[B,N], dist2[B,M], idx1 i32[B,N], idx2 i32[B,M]); squared distances."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, xyz1, xyz2):
        out = _chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, out[2], out[3])
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        g1 = torch.zeros_like(xyz1)
        g2 = torch.zeros_like(xyz2)
        chamfer_3D.backward(xyz1, xyz2, g1, g2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)
        return g1, g2


class chamfer_3DDist(nn.Module):
    def forward(self, input1, input2):
        return chamfer_3DFunction.apply(input1.contiguous(), input2.contiguous())


def _make_chamfer_nd(dim, ext):
    """chamfer_<dim>DFunction / chamfer_<dim>DDist with the reference's contract (dist_chamfer_6D.py:30-82)."""

    class _Fn(Function):
        @staticmethod
        def forward(ctx, xyz1, xyz2):
            b, n, d = xyz1.size()
            assert d == dim, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
            _, m, d2 = xyz2.size()
            assert d2 == dim, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
            dev = xyz1.device
            dist1 = torch.zeros(b, n, device=dev)
            dist2 = torch.zeros(b, m, device=dev)
            idx1 = torch.zeros(b, n, dtype=torch.int32, device=dev)
            idx2 = torch.zeros(b, m, dtype=torch.int32, device=dev)
            ext.forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
            ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
            return dist1, dist2, idx1, idx2

        @staticmethod
        def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
            xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
            g1 = torch.zeros_like(xyz1)
            g2 = torch.zeros_like(xyz2)
            ext.backward(xyz1, xyz2, g1, g2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)
            return g1, g2

    class _Dist(nn.Module):
        def forward(self, input1, input2):
            return _Fn.apply(input1.contiguous().float(), input2.contiguous().float())

    _Fn.__name__ = _Fn.__qualname__ = f"chamfer_{dim}DFunction"
    _Dist.__name__ = _Dist.__qualname__ = f"chamfer_{dim}DDist"
    return _Fn, _Dist


chamfer_2DFunction, chamfer_2DDist = _make_chamfer_nd(2, chamfer_2d)
chamfer_5DFunction, chamfer_5DDist = _make_chamfer_nd(5, chamfer_5d)
chamfer_6DFunction, chamfer_6DDist = _make_chamfer_nd(6, chamfer_6d)


class chamfer_3DFunction_noGrad(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        return _chamfer_forward(xyz1, xyz2)


class chamfer_3DDist_nograd(nn.Module):
    def forward(self, input1, input2):
        return chamfer_3DFunction_noGrad.apply(input1.contiguous(), input2.contiguous())


class EarthMoverDistanceFunction(Function):
    """cost[B] = sum d^2 * match;  the gradient flows through matchcost only (match is a constant, emd.py:20-24)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        assert xyz1.is_cuda and xyz2.is_cuda, "Only support cuda currently."
        match = emd_cuda.approxmatch_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, match)
        return emd_cuda.matchcost_forward(xyz1, xyz2, match)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_cost):
        xyz1, xyz2, match = ctx.saved_tensors
        g1, g2 = emd_cuda.matchcost_backward(grad_cost.contiguous(), xyz1, xyz2, match)
        return g1, g2


class EarthMoverDistanceFunctionNoGrad(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        assert xyz1.is_cuda and xyz2.is_cuda, "Only support cuda currently."
        return emd_cuda.matchcost_forward(xyz1, xyz2, emd_cuda.approxmatch_forward(xyz1, xyz2))


def _emd_prepare(xyz1, xyz2, transpose):
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1, xyz2 = xyz1.transpose(1, 2), xyz2.transpose(1, 2)
    assert xyz1.shape[-1] == 3, f"require it to be B,N,3; get: {xyz1.shape}"
    return xyz1, xyz2


def earth_mover_distance(xyz1, xyz2, transpose=True):
    """Approximate EMD per batch element divided by N (emd.py:27-51). transpose=True accepts [B,3,N] inputs."""
    xyz1, xyz2 = _emd_prepare(xyz1, xyz2, transpose)
    return EarthMoverDistanceFunction.apply(xyz1, xyz2) / float(xyz1.shape[1])


def earth_mover_distance_nograd(xyz1, xyz2, transpose=True):
    xyz1, xyz2 = _emd_prepare(xyz1, xyz2, transpose)
    return EarthMoverDistanceFunctionNoGrad.apply(xyz1, xyz2) / float(xyz1.shape[1])


def fscore(dist1, dist2, threshold=0.001):
    """F-score of two clouds from the Chamfer outputs (ChamferDistancePytorch/fscore.py:3-16).

    dist1 [B,N], dist2 [B,M]: SQUARED nearest-neighbour distances as returned by chamfer_3DDist, so `threshold` is a
    squared distance too.  Returns (fscore [B], precision_1 [B], precision_2 [B]): the fraction of points of each
    cloud closer than the threshold to the other cloud and their harmonic mean (0 where both fractions are 0).
    """
    hit1 = dist1.lt(threshold).to(torch.float32).mean(dim=1)
    hit2 = dist2.lt(threshold).to(torch.float32).mean(dim=1)
    f = 2 * hit1 * hit2 / (hit1 + hit2)
    return torch.nan_to_num(f, nan=0.0), hit1, hit2
