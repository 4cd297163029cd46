"""Operator API of the hot path: the reference's `modules.functional.*` names over the B200 backend.

Mirrors third_party/pvcnn/modules/functional/{voxelization,devoxelization,ball_query,grouping,sampling,
interpolatation,loss}.py — same callables, argument order, gradients (features only; coordinates never get a
gradient, functional/voxelization.py:37, devoxelization.py:39) and the same failure mode for backward through
an eval-mode devoxelize (nothing was saved, devoxelization.py:24-26).
"""
import numpy as np
import torch
import torch.nn.functional as TF
from torch.autograd import Function

from .backend import pvcnn as _backend

__all__ = ["avg_voxelize", "trilinear_devoxelize", "ball_query", "grouping", "gather", "furthest_point_sample",
           "logits_mask", "nearest_neighbor_interpolate", "kl_loss", "huber_loss", "voxel_coords"]


class AvgVoxelization(Function):
    """features f32[B,C,N], coords (integer voxel coords) [B,3,N], resolution[, float_coords] -> f32[B,C,R,R,R]."""

    @staticmethod
    def forward(ctx, features, coords, resolution, float_coords=None):
        feats = features.contiguous()
        vox = coords.int().contiguous()
        r = int(resolution)
        # float_coords (extension, what pcfm.nn.Voxelization passes): where the block will devoxelize; the trilinear
        # cell plan of those coordinates — used by the devoxelize forward and, in training, by its backward — is built
        # by the same kernel as the plan of this call
        fc = float_coords
        out, ind, cnt = _backend.avg_voxelize_forward(feats, vox, r, fc)
        ctx.save_for_backward(ind, cnt)
        return out.view(feats.shape[0], feats.shape[1], r, r, r)

    @staticmethod
    def backward(ctx, grad_output):
        ind, cnt = ctx.saved_tensors
        b, c = grad_output.shape[0], grad_output.shape[1]
        flat = grad_output.contiguous().view(b, c, -1)
        return _backend.avg_voxelize_backward(flat, ind, cnt), None, None, None


class TrilinearDevoxelization(Function):
    """features f32[B,C,R,R,R], coords f32[B,3,N] in [0,R-1], resolution, is_training -> f32[B,C,N]."""

    @staticmethod
    def forward(ctx, features, coords, resolution, is_training=True):
        b, c = features.shape[0], features.shape[1]
        grid = features.contiguous().view(b, c, -1)
        coords = coords.contiguous()
        outs, inds, wgts = _backend.trilinear_devoxelize_forward(int(resolution), bool(is_training), coords, grid)
        if is_training:
            # inds / wgts are saved like the reference does; the coordinates ride along for the cell-plan backward
            ctx.save_for_backward(inds, wgts, coords)
            ctx.r = int(resolution)
        return outs

    @staticmethod
    def backward(ctx, grad_output):
        inds, wgts, coords = ctx.saved_tensors  # eval mode saved nothing: unpacking fails, as in the reference
        r = ctx.r
        if r <= _backend.MAX_FAST_DEVOX_BWD_R:
            g = _backend.trilinear_devoxelize_backward_coords(grad_output.contiguous(), coords, r)
        else:
            g = _backend.trilinear_devoxelize_backward(grad_output.contiguous(), inds, wgts, r)
        return g.view(grad_output.shape[0], grad_output.shape[1], r, r, r), None, None, None


class Grouping(Function):
    """features f32[B,C,N], indices i32[B,M,U] -> f32[B,C,M,U]."""

    @staticmethod
    def forward(ctx, features, indices):
        feats, idx = features.contiguous(), indices.contiguous()
        ctx.save_for_backward(idx)
        ctx.num_points = feats.shape[-1]
        return _backend.grouping_forward(feats, idx)

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        return _backend.grouping_backward(grad_output.contiguous(), idx, ctx.num_points), None


class Gather(Function):
    """features f32[B,C,N], indices [B,M] -> f32[B,C,M]."""

    @staticmethod
    def forward(ctx, features, indices):
        feats, idx = features.contiguous(), indices.int().contiguous()
        ctx.save_for_backward(idx)
        ctx.num_points = feats.shape[-1]
        return _backend.gather_features_forward(feats, idx)

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        return _backend.gather_features_backward(grad_output.contiguous(), idx, ctx.num_points), None


class NeighborInterpolation(Function):
    """points f32[B,3,N], centers f32[B,3,M], centers_features f32[B,C,M] -> f32[B,C,N] (3-NN, inverse distance)."""

    @staticmethod
    def forward(ctx, points_coords, centers_coords, centers_features):
        pts, ctr = points_coords.contiguous(), centers_coords.contiguous()
        out, idx, wgt = _backend.three_nearest_neighbors_interpolate_forward(pts, ctr, centers_features.contiguous())
        ctx.save_for_backward(idx, wgt)
        ctx.num_centers = ctr.shape[-1]
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, wgt = ctx.saved_tensors
        g = _backend.three_nearest_neighbors_interpolate_backward(grad_output.contiguous(), idx, wgt, ctx.num_centers)
        return None, None, g


avg_voxelize = AvgVoxelization.apply
trilinear_devoxelize = TrilinearDevoxelization.apply
grouping = Grouping.apply
gather = Gather.apply
nearest_neighbor_interpolate = NeighborInterpolation.apply


def voxel_coords(centred, scale, resolution):
    """Fused elementwise tail of Voxelization.forward (modules/voxelization.py:19-24): centred f32[B,3,N], scale
    f32[B] or None -> (norm_coords f32[B,3,N], vox_coords i32[B,3,N]).  No gradient (the reference detaches coords)."""
    with torch.no_grad():
        return _backend.voxel_coords(centred.contiguous(), None if scale is None else scale.contiguous(), resolution)


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """centers f32[B,3,M], points f32[B,3,N] -> i32[B,M,U] (first U neighbours inside the ball, index order)."""
    return _backend.ball_query(centers_coords.contiguous(), points_coords.contiguous(), radius, num_neighbors)


def furthest_point_sample(coords, num_samples):
    """coords f32[B,3,N] -> coordinates f32[B,3,M] of the iteratively furthest points (index 0 first)."""
    coords = coords.contiguous()
    return gather(coords, _backend.furthest_point_sampling(coords, num_samples))


def logits_mask(coords, logits, num_points_per_object):
    """Sample M foreground points per object from binary logits (functional/sampling.py:51-84).

    Returns (selected_coords f32[B,3,M] centred on the foreground mean, foreground mean f32[B,3], mask bool[B,N]).
    """
    bsz, _, npts = coords.shape
    mask = logits[:, 0, :] < logits[:, 1, :]
    n_fg = mask.sum(dim=-1, keepdim=True)
    fg_coords = coords * mask.view(bsz, 1, npts)
    fg_mean = fg_coords.sum(dim=-1) / torch.max(n_fg, torch.ones_like(n_fg)).float()
    picked = torch.zeros((bsz, num_points_per_object), device=coords.device, dtype=torch.int32)
    for i in range(bsz):
        cand = mask[i].nonzero().view(-1)
        k = cand.numel()
        if k >= num_points_per_object:
            sel = np.random.choice(k, num_points_per_object, replace=False)
        elif k > 0:
            sel = np.concatenate([np.arange(k).repeat(num_points_per_object // k),
                                  np.random.choice(k, num_points_per_object % k, replace=False)])
            np.random.shuffle(sel)
        else:
            continue
        picked[i] = cand[sel]
    return gather(fg_coords - fg_mean.view(bsz, -1, 1), picked), fg_mean, mask


def kl_loss(x, y):
    p = TF.softmax(x.detach(), dim=1)
    return torch.mean(torch.sum(p * (torch.log(p) - TF.log_softmax(y, dim=1)), dim=1))


def huber_loss(error, delta):
    a = torch.abs(error)
    q = torch.clamp(a, max=delta)
    return torch.mean(0.5 * q * q + delta * (a - q))
