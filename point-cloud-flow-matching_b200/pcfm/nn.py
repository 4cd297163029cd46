"""nn.Modules of the hot path with the reference's constructor signatures, forward contracts and parameter names
(state_dicts of the reference load unchanged): third_party/pvcnn/modules/{voxelization,pvconv,shared_mlp,se,
ball_query,pointnet,loss}.py.  Dense layers (Conv3d / Conv1d / BatchNorm / Linear) stay cuDNN / cuBLAS calls.
"""
import threading
import weakref

import torch
import torch.nn as nn

from . import _lib
from . import functional as F

__all__ = ["Voxelization", "PVConv", "SharedMLP", "SE3d", "BallQuery", "PointNetAModule", "PointNetSAModule",
           "PointNetFPModule", "KLLoss"]


class _VoxelCoordCache(threading.local):
    """Per-thread memo of the coordinate work of Voxelization.forward for the LAST coords tensor seen.

    The PVConv layers of one backbone pass the same `coords` tensor object through (pvconv.py:39 returns it
    unchanged), so every layer after the first repeats identical reductions and elementwise kernels.  An entry is
    valid only while it is the same tensor object with the same autograd version counter (any in-place write
    bumps it) under the same CUDA-graph capture (or outside any): a value computed before a capture must not be
    reused inside it, the graph would replay without the kernels that produce it.  The tensor is held by weak
    reference.  Level 1: centred coords and max norm (shared by all resolutions); level 2: (norm_coords,
    vox_coords) per (resolution, normalize, eps).
    """

    def __init__(self):
        self.ref, self.version, self.capture = None, -1, 0
        self.centred, self.maxnorm, self.by_cfg = None, None, {}

    def entry(self, coords):
        capture = 0
        if coords.is_cuda:
            capture = _lib.stream_capture_id(torch.cuda.current_stream(coords.device).cuda_stream)
        try:
            version = coords._version
        except RuntimeError:   # inference tensors track no version counter: nothing can be memoised on them
            return _VoxelCoordCache()
        if (self.ref is None or self.ref() is not coords or version != self.version
                or capture != self.capture):
            self.ref, self.version, self.capture = weakref.ref(coords), version, capture
            self.centred, self.maxnorm, self.by_cfg = None, None, {}
        return self


_coord_cache = _VoxelCoordCache()


class Voxelization(nn.Module):
    """(features[B,C,N], coords[B,3,N]) -> (voxels[B,C,R,R,R], float voxel coords[B,3,N] in [0,R-1]).

    modules/voxelization.py:16-25.  The two reductions (mean over N, max norm) stay torch ops in the reference's
    order — the voxel index is round_half_even() of the result, so a differently ordered reduction could flip
    indices; the six elementwise ops after them are one kernel (pcfm_voxel_coords, bit-identical to the torch
    ops).  `Voxelization.cache = False` disables the per-coords memo (see _VoxelCoordCache).
    """

    cache = True

    def __init__(self, resolution, normalize=True, eps=0):
        super().__init__()
        self.r = int(resolution)
        self.normalize = normalize
        self.eps = eps

    def _coords(self, coords):
        e = _coord_cache.entry(coords) if Voxelization.cache else _VoxelCoordCache()
        key = (self.r, bool(self.normalize), float(self.eps))
        hit = e.by_cfg.get(key)
        if hit is not None:
            return hit
        if e.centred is None:
            c = coords.detach()
            e.centred = (c - c.mean(2, keepdim=True)).contiguous()
        scale = None
        if self.normalize:
            if e.maxnorm is None:
                e.maxnorm = e.centred.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values
            scale = (e.maxnorm * 2.0 + self.eps).reshape(-1)
        if e.centred.is_cuda and e.centred.dtype == torch.float32:
            nc, vox = F.voxel_coords(e.centred, scale, self.r)
        else:  # other dtypes: the reference's op sequence (the CUDA kernels themselves are fp32-only)
            nc = e.centred / scale.view(-1, 1, 1) + 0.5 if self.normalize else (e.centred + 1) / 2.0
            nc = torch.clamp(nc * self.r, 0, self.r - 1)
            vox = torch.round(nc).to(torch.int32)
        e.by_cfg[key] = (nc, vox)
        return nc, vox

    def forward(self, features, coords):
        nc, vox = self._coords(coords)
        if nc.is_cuda and nc.dtype == torch.float32 and nc.is_contiguous():
            return F.avg_voxelize(features, vox, self.r, nc), nc   # nc: lets a training pass build both cell plans at once
        return F.avg_voxelize(features, vox, self.r), nc

    def extra_repr(self):
        return f"resolution={self.r}" + (f", normalized eps = {self.eps}" if self.normalize else "")


class SharedMLP(nn.Module):
    """Stack of 1x1 Conv(1d|2d)+BN+ReLU; extra tuple items pass through (shared_mlp.py:6-33)."""

    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        try:
            conv, bn = {1: (nn.Conv1d, nn.BatchNorm1d), 2: (nn.Conv2d, nn.BatchNorm2d)}[dim]
        except KeyError:
            raise ValueError from None
        widths = list(out_channels) if isinstance(out_channels, (list, tuple)) else [out_channels]
        seq = []
        for w in widths:
            seq += [conv(in_channels, w, 1), bn(w), nn.ReLU(True)]
            in_channels = w
        self.layers = nn.Sequential(*seq)

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            return (self.layers(inputs[0]), *inputs[1:])
        return self.layers(inputs)


class SE3d(nn.Module):
    """Squeeze-excite over a voxel grid (se.py:6-17)."""

    def __init__(self, channel, reduction=8):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def forward(self, inputs):
        gate = self.fc(inputs.mean(-1).mean(-1).mean(-1))
        return inputs * gate.view(inputs.shape[0], inputs.shape[1], 1, 1, 1)


class PVConv(nn.Module):
    """Point-voxel convolution (pvconv.py:11-39): voxelize -> 2x(Conv3d,BN3d,LeakyReLU)[+SE] -> devoxelize,
    fused with a point-wise SharedMLP branch.  forward((features[B,Cin,N], coords[B,3,N])) -> (features, coords).

    `PVConv.channels_last` (class attribute, default False): the voxel grid is handed to the Conv3d /
    BatchNorm3d stack in torch.channels_last_3d memory format (NDHWC), where cuDNN runs its tensor-core kernels; the
    grid comes back to the reference's NCDHW layout in front of the devoxelization, so the point-voxel primitives, the
    parameters and the results are unchanged (same values up to cuDNN's algorithm choice).  It is a host-side layout
    switch around library calls: two extra transposes of the grid per block and direction, paid back by the faster
    convolutions (bench.py `pvconv_modules`).
    """

    channels_last = False

    def __init__(self, in_channels, out_channels, kernel_size, resolution, with_se=False, normalize=True, eps=0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.resolution = kernel_size, resolution
        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)
        pad = kernel_size // 2
        seq = []
        for cin in (in_channels, out_channels):
            seq += [nn.Conv3d(cin, out_channels, kernel_size, stride=1, padding=pad),
                    nn.BatchNorm3d(out_channels, eps=1e-4), nn.LeakyReLU(0.1, True)]
        if with_se:
            seq.append(SE3d(out_channels))
        self.voxel_layers = nn.Sequential(*seq)
        self.point_features = SharedMLP(in_channels, out_channels)

    def forward(self, inputs):
        features, coords = inputs
        grid, vcoords = self.voxelization(features, coords)
        if PVConv.channels_last and grid.is_cuda:
            grid = self.voxel_layers(grid.contiguous(memory_format=torch.channels_last_3d))
            grid = grid.contiguous(memory_format=torch.contiguous_format)   # the primitives keep the reference layout
        else:
            grid = self.voxel_layers(grid)
        from_voxels = F.trilinear_devoxelize(grid, vcoords, self.resolution, self.training)
        return from_voxels + self.point_features(features), coords


class BallQuery(nn.Module):
    """Neighbourhood grouping (modules/ball_query.py:9-30)."""

    def __init__(self, radius, num_neighbors, include_coordinates=True):
        super().__init__()
        self.radius, self.num_neighbors, self.include_coordinates = radius, num_neighbors, include_coordinates

    def forward(self, points_coords, centers_coords, points_features=None):
        points_coords, centers_coords = points_coords.contiguous(), centers_coords.contiguous()
        idx = F.ball_query(centers_coords, points_coords, self.radius, self.num_neighbors)
        rel = F.grouping(points_coords, idx) - centers_coords.unsqueeze(-1)
        if points_features is None:
            assert self.include_coordinates, "No Features For Grouping"
            return rel
        grouped = F.grouping(points_features, idx)
        return torch.cat([rel, grouped], dim=1) if self.include_coordinates else grouped

    def extra_repr(self):
        return "radius={}, num_neighbors={}{}".format(self.radius, self.num_neighbors,
                                                     ", include coordinates" if self.include_coordinates else "")


def _as_nested(out_channels, n):
    if not isinstance(out_channels, (list, tuple)):
        return [[out_channels]] * n
    if not isinstance(out_channels[0], (list, tuple)):
        return [out_channels] * n
    return out_channels


class PointNetAModule(nn.Module):
    """Global max-pool PointNet block (pointnet.py:12-46)."""

    def __init__(self, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        nested = _as_nested(out_channels, 1)
        extra = 3 if include_coordinates else 0
        self.mlps = nn.ModuleList([SharedMLP(in_channels + extra, oc, dim=1) for oc in nested])
        self.include_coordinates = include_coordinates
        self.out_channels = sum(oc[-1] for oc in nested)

    def forward(self, inputs):
        features, coords = inputs
        if self.include_coordinates:
            features = torch.cat([features, coords], dim=1)
        coords = torch.zeros((coords.size(0), 3, 1), device=coords.device)
        pooled = [mlp(features).max(dim=-1, keepdim=True).values for mlp in self.mlps]
        return (torch.cat(pooled, dim=1) if len(pooled) > 1 else pooled[0]), coords

    def extra_repr(self):
        return f"out_channels={self.out_channels}, include_coordinates={self.include_coordinates}"


class PointNetSAModule(nn.Module):
    """Set abstraction: FPS -> ball query/grouping -> SharedMLP -> max (pointnet.py:49-92)."""

    def __init__(self, num_centers, radius, num_neighbors, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        radius = list(radius) if isinstance(radius, (list, tuple)) else [radius]
        if not isinstance(num_neighbors, (list, tuple)):
            num_neighbors = [num_neighbors] * len(radius)
        assert len(radius) == len(num_neighbors)
        nested = _as_nested(out_channels, len(radius))
        assert len(radius) == len(nested)
        extra = 3 if include_coordinates else 0
        self.groupers = nn.ModuleList([BallQuery(radius=r, num_neighbors=k, include_coordinates=include_coordinates)
                                       for r, k in zip(radius, num_neighbors)])
        self.mlps = nn.ModuleList([SharedMLP(in_channels + extra, oc, dim=2) for oc in nested])
        self.num_centers = num_centers
        self.out_channels = sum(oc[-1] for oc in nested)

    def forward(self, inputs):
        features, coords = inputs
        centers = F.furthest_point_sample(coords, self.num_centers)
        outs = [mlp(g(coords, centers, features)).max(dim=-1).values for g, mlp in zip(self.groupers, self.mlps)]
        return (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]), centers

    def extra_repr(self):
        return f"num_centers={self.num_centers}, out_channels={self.out_channels}"


class PointNetFPModule(nn.Module):
    """Feature propagation by 3-NN interpolation (pointnet.py:95-111)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.mlp = SharedMLP(in_channels=in_channels, out_channels=out_channels, dim=1)

    def forward(self, inputs):
        if len(inputs) == 3:
            points_coords, centers_coords, centers_features = inputs
            points_features = None
        else:
            points_coords, centers_coords, centers_features, points_features = inputs
        x = F.nearest_neighbor_interpolate(points_coords, centers_coords, centers_features)
        if points_features is not None:
            x = torch.cat([x, points_features], dim=1)
        return self.mlp(x), points_coords


class KLLoss(nn.Module):
    def forward(self, x, y):
        return F.kl_loss(x, y)
