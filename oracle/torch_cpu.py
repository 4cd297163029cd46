"""The reference path expressed in stock PyTorch CPU ops — the timed `cpu_baseline` / `--impl reference` arm.

north_star: "the same ops expressed in PyTorch CPU (scatter_add voxelization, grid_sample-equivalent devoxelize,
cdist+min), timed on the GPU box's host cores".  TEST / BENCH INFRASTRUCTURE ONLY, never on the product path.
Semantics follow vox.cu:31,64-69, trilinear_devox.cu:38-103 (grid order (z,y,x): flat index is x*r^2+y*r+z) and
train.py:83-84 / chamfer_python.py:18-43.
"""
import torch
import torch.nn.functional as TF


def normalize_coords(coords, r, normalize=True, eps=0.0):
    nc = coords - coords.mean(2, keepdim=True)
    if normalize:
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5
    else:
        nc = (nc + 1) / 2.0
    return torch.clamp(nc * r, 0, r - 1)


def avg_voxelize(features, vox_coords, r):
    b, c, n = features.shape
    ind = (vox_coords[:, 0] * r * r + vox_coords[:, 1] * r + vox_coords[:, 2]).long()
    cnt = torch.zeros(b, r ** 3, dtype=features.dtype).scatter_add_(1, ind, torch.ones(b, n, dtype=features.dtype))
    w = 1.0 / cnt.gather(1, ind)
    out = torch.zeros(b, c, r ** 3, dtype=features.dtype)
    out = out.scatter_add(2, ind[:, None, :].expand(b, c, n), features * w[:, None, :])
    return out.view(b, c, r, r, r)


def trilinear_devoxelize(voxels, coords, r):
    """voxels [B,C,r,r,r], coords float [B,3,N] in [0,r-1] -> [B,C,N]."""
    b, c = voxels.shape[:2]
    g = coords.permute(0, 2, 1) / (r - 1) * 2 - 1  # [B,N,3] (x,y,z)
    grid = torch.stack([g[..., 2], g[..., 1], g[..., 0]], dim=-1)[:, :, None, None, :]  # (z,y,x) <-> (W,H,D)
    out = TF.grid_sample(voxels, grid, mode="bilinear", padding_mode="border", align_corners=True)
    return out.view(b, c, -1)


def pvconv_primitives_fwd_bwd(features, coords, r):
    """voxelize fwd -> devoxelize fwd -> (sum).backward(): the four primitives of one PVConv, no Conv3d."""
    features = features.detach().requires_grad_(True)
    nc = normalize_coords(coords, r)
    vox = avg_voxelize(features, torch.round(nc).to(torch.int64), r)
    out = trilinear_devoxelize(vox, nc, r)
    out.sum().backward()
    return out.detach(), features.grad


def chamfer(xyz1, xyz2):
    d = torch.cdist(xyz1, xyz2).pow(2)
    d1, i1 = d.min(2)
    d2, i2 = d.min(1)
    return d1, d2, i1.int(), i2.int()


def chamfer_allpairs(s, r):
    """[ns,N,3] x [nr,M,3] -> [ns,nr] of dist1.mean()+dist2.mean()."""
    out = torch.empty(s.shape[0], r.shape[0])
    for i in range(s.shape[0]):
        d = torch.cdist(s[i:i + 1].expand(r.shape[0], -1, -1), r).pow(2)
        out[i] = d.min(2).values.mean(1) + d.min(1).values.mean(1)
    return out
