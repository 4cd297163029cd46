"""Generate golden vectors by running the UNMODIFIED reference CUDA extensions (oracle/_ref, built by
oracle/build_ref.py from /root/reference) on a GPU.  Run on the B200 box:

    python tests/golden/make_golden.py            # writes tests/golden/ref_b200.npz (and gpurun_out/ref_b200.npz)
    python tests/golden/make_golden.py chamfer6d  # writes tests/golden/ref_b200_chamfer6d.npz

Only integer outputs (bit-exact) and small float outputs are stored; inputs are regenerated from seeds (cases.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import cases  # noqa: E402
from oracle import build_ref  # noqa: E402


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def n(x):
    return x.detach().cpu().numpy()


def chamfer6d():
    """tests/golden/ref_b200_chamfer6d.npz: the compiled reference chamfer_6D on the 6-D cases."""
    c6 = build_ref.load_ref("chamfer6d")
    assert c6, "oracle/_ref/chamfer6d is not built"
    out = {}
    for name, kw in cases.CHAMFER6D.items():
        a, c = cases.cloud_pair(**kw)
        b, nn_, m = a.shape[0], a.shape[1], c.shape[1]
        d1 = torch.zeros(b, nn_).cuda(); d2 = torch.zeros(b, m).cuda()
        i1 = torch.zeros(b, nn_, dtype=torch.int32).cuda(); i2 = torch.zeros(b, m, dtype=torch.int32).cuda()
        assert c6.forward(t(a), t(c), d1, d2, i1, i2) == 1
        g1 = torch.zeros(b, nn_, 6).cuda(); g2 = torch.zeros(b, m, 6).cuda()
        gd1 = torch.ones(b, nn_).cuda(); gd2 = (torch.arange(m).float() / m).repeat(b, 1).cuda()
        c6.backward(t(a), t(c), g1, g2, gd1, gd2, i1, i2)
        out[name + "/dist1"] = n(d1); out[name + "/dist2"] = n(d2)
        out[name + "/idx1"] = n(i1); out[name + "/idx2"] = n(i2)
        out[name + "/g1"] = n(g1); out[name + "/g2"] = n(g2)
    for dst in (os.path.join(HERE, "ref_b200_chamfer6d.npz"), os.path.join(ROOT, "gpurun_out", "ref_b200_chamfer6d.npz")):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        np.savez_compressed(dst, **out)
        print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


def main():
    if "chamfer6d" in sys.argv[1:]:
        return chamfer6d()
    pv = build_ref.load_ref("pvcnn")
    ch = build_ref.load_ref("chamfer")
    em = build_ref.load_ref("emd")
    assert pv and ch and em, "oracle/_ref is not built"
    out = {}
    for name, kw in cases.VOXEL.items():
        feat, coords, res = cases.voxel_case(**kw)
        c = torch.from_numpy(coords).cuda()  # modules/voxelization.py:16-24, all torch ops on the GPU
        nc = c - c.mean(2, keepdim=True)
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + 0) + 0.5
        nc = torch.clamp(nc * res, 0, res - 1)
        vox = torch.round(nc).to(torch.int32)
        o, ind, cnt = pv.avg_voxelize_forward(t(feat), vox.contiguous(), res)
        gy = torch.from_numpy(np.random.default_rng(kw["seed"] + 100).standard_normal(o.shape).astype(np.float32)).cuda()
        gx = pv.avg_voxelize_backward(gy, ind, cnt)
        out[name + "/voxcoords"] = n(vox).astype(np.int8 if res <= 127 else np.int16)
        out[name + "/ind"] = n(ind)
        out[name + "/cnt_nz"] = np.stack(np.nonzero(n(cnt)) + (n(cnt)[np.nonzero(n(cnt))],)).astype(np.int32)
        out[name + "/out_sum"] = n(o.double().sum(dim=2)).astype(np.float64)
        out[name + "/out_sample"] = n(o[:, :2, :]).astype(np.float32)[..., ::7]
        out[name + "/gx_sample"] = n(gx[:, :2, :]).astype(np.float32)
    for name, kw in cases.DEVOX.items():
        coords, vox, gy, res = cases.devox_case(**kw)
        o, inds, wgts = pv.trilinear_devoxelize_forward(res, True, t(coords), t(vox))
        gx = pv.trilinear_devoxelize_backward(t(gy), inds, wgts, res)
        out[name + "/inds"] = n(inds)
        out[name + "/wgts"] = n(wgts)
        out[name + "/outs"] = n(o)
        out[name + "/gx"] = n(gx)
    for name, kw in cases.CHAMFER.items():
        a, c = cases.cloud_pair(**kw)
        b, nn_, m = a.shape[0], a.shape[1], c.shape[1]
        d1 = torch.zeros(b, nn_).cuda(); d2 = torch.zeros(b, m).cuda()
        i1 = torch.zeros(b, nn_, dtype=torch.int32).cuda(); i2 = torch.zeros(b, m, dtype=torch.int32).cuda()
        assert ch.forward(t(a), t(c), d1, d2, i1, i2) == 1
        g1 = torch.zeros(b, nn_, 3).cuda(); g2 = torch.zeros(b, m, 3).cuda()
        gd1 = torch.ones(b, nn_).cuda(); gd2 = (torch.arange(m).float() / m).repeat(b, 1).cuda()
        ch.backward(t(a), t(c), g1, g2, gd1, gd2, i1, i2)
        out[name + "/dist1"] = n(d1); out[name + "/dist2"] = n(d2)
        out[name + "/idx1"] = n(i1); out[name + "/idx2"] = n(i2)
        out[name + "/g1"] = n(g1); out[name + "/g2"] = n(g2)
    for name, kw in cases.EMD.items():
        a, c = cases.cloud_pair(**kw)
        match = em.approxmatch_forward(t(a), t(c))
        cost = em.matchcost_forward(t(a), t(c), match)
        gc = torch.linspace(0.5, 1.5, a.shape[0]).cuda()
        g1, g2 = em.matchcost_backward(gc, t(a), t(c), match)
        out[name + "/cost"] = n(cost)
        out[name + "/match_rowsum"] = n(match.sum(dim=2)); out[name + "/match_colsum"] = n(match.sum(dim=1))
        out[name + "/g1"] = n(g1); out[name + "/g2"] = n(g2)
    # EMD known answer of the reference's own test (test_emd_loss.py:7-10)
    p1 = np.array([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]], np.float32).repeat(3, 0)
    p2 = np.array([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]], np.float32).repeat(3, 0)
    mt = em.approxmatch_forward(t(p1), t(p2))
    out["emd_known/match"] = n(mt); out["emd_known/cost"] = n(em.matchcost_forward(t(p1), t(p2), mt))
    for name, kw in cases.POINTNET.items():
        pts, feat, cfeat = cases.pointnet_case(**kw)
        m = kw["m"]
        fidx = pv.furthest_point_sampling(t(pts), m)
        centers = pv.gather_features_forward(t(pts), fidx)
        bq = pv.ball_query(centers, t(pts), cases.BALL["radius"], cases.BALL["u"])
        grp = pv.grouping_forward(t(feat), bq)
        ggy = torch.from_numpy(np.random.default_rng(kw["seed"] + 7).standard_normal(tuple(grp.shape)).astype(np.float32)).cuda()
        ggx = pv.grouping_backward(ggy, bq, kw["n"])
        o3, i3, w3 = pv.three_nearest_neighbors_interpolate_forward(t(pts), centers, t(cfeat))
        g3 = pv.three_nearest_neighbors_interpolate_backward(t(feat), i3, w3, m)
        out[name + "/fps"] = n(fidx); out[name + "/ball"] = n(bq)
        out[name + "/grp_bwd"] = n(ggx)
        out[name + "/nn_idx"] = n(i3); out[name + "/nn_w"] = n(w3); out[name + "/nn_out"] = n(o3); out[name + "/nn_bwd"] = n(g3)
    for dst in (os.path.join(HERE, "ref_b200.npz"), os.path.join(ROOT, "gpurun_out", "ref_b200.npz")):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        np.savez_compressed(dst, **out)
        print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    main()
