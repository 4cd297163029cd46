"""Every kernel once on tiny shapes — meant to run under `compute-sanitizer --tool memcheck` on the GPU box
(tests/ only: uses nothing but the product API).  Prints OK at the end."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-flow-matching_b200"))
from pcfm import _lib, backend as be  # noqa: E402
from pcfm.nn import PVConv  # noqa: E402


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    for (b, c, n, r) in [(2, 8, 1001, 7), (1, 16, 3000, 16), (2, 4, 2048, 32), (1, 2, 60000, 4)]:
        feat = rn(b, c, n)
        vc = torch.randint(0, r, (b, 3, n), device="cuda", dtype=torch.int32, generator=g)
        for variant in (0, 1, 2):
            _lib.set_option("vox_fwd", variant)
            out, ind, cnt = be.pvcnn.avg_voxelize_forward(feat, vc, r)
        _lib.set_option("vox_fwd", 0)
        be.pvcnn.avg_voxelize_backward(rn(b, c, r ** 3), ind, cnt)
        fc = torch.rand(b, 3, n, device="cuda", generator=g) * (r - 1)
        fc[:, :, :5] = r - 1
        for stage in (0, 1):
            _lib.set_option("stage", stage)
            outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(r, True, fc, out)
            be.pvcnn.trilinear_devoxelize_forward(r, False, fc, out)
        _lib.set_option("stage", 0)
        for variant in (0, 1):
            _lib.set_option("devox_bwd", variant)
            be.pvcnn.trilinear_devoxelize_backward(rn(b, c, n), inds, wgts, r)
        _lib.set_option("devox_bwd", 0)
        for stage in (0, 1):
            _lib.set_option("stage", stage)
            be.pvcnn.trilinear_devoxelize_backward_coords(rn(b, c, n), fc, r)
        _lib.set_option("stage", 0)
    # plan reuse (second call on the same coords tensors -> *_planned entries), incl. heavy cells split into pieces,
    # and the optional channel-lane devoxelize kernel (C % 32 == 0, r <= 10)
    be.plan_memo.clear()
    for (b, c, n, r) in [(2, 32, 3001, 8), (1, 64, 5000, 5), (1, 32, 700, 10), (1, 32, 4000, 40)]:
        vc = torch.randint(0, min(r, 3), (b, 3, n), device="cuda", dtype=torch.int32, generator=g)   # dense: heavy cells
        fc = (torch.rand(b, 3, n, device="cuda", generator=g) * (r - 1)).contiguous()
        h0 = be.plan_memo.hits
        for _ in range(2):
            out, ind, cnt = be.pvcnn.avg_voxelize_forward(rn(b, c, n), vc, r)
            be.pvcnn.trilinear_devoxelize_backward_coords(rn(b, c, n), fc, r)
        assert be.plan_memo.hits - h0 == 2
        for opt in (1, 2, 0):
            _lib.set_option("devox_cl", opt)
            be.pvcnn.trilinear_devoxelize_forward(r, True, fc, out)
    # plane too large for shared memory (r = 40 -> 64000 floats): L2 path
    be.pvcnn.trilinear_devoxelize_forward(40, True, torch.rand(1, 3, 500, device="cuda") * 39, rn(1, 2, 64000))
    pts = torch.rand(2, 3, 700, device="cuda", generator=g)
    fidx = be.pvcnn.furthest_point_sampling(pts, 100)
    ctr = be.pvcnn.gather_features_forward(pts, fidx)
    bq = be.pvcnn.ball_query(ctr, pts, 0.2, 9)
    grp = be.pvcnn.grouping_forward(rn(2, 5, 700), bq)
    be.pvcnn.grouping_backward(torch.randn_like(grp), bq, 700)
    be.pvcnn.gather_features_backward(rn(2, 3, 100), fidx, 700)
    o3, i3, w3 = be.pvcnn.three_nearest_neighbors_interpolate_forward(pts, ctr, rn(2, 5, 100))
    be.pvcnn.three_nearest_neighbors_interpolate_backward(rn(2, 5, 700), i3, w3, 100)
    be.pvcnn.furthest_point_sampling(torch.rand(1, 3, 15000, device="cuda", generator=g), 20)  # global-memory path
    for (b, n, m) in [(2, 100, 77), (1, 2500, 2100), (3, 5, 1)]:
        a, c = rn(b, n, 3), rn(b, m, 3)
        d1 = torch.zeros(b, n, device="cuda"); d2 = torch.zeros(b, m, device="cuda")
        i1 = torch.zeros(b, n, dtype=torch.int32, device="cuda"); i2 = torch.zeros(b, m, dtype=torch.int32, device="cuda")
        assert be.chamfer.forward(a, c, d1, d2, i1, i2) == 1
        g1 = torch.zeros(b, n, 3, device="cuda"); g2 = torch.zeros(b, m, 3, device="cuda")
        assert be.chamfer.backward(a, c, g1, g2, torch.ones_like(d1), torch.ones_like(d2), i1, i2) == 1
    for (b, n, m) in [(2, 64, 64), (1, 300, 150), (1, 1100, 2500)]:
        a, c = torch.rand(b, n, 3, device="cuda", generator=g), torch.rand(b, m, 3, device="cuda", generator=g)
        match = be.emd.approxmatch_forward(a, c)
        cost = be.emd.matchcost_forward(a, c, match)
        be.emd.matchcost_backward(torch.ones_like(cost), a, c, match)
    s = torch.rand(3, 300, 3, device="cuda", generator=g); q = torch.rand(2, 257, 3, device="cuda", generator=g)
    be.chamfer3d_allpairs(s, q)
    be.emd_allpairs(s, q)
    be.emd_allpairs(torch.rand(2, 2048, 3, device="cuda", generator=g), torch.rand(1, 2048, 3, device="cuda", generator=g))
    m = PVConv(8, 8, 3, 8, with_se=True, normalize=True, eps=1e-6).cuda()
    f = rn(2, 8, 300).requires_grad_(True)
    out, _ = m((f, rn(2, 3, 300)))
    out.sum().backward()
    torch.cuda.synchronize()
    print("OK", _lib.launch_count(), "launches")


if __name__ == "__main__":
    main()
