"""CPU tests of the oracle (oracle/pcfm_oracle.c via oracle/cpu_ref.py): pinned against
  (1) golden vectors produced by the UNMODIFIED reference CUDA kernels on a B200 (tests/golden/ref_b200.npz,
      generator tests/golden/make_golden.py),
  (2) the reference's own known answers (unit_test.py, test_emd_loss.py, README CD(x,x)=0),
  (3) an independent stock-PyTorch statement of the same ops (oracle/torch_cpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402
from util import assert_close, assert_equal  # noqa: E402

from oracle import cpu_ref as o  # noqa: E402
from oracle import torch_cpu as tc  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "ref_b200.npz")


@pytest.fixture(scope="module")
def golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/ref_b200.npz not generated yet (python tests/golden/make_golden.py on a GPU box)")
    return np.load(GOLDEN)


# ---- (1) golden vectors of the real reference -------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.VOXEL))
def test_golden_voxelize(golden, name):
    kw = cases.VOXEL[name]
    feat, _, res = cases.voxel_case(**kw)
    vox = golden[name + "/voxcoords"].astype(np.int32)
    out, ind, cnt = o.avg_voxelize_forward(feat, vox, res)
    assert_equal(ind, golden[name + "/ind"], "ind")
    nz = golden[name + "/cnt_nz"]
    want = np.zeros_like(cnt)
    want[nz[0], nz[1]] = nz[2]
    assert_equal(cnt, want, "cnt")
    assert_close(out.astype(np.float64).sum(2), golden[name + "/out_sum"], "sum over voxels", rtol=1e-5, atol_frac=1e-5)
    assert_close(out[:, :2, ::7], golden[name + "/out_sample"], "out sample")
    gy = np.random.default_rng(kw["seed"] + 100).standard_normal(out.shape).astype(np.float32)
    assert_close(o.avg_voxelize_backward(gy, ind, cnt)[:, :2], golden[name + "/gx_sample"], "grad sample")


@pytest.mark.parametrize("name", sorted(cases.DEVOX))
def test_golden_devoxelize(golden, name):
    coords, vox, gy, res = cases.devox_case(**cases.DEVOX[name])
    outs, inds, wgts = o.trilinear_devoxelize_forward(res, coords, vox)
    assert_equal(inds, golden[name + "/inds"], "inds")
    assert_equal(wgts, golden[name + "/wgts"], "wgts")
    assert_close(outs, golden[name + "/outs"], "outs")
    assert_close(o.trilinear_devoxelize_backward(gy, inds, wgts, res), golden[name + "/gx"], "grad")


@pytest.mark.parametrize("name", sorted(cases.CHAMFER))
def test_golden_chamfer(golden, name):
    a, c = cases.cloud_pair(**cases.CHAMFER[name])
    d1, d2, i1, i2 = o.chamfer_forward(a, c)
    assert_equal(i1, golden[name + "/idx1"], "idx1"); assert_equal(i2, golden[name + "/idx2"], "idx2")
    assert_equal(d1, golden[name + "/dist1"], "dist1"); assert_equal(d2, golden[name + "/dist2"], "dist2")
    m = c.shape[1]
    gd1 = np.ones_like(d1); gd2 = np.tile(np.arange(m, dtype=np.float32) / m, (a.shape[0], 1))
    g1, g2 = o.chamfer_backward(a, c, gd1, gd2, i1, i2)
    assert_close(g1, golden[name + "/g1"], "g1"); assert_close(g2, golden[name + "/g2"], "g2")


GOLDEN6 = os.path.join(HERE, "golden", "ref_b200_chamfer6d.npz")


@pytest.mark.parametrize("name", sorted(cases.CHAMFER6D))
def test_golden_chamfer_6d(name):
    """The compiled reference chamfer_6D (oracle/_ref/chamfer6d) run on a B200: bit-exact distances and indices."""
    if not os.path.exists(GOLDEN6):
        pytest.skip("tests/golden/ref_b200_chamfer6d.npz not generated (python tests/golden/make_golden.py chamfer6d)")
    golden = np.load(GOLDEN6)
    a, c = cases.cloud_pair(**cases.CHAMFER6D[name])
    d1, d2, i1, i2 = o.chamfer_nd_forward(a, c)
    assert_equal(i1, golden[name + "/idx1"], "idx1"); assert_equal(i2, golden[name + "/idx2"], "idx2")
    assert_equal(d1, golden[name + "/dist1"], "dist1"); assert_equal(d2, golden[name + "/dist2"], "dist2")
    m = c.shape[1]
    gd1 = np.ones_like(d1); gd2 = np.tile(np.arange(m, dtype=np.float32) / m, (a.shape[0], 1))
    g1, g2 = o.chamfer_nd_backward(a, c, gd1, gd2, i1, i2)
    assert_close(g1, golden[name + "/g1"], "g1"); assert_close(g2, golden[name + "/g2"], "g2")


def test_chamfer_nd_reduces_to_3d_and_matches_float64():
    """With the extra coordinates equal in both clouds the 6-D operator is the 3-D one bit for bit; on random 5-D /
    2-D clouds it matches the float64 pairwise method (unit_test.py:14-35 criterion)."""
    a, c = cases.cloud_pair(seed=71, b=2, n=400, m=300, kind="lattice")
    a6 = np.concatenate([a, np.full_like(a, 0.5)], 2); c6 = np.concatenate([c, np.full_like(c, 0.5)], 2)
    for x, y in zip(o.chamfer_nd_forward(a6, c6), o.chamfer_forward(a, c)):
        assert_equal(x, y, "6D with constant tail vs 3D")
    for dim in (2, 5):
        p, q = cases.cloud_pair(seed=72 + dim, b=3, n=150, m=260, kind="rand", dim=dim)
        d1, d2, i1, i2 = o.chamfer_nd_forward(p, q)
        D = ((p[:, :, None, :].astype(np.float64) - q[:, None, :, :]) ** 2).sum(-1)
        assert float(((d1 - D.min(2)) ** 2).mean() + ((d2 - D.min(1)) ** 2).mean()) < 1e-8
        assert_equal(i1, D.argmin(2).astype(np.int32), "idx1"); assert_equal(i2, D.argmin(1).astype(np.int32), "idx2")


@pytest.mark.parametrize("name", sorted(cases.EMD))
def test_golden_emd(golden, name):
    a, c = cases.cloud_pair(**cases.EMD[name])
    match = o.emd_approxmatch(a, c)
    assert_close(o.emd_matchcost(a, c, match), golden[name + "/cost"], "cost", rtol=2e-5)
    assert_close(match.sum(2), golden[name + "/match_rowsum"], "row sums", rtol=2e-5, atol_frac=2e-6)
    assert_close(match.sum(1), golden[name + "/match_colsum"], "col sums", rtol=2e-5, atol_frac=2e-6)
    gc = np.linspace(0.5, 1.5, a.shape[0]).astype(np.float32)
    g1, g2 = o.emd_matchcost_backward(gc, a, c, match)
    # The reference is compiled with --use_fast_math (approximate ex2 / division, FTZ), which a CPU cannot restate
    # bit for bit; cost and marginals agree to 2e-5, the gradient is a cancelling sum over `match` and agrees to
    # 5e-4 of max|g|.  (The CUDA kernels, which use the same intrinsics, are held to 1e-5 in test_gpu_parity.py.)
    assert_close(g1, golden[name + "/g1"], "g1", rtol=5e-4, atol_frac=5e-4)
    assert_close(g2, golden[name + "/g2"], "g2", rtol=5e-4, atol_frac=5e-4)


@pytest.mark.parametrize("name", sorted(cases.POINTNET))
def test_golden_pointnet_ops(golden, name):
    kw = cases.POINTNET[name]
    pts, feat, cfeat = cases.pointnet_case(**kw)
    fidx = o.furthest_point_sampling(pts, kw["m"])
    assert_equal(fidx, golden[name + "/fps"], "fps")
    centers = o.gather_forward(pts, fidx)
    bq = o.ball_query(centers, pts, cases.BALL["radius"], cases.BALL["u"])
    assert_equal(bq, golden[name + "/ball"], "ball query")
    grp = o.grouping_forward(feat, bq)
    ggy = np.random.default_rng(kw["seed"] + 7).standard_normal(grp.shape).astype(np.float32)
    assert_close(o.grouping_backward(ggy, bq, kw["n"]), golden[name + "/grp_bwd"], "grouping bwd")
    out, idx, w = o.three_nn_interpolate_forward(pts, centers, cfeat)
    assert_equal(idx, golden[name + "/nn_idx"], "3nn idx")
    assert_equal(w, golden[name + "/nn_w"], "3nn weights")
    assert_close(out, golden[name + "/nn_out"], "3nn out")
    assert_close(o.three_nn_interpolate_backward(feat, idx, w, kw["m"]), golden[name + "/nn_bwd"], "3nn bwd")


def test_golden_emd_known_answer_of_reference(golden):
    assert np.allclose(golden["emd_known/cost"] / 2, 0.355, atol=1e-5)


# ---- (2) the reference's own known answers -----------------------------------------------------------------
def test_emd_known_answer():
    """test_emd_loss.py:7-43 -> d = 0.355 for each of the three identical batch elements."""
    p1 = np.array([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]], np.float32).repeat(3, 0)
    p2 = np.array([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]], np.float32).repeat(3, 0)
    d = o.earth_mover_distance(p1, p2)
    assert np.allclose(d, 0.355, atol=1e-5)
    assert abs(d[0] / 2 + d[1] * 2 + d[2] / 3 - 2.011667 / 2) < 1e-4   # gt_dist = 2.011667 is the un-normalised cost; d = cost / N
    match = o.emd_approxmatch(p1, p2)
    assert np.allclose(match[0], [[0, 1], [1, 0]], atol=1e-4)


def test_chamfer_same_cloud_is_zero():
    """README.md:116-134."""
    x = np.random.default_rng(0).standard_normal((2, 2048, 3)).astype(np.float32)
    d1, d2, i1, i2 = o.chamfer_forward(x, x)
    assert d1.mean() + d2.mean() == 0.0
    assert_equal(i1, np.tile(np.arange(2048, dtype=np.int32), (2, 1)), "idx")


def test_chamfer_vs_float64_method():
    """unit_test.py:14-35 with chamfer_python.py:18-43's float64 method: MSE < 1e-8, identical indices."""
    p1 = torch.rand(4, 100, 3, generator=torch.Generator().manual_seed(3))
    p2 = torch.rand(4, 200, 3, generator=torch.Generator().manual_seed(4))
    d1, d2, i1, i2 = o.chamfer_forward(p1.numpy(), p2.numpy())
    x, y = p1.double(), p2.double()
    P = x.pow(2).sum(2)[:, :, None] + y.pow(2).sum(2)[:, None, :] - 2 * torch.bmm(x, y.transpose(2, 1))
    md1, mi1 = P.min(2); md2, mi2 = P.min(1)
    assert ((d1 - md1.numpy()) ** 2).mean() + ((d2 - md2.numpy()) ** 2).mean() < 1e-8
    assert np.array_equal(i1, mi1.numpy()) and np.array_equal(i2, mi2.numpy())


def test_chamfer_lowest_index_tie_break():
    a = np.zeros((1, 3, 3), np.float32)
    c = np.array([[[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]]], np.float32)   # four equidistant points
    _, _, i1, i2 = o.chamfer_forward(a, c)
    assert (i1 == 0).all() and (i2 == 0).all()


# ---- (3) independent PyTorch statement ------------------------------------------------------------------------
@pytest.mark.parametrize("res,n", [(8, 500), (32, 2048)])
def test_voxelize_devoxelize_vs_torch(res, n):
    r = np.random.default_rng(5)
    feat = r.standard_normal((2, 6, n)).astype(np.float32)
    coords = r.standard_normal((2, 3, n)).astype(np.float32)
    out, nc, vox, ind, cnt = o.voxelization_module(feat, coords, res, True, 0.0)
    t_nc = tc.normalize_coords(torch.from_numpy(coords), res)
    t_vox = tc.avg_voxelize(torch.from_numpy(feat), torch.round(t_nc).long(), res)
    assert_close(out, t_vox.reshape(2, 6, -1).numpy(), "avg_voxelize vs scatter_add", rtol=1e-5, atol_frac=1e-5)
    outs, inds, wgts = o.trilinear_devoxelize_forward(res, nc, out)
    t_out = tc.trilinear_devoxelize(t_vox, t_nc, res)
    assert_close(outs, t_out.numpy(), "devoxelize vs grid_sample", rtol=2e-5, atol_frac=2e-5)
    assert np.allclose(wgts.sum(1), 1.0, atol=1e-6)
    assert inds.min() >= 0 and inds.max() < res ** 3
    # backward of the two ops vs autograd through the torch statement
    gy = r.standard_normal(outs.shape).astype(np.float32)
    f = torch.from_numpy(feat).requires_grad_(True)
    tv = tc.avg_voxelize(f, torch.round(t_nc).long(), res)
    to = tc.trilinear_devoxelize(tv, t_nc, res)
    to.backward(torch.from_numpy(gy))
    gvox = o.trilinear_devoxelize_backward(gy, inds, wgts, res)
    gfeat = o.avg_voxelize_backward(gvox, ind, cnt)
    assert_close(gfeat, f.grad.numpy(), "feature gradient vs autograd", rtol=5e-5, atol_frac=5e-5)


def test_devox_edge_cases():
    """frac == 0 never reads the next plane (trilinear_devox.cu:64-75): coords exactly r-1 stay in range."""
    res = 4
    coords = np.array([[[3.0, 0.0, 1.5], [3.0, 0.0, 2.0], [3.0, 0.0, 0.25]]], np.float32)
    vox = np.arange(res ** 3, dtype=np.float32)[None, None]
    outs, inds, wgts = o.trilinear_devoxelize_forward(res, coords, vox)
    assert inds.max() == res ** 3 - 1 and inds.min() == 0
    assert outs[0, 0, 0] == res ** 3 - 1 and outs[0, 0, 1] == 0


def test_ball_query_semantics():
    pts = np.array([[[0, .1, .2, .3, 5], [0, 0, 0, 0, 5], [0, 0, 0, 0, 5]]], np.float32)
    ctr = np.array([[[0.0, 9.0], [0, 9], [0, 9]]], np.float32)
    bq = o.ball_query(ctr, pts, 0.25, 4)
    assert bq[0, 0].tolist() == [0, 1, 2, 0]      # first hit pre-fills the unused slots
    assert bq[0, 1].tolist() == [0, 0, 0, 0]      # no hit: zero row


def test_fps_tie_break_differs_from_plain_argmax():
    """sampling.cu:141,154: ties go to the smallest (k mod 512), then the smallest k."""
    n = 1100
    pts = np.zeros((1, 3, n), np.float32)
    pts[0, 0, 600] = 1.0     # slot 88
    pts[0, 0, 520] = -1.0    # slot 8: same distance from point 0, higher priority than 600 despite... lower k too
    pts[0, 0, 1030] = 1.0    # slot 6 (1030 % 512): wins over both although its index is the largest
    idx = o.furthest_point_sampling(pts, 2)
    assert idx[0].tolist() == [0, 1030]


def test_three_nn_weights_sum_to_one():
    r = np.random.default_rng(3)
    pts = r.random((1, 3, 50)).astype(np.float32); ctr = r.random((1, 3, 9)).astype(np.float32)
    out, idx, w = o.three_nn_interpolate_forward(pts, ctr, r.random((1, 2, 9)).astype(np.float32))
    assert np.allclose(w.sum(1), 1, atol=1e-6) and idx.max() < 9
