"""Property tests (hypothesis) of the CPU oracle against brute-force numpy statements of the reference's tie-break rules
(SURVEY.md 7.4.2): Chamfer lowest index, ball query first-U-in-order with first-hit fill, FPS (max dist, smallest k mod
512, smallest k), 3-NN earliest centre; plus algebraic properties of voxelize / devoxelize."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import cpu_ref as o

SET = settings(max_examples=40, deadline=None)
f32 = np.float32


def lattice(draw, n, lo=0, hi=4):
    """Points on a coarse lattice: distances are exact in fp32 and ties are everywhere."""
    vals = draw(st.lists(st.integers(lo, hi), min_size=3 * n, max_size=3 * n))
    return (np.array(vals, dtype=f32).reshape(n, 3) * f32(0.25))


@SET
@given(st.data())
def test_chamfer_lowest_index(data):
    n = data.draw(st.integers(1, 40)); m = data.draw(st.integers(1, 40))
    a = lattice(data.draw, n)[None]; c = lattice(data.draw, m)[None]
    d1, d2, i1, i2 = o.chamfer_forward(a, c)
    D = ((a[0][:, None, :] - c[0][None, :, :]) ** 2).sum(-1)          # exact on the lattice
    assert np.array_equal(i1[0], D.argmin(1)) and np.array_equal(i2[0], D.argmin(0))   # numpy argmin = first minimum
    assert np.array_equal(d1[0], D.min(1)) and np.array_equal(d2[0], D.min(0))


@SET
@given(st.data())
def test_ball_query_first_u_in_index_order(data):
    n = data.draw(st.integers(1, 60)); m = data.draw(st.integers(1, 10)); u = data.draw(st.integers(1, 8))
    pts = lattice(data.draw, n).T[None].copy(); ctr = lattice(data.draw, m).T[None].copy()
    radius = data.draw(st.sampled_from([0.2, 0.3, 0.6, 5.0]))
    got = o.ball_query(ctr, pts, radius, u)[0]
    D = ((ctr[0].T[:, None, :] - pts[0].T[None, :, :]) ** 2).sum(-1)
    for j in range(m):
        hits = np.nonzero(D[j] < f32(radius) * f32(radius))[0][:u]
        want = np.zeros(u, np.int32)
        if len(hits):
            want[:] = hits[0]
            want[: len(hits)] = hits
        assert np.array_equal(got[j], want)


@SET
@given(st.data())
def test_fps_tie_break_rule(data):
    n = data.draw(st.integers(1, 1300)); m = data.draw(st.integers(1, 6))
    r = np.random.default_rng(data.draw(st.integers(0, 10 ** 6)))
    pts = (r.integers(0, 3, size=(3, n)) * 0.5).astype(f32)[None]
    got = o.furthest_point_sampling(pts, m)[0]
    dist = np.full(n, 1e38, f32)
    old, want = 0, [0]
    for _ in range(1, m):
        d = ((pts[0] - pts[0][:, old:old + 1]) ** 2).sum(0).astype(f32)
        dist = np.minimum(dist, d)
        best = dist.max()
        cand = np.nonzero(dist == best)[0]
        old = int(min(cand, key=lambda k: (k % 512, k)))            # sampling.cu:141,154
        want.append(old)
    assert got.tolist() == want


@SET
@given(st.data())
def test_three_nn_earliest_centre_and_weights(data):
    n = data.draw(st.integers(1, 30)); m = data.draw(st.integers(3, 20))
    pts = lattice(data.draw, n).T[None].copy(); ctr = lattice(data.draw, m, 0, 6).T[None].copy()
    feat = np.arange(2 * m, dtype=f32).reshape(1, 2, m)
    out, idx, w = o.three_nn_interpolate_forward(pts, ctr, feat)
    D = ((pts[0].T[:, None, :] - ctr[0].T[None, :, :]) ** 2).sum(-1)
    order = np.argsort(D, axis=1, kind="stable")[:, :3]              # stable sort = earliest centre on ties
    assert np.array_equal(idx[0].T, order)
    assert np.allclose(w.sum(1), 1.0, atol=1e-6) and (w >= 0).all()


@SET
@given(st.integers(2, 9), st.integers(1, 200), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_voxelize_is_a_mean_and_backward_is_its_transpose(res, n, c, seed):
    r = np.random.default_rng(seed)
    feat = r.standard_normal((1, c, n)).astype(f32)
    vox = r.integers(0, res, size=(1, 3, n)).astype(np.int32)
    out, ind, cnt = o.avg_voxelize_forward(feat, vox, res)
    assert cnt.sum() == n and np.array_equal(ind[0], vox[0, 0] * res * res + vox[0, 1] * res + vox[0, 2])
    ones, _, _ = o.avg_voxelize_forward(np.ones_like(feat), vox, res)
    assert np.allclose(ones[0, 0][cnt[0] > 0], 1.0, atol=1e-6) and (ones[0, 0][cnt[0] == 0] == 0).all()
    gy = r.standard_normal(out.shape).astype(f32)
    gx = o.avg_voxelize_backward(gy, ind, cnt)
    assert abs(float((out.astype(np.float64) * gy).sum()) - float((feat.astype(np.float64) * gx).sum())) < 1e-3   # <Ax,y> = <x,A^T y>


@SET
@given(st.integers(2, 9), st.integers(1, 100), st.integers(0, 10 ** 6))
def test_devoxelize_partition_of_unity_and_adjoint(res, n, seed):
    r = np.random.default_rng(seed)
    coords = (r.random((1, 3, n)) * (res - 1)).astype(f32)
    coords[0, :, ::4] = np.round(coords[0, :, ::4])
    vox = r.standard_normal((1, 2, res ** 3)).astype(f32)
    outs, inds, wgts = o.trilinear_devoxelize_forward(res, coords, vox)
    assert np.allclose(wgts.sum(1), 1.0, atol=1e-6) and inds.min() >= 0 and inds.max() < res ** 3
    const, _, _ = o.trilinear_devoxelize_forward(res, coords, np.full_like(vox, 3.0))
    assert np.allclose(const, 3.0, atol=1e-5)
    gy = r.standard_normal(outs.shape).astype(f32)
    gx = o.trilinear_devoxelize_backward(gy, inds, wgts, res)
    assert abs(float((outs.astype(np.float64) * gy).sum()) - float((vox.astype(np.float64) * gx).sum())) < 1e-3


def test_emd_properties():
    r = np.random.default_rng(3)
    a = r.random((2, 64, 3)).astype(f32); c = r.random((2, 64, 3)).astype(f32)
    match = o.emd_approxmatch(a, c)
    assert np.allclose(match.sum(1), 1.0, atol=2e-3) and np.allclose(match.sum(2), 1.0, atol=2e-3)   # doubly stochastic
    same = o.earth_mover_distance(a, a)
    assert (same < 1e-3).all() and (o.earth_mover_distance(a, c) > same).all()
