"""GPU parity tests: the sm_100a kernels, called through the C ABI (pcfm.backend is a ctypes shim over
include/pcfm_b200.h), against the CPU oracle on the same seeded inputs, and — when oracle/_ref was built —
against the UNMODIFIED reference CUDA kernels running on the same GPU.

Bars (north_star): integer outputs (voxel / neighbour / arg-min indices, counts) bit-exact;
fp32 values within 1e-5 (tests/util.py states the exact rule)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
from util import assert_close, assert_equal, rng  # noqa: E402

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(x):
    return x.detach().cpu().numpy()


@pytest.fixture(scope="module")
def be():
    from pcfm import backend
    return backend


@pytest.fixture(scope="module")
def oracle():
    from oracle import cpu_ref
    cpu_ref.build()
    return cpu_ref


@pytest.fixture(scope="module")
def ref():
    """The compiled reference extensions (None if oracle/_ref is absent)."""
    from oracle import build_ref
    out = {}
    for k in ("pvcnn", "chamfer", "emd", "chamfer6d"):
        try:
            out[k] = build_ref.load_ref(k)
        except Exception as e:  # noqa: BLE001
            print(f"oracle/_ref {k} not loadable: {e}")
            out[k] = None
    return out


def gpu_voxel_coords(coords, res):
    """modules/voxelization.py:16-24 on the GPU (what both the reference module and pcfm.nn.Voxelization run)."""
    c = cu(coords)
    nc = c - c.mean(2, keepdim=True)
    nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + 0) + 0.5
    nc = torch.clamp(nc * res, 0, res - 1)
    return torch.round(nc).to(torch.int32).contiguous(), nc


# ------------------------------------------------------------------------------------------------ voxelize
VOX_SHAPES = [
    dict(seed=1, b=2, c=64, n=2048, res=32),                    # BASELINE config 1
    dict(seed=2, b=2, c=3, n=1001, res=7),                      # odd sizes: no 16-byte alignment anywhere
    dict(seed=3, b=1, c=4, n=4096, res=4, kind="clustered"),    # hundreds of points per voxel
    dict(seed=4, b=1, c=1, n=1, res=2),
    dict(seed=5, b=3, c=16, n=20000, res=32),                   # config-2 point count
    dict(seed=6, b=2, c=32, n=20000, res=8),                    # long segments (config-2 stage 3)
]


@pytest.mark.parametrize("kw", VOX_SHAPES, ids=lambda k: f"b{k['b']}c{k['c']}n{k['n']}r{k['res']}")
@pytest.mark.parametrize("variant", [0, 1], ids=["csr", "atomic"])
def test_avg_voxelize(be, oracle, ref, kw, variant):
    from pcfm import _lib
    feat, coords, res = cases.voxel_case(**kw)
    vox, _ = gpu_voxel_coords(coords, res)
    _lib.set_option("vox_fwd", variant)
    try:
        out, ind, cnt = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
    finally:
        _lib.set_option("vox_fwd", 0)
    o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, host(vox), res)
    assert_equal(host(ind), o_ind, "ind")
    assert_equal(host(cnt), o_cnt, "cnt")
    if variant == 0 and o_cnt.max() <= 16:  # kSortMax
        # voxels with <= 16 points (cellplan.cuh kLightMax) are summed in ascending point order, like the oracle
        assert_equal(host(out), o_out, "out (ascending-order sum is bit-identical to the oracle)")
    else:
        assert_close(host(out), o_out, "out")
    gy = rng(kw["seed"] + 100).standard_normal(o_out.shape).astype(np.float32)
    gx = be.pvcnn.avg_voxelize_backward(cu(gy), ind, cnt)
    assert_equal(host(gx), oracle.avg_voxelize_backward(gy, o_ind, o_cnt), "grad_x")
    if ref["pvcnn"] is not None and variant == 0:
        r_out, r_ind, r_cnt = ref["pvcnn"].avg_voxelize_forward(cu(feat), vox, res)
        assert_equal(host(ind), host(r_ind), "ind vs reference")
        assert_equal(host(cnt), host(r_cnt), "cnt vs reference")
        assert_close(host(out), host(r_out), "out vs reference")
        assert_close(host(gx), host(ref["pvcnn"].avg_voxelize_backward(cu(gy), r_ind, r_cnt)), "grad vs reference")


@pytest.mark.parametrize("n", [700, 20000])
def test_avg_voxelize_all_points_in_few_voxels(be, oracle, n):
    """Degenerate clouds: every point in ONE voxel (a single heavy cell, listed as pieces of <= 512 records that
    several warps reduce and combine with RED.ADD) and every point in one of two voxels."""
    r0 = rng(400 + n)
    feat = r0.standard_normal((2, 5, n)).astype(np.float32)
    vox = np.zeros((2, 3, n), np.int32)
    vox[0] = 3                                   # cloud 0: one voxel
    vox[1, 2] = r0.integers(0, 2, size=n)        # cloud 1: two voxels
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(cu(feat), cu(vox), 8)
    o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, vox, 8)
    assert_equal(host(ind), o_ind, "ind"); assert_equal(host(cnt), o_cnt, "cnt")
    assert_close(host(out), o_out, "out", atol_frac=1e-5)


def test_avg_voxelize_stage_variants_agree(be):
    from pcfm import _lib
    feat, coords, res = cases.voxel_case(seed=9, b=2, c=8, n=4000, res=16)
    vox, _ = gpu_voxel_coords(coords, res)
    a = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
    _lib.set_option("stage", 1)
    try:
        b = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
    finally:
        _lib.set_option("stage", 0)
    light = (host(a[2]) <= 16)[:, None, :]          # cells of <= 16 points are summed in ascending point order
    assert_equal(np.where(light, host(a[0]), 0), np.where(light, host(b[0]), 0), "TMA-bulk vs plain staging")
    assert_close(host(a[0]), host(b[0]), "TMA-bulk vs plain staging (cells of > 16 points: arrival order)")


@pytest.mark.parametrize("res,n", [(32, 20000), (16, 5000), (8, 20000), (3, 50)])
def test_cell_plan_fused_and_multi_kernel_builds_agree(be, res, n):
    """The cell plan is built by one fused kernel (per-cell table in shared memory) when the grid fits, else by
    separate kernels; both must give the same voxelize / devoxelize-backward results."""
    from pcfm import _lib
    feat, coords, _ = cases.voxel_case(seed=90 + res, b=2, c=6, n=n, res=res)
    vox, nc = gpu_voxel_coords(coords, res)
    gy = cu(rng(91).standard_normal((2, 6, n)).astype(np.float32))
    a = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
    nc = nc.contiguous()
    ga = be.pvcnn.trilinear_devoxelize_backward_coords(gy, nc, res)
    _lib.set_option("plan", 1)
    try:
        b = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
        gb = be.pvcnn.trilinear_devoxelize_backward_coords(gy, nc, res)
    finally:
        _lib.set_option("plan", 0)
    assert_equal(host(a[1]), host(b[1]), "ind")
    assert_equal(host(a[2]), host(b[2]), "cnt")
    assert_close(host(a[0]), host(b[0]), "voxelize out")
    assert_close(host(ga), host(gb), "devoxelize backward")


# ------------------------------------------------------------------------------------------------ devoxelize
DEVOX_SHAPES = [
    dict(seed=11, b=2, c=64, n=2048, res=32),
    dict(seed=12, b=1, c=3, n=333, res=5),
    dict(seed=13, b=2, c=16, n=20000, res=16),
    dict(seed=14, b=2, c=16, n=20000, res=8),
    dict(seed=15, b=1, c=5, n=7, res=2),
    dict(seed=18, b=2, c=304, n=3001, res=8),     # enough rows for the two-channels-per-CTA backward (unaligned rows)
    dict(seed=19, b=2, c=300, n=6000, res=16),    # same, 16-byte aligned rows (TMA staging)
    dict(seed=20, b=2, c=608, n=1000, res=8),     # four channels per CTA: channel-interleaved forward kernel <4,8>
    dict(seed=21, b=2, c=592, n=1500, res=16),    # <4,16>
    dict(seed=22, b=2, c=592, n=700, res=12),     # <4,runtime r>
    dict(seed=27, b=1, c=32, n=2500, res=10),     # channel-lane kernel under devox_cl=2 (one CTA per SM)
    dict(seed=23, b=2, c=32, n=20000, res=8),     # C % 32 == 0, r <= 8: also run through the optional channel-lane kernel
    dict(seed=24, b=1, c=64, n=1001, res=5),      # same, odd point count and grid
    dict(seed=25, b=2, c=32, n=5, res=2),         # same, fewer points than one warp step
    dict(seed=26, b=3, c=96, n=4099, res=7),
]


@pytest.mark.parametrize("kw", DEVOX_SHAPES, ids=lambda k: f"b{k['b']}c{k['c']}n{k['n']}r{k['res']}")
def test_trilinear_devoxelize(be, oracle, ref, kw):
    from pcfm import _lib
    coords, vox, gy, res = cases.devox_case(**kw)
    outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(coords), cu(vox))
    o_outs, o_inds, o_wgts = oracle.trilinear_devoxelize_forward(res, coords, vox)
    assert_equal(host(inds), o_inds, "inds")
    assert_equal(host(wgts), o_wgts, "wgts (products of exact factors: bit-exact)")
    assert_close(host(outs), o_outs, "outs")
    e_outs, e_inds, e_wgts = be.pvcnn.trilinear_devoxelize_forward(res, False, cu(coords), cu(vox))
    assert tuple(e_inds.shape) == (1,) and tuple(e_wgts.shape) == (1,)     # trilinear_devox.cpp:45-49
    assert_equal(host(e_outs), host(outs), "eval-mode outs")
    # the plane-staging point-lane kernels (the forward without a cell plan: devox_fwd=1, and what few-channel calls
    # use anyway) and the optional channel-lane kernel (devox_cl: 1 = up to r = 8, 2 = up to r = 10)
    _lib.set_option("devox_fwd", 1)
    try:
        q_outs, q_inds, q_wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(coords), cu(vox))
        assert_equal(host(q_inds), o_inds, "inds (plane-staging kernels)")
        assert_equal(host(q_wgts), o_wgts, "wgts (plane-staging kernels)")
        assert_close(host(q_outs), o_outs, "outs (plane-staging kernels)")
        assert_close(host(q_outs), host(outs), "outs: cell walk vs plane-staging kernels")
        if kw["c"] % 32 == 0 and res <= 10:
            for opt in (1, 2):
                _lib.set_option("devox_cl", opt)
                try:
                    p_outs, p_inds, p_wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(coords), cu(vox))
                finally:
                    _lib.set_option("devox_cl", 0)
                assert_equal(host(p_inds), o_inds, f"inds (devox_cl={opt})")
                assert_equal(host(p_wgts), o_wgts, f"wgts (devox_cl={opt})")
                assert_equal(host(p_outs), host(q_outs), f"outs (devox_cl={opt}): same FMA order in the point / channel-lane kernels")
    finally:
        _lib.set_option("devox_fwd", 0)
    # a second call on the SAME coords tensor reuses the memoised plan (pcfm_trilinear_devoxelize_forward_cells, planned)
    cc, vv = cu(coords), cu(vox)
    first = be.pvcnn.trilinear_devoxelize_forward(res, True, cc, vv)
    again = be.pvcnn.trilinear_devoxelize_forward(res, True, cc, vv)
    assert_equal(host(again[0]), host(first[0]), "outs on the reused plan")
    assert_equal(host(again[1]), o_inds, "inds on the reused plan")
    o_gx = oracle.trilinear_devoxelize_backward(gy, o_inds, o_wgts, res)
    for variant in (0, 1):
        _lib.set_option("devox_bwd", variant)
        try:
            gx = be.pvcnn.trilinear_devoxelize_backward(cu(gy), inds, wgts, res)
        finally:
            _lib.set_option("devox_bwd", 0)
        # thousands of signed contributions per voxel at r=8: summation-order noise is bounded by eps*sum|terms|, so
        # the absolute term of the elementwise rule is 1e-5*max|b| here (the normwise 1e-5 bar is unchanged)
        assert_close(host(gx), o_gx, f"grad_x variant {variant}", atol_frac=1e-5)
    # the autograd path: backward from the coordinates (cell plan, every point visited once)
    gxc = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)
    assert_close(host(gxc), o_gx, "grad_x from coords", atol_frac=1e-5)
    assert_equal(host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)) != 0, o_gx != 0,
                 "sparsity pattern of grad_x")
    if ref["pvcnn"] is not None:
        r_outs, r_inds, r_wgts = ref["pvcnn"].trilinear_devoxelize_forward(res, True, cu(coords), cu(vox))
        assert_equal(host(inds), host(r_inds), "inds vs reference")
        assert_equal(host(wgts), host(r_wgts), "wgts vs reference")
        assert_close(host(outs), host(r_outs), "outs vs reference")
        # the reference adds ~40 signed terms per voxel with global atomics in arrival order: its own run-to-run spread
        # needs the same 1e-5 absolute term as the oracle comparison above (1 failure in 25 runs at 1e-6)
        assert_close(host(be.pvcnn.trilinear_devoxelize_backward(cu(gy), inds, wgts, res)),
                     host(ref["pvcnn"].trilinear_devoxelize_backward(cu(gy), r_inds, r_wgts, res)), "grad vs reference",
                     atol_frac=1e-5)


@pytest.mark.parametrize("kw", [DEVOX_SHAPES[0], DEVOX_SHAPES[2], DEVOX_SHAPES[3], DEVOX_SHAPES[6], DEVOX_SHAPES[14]],
                         ids=lambda k: f"b{k['b']}c{k['c']}n{k['n']}r{k['res']}")
def test_trilinear_walk_variants(be, oracle, kw):
    """Devoxelize backward, how light cells reach the plane (pcfm_set_option("bwd_phases")): 0 = shared-memory atomics
    (default), 2 = deferred conflict-free corner phases (sums parked in the output row; fixed summation order) — each
    against the oracle and each other."""
    from pcfm import _lib
    coords, vox, gy, res = cases.devox_case(**kw)
    o_outs, o_inds, o_wgts = oracle.trilinear_devoxelize_forward(res, coords, vox)
    o_gx = oracle.trilinear_devoxelize_backward(gy, o_inds, o_wgts, res)
    got = {}
    for name, opts in (("default", {}), ("bwd_deferred", {"bwd_phases": 2})):
        for k, v in opts.items():
            _lib.set_option(k, v)
        try:
            cc = cu(coords)
            outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cc, cu(vox))
            gx = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cc, res)          # plan left by the forward
            gx2 = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)  # fresh plan (PDL behind the build)
        finally:
            for k in opts:
                _lib.set_option(k, 0)
        assert_equal(host(inds), o_inds, f"inds ({name})")
        assert_close(host(outs), o_outs, f"outs ({name})")
        assert_close(host(gx), o_gx, f"grad_x ({name})", atol_frac=1e-5)
        assert_close(host(gx2), o_gx, f"grad_x, fresh plan ({name})", atol_frac=1e-5)
        got[name] = (host(outs), host(gx))
    for name in ("bwd_deferred",):
        assert_close(got[name][0], got["default"][0], f"outs: {name} vs default")
        assert_close(got[name][1], got["default"][1], f"grad_x: {name} vs default", atol_frac=1e-5)
    # the deferred phases add a voxel's light contributions in a fixed order: bit-identical run to run when no heavy
    # cell (> 64 points, added with atomics) is involved
    if kw["n"] <= 2048:
        _lib.set_option("bwd_phases", 2)
        try:
            cc = cu(coords)
            a = host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cc, res))
            b = host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cc, res))
        finally:
            _lib.set_option("bwd_phases", 0)
        if int(np.bincount(np.floor(coords[0]).astype(np.int64).T.dot([res * res, res, 1])).max()) <= 64:
            assert_equal(a, b, "deferred-phase backward, run to run")


def test_devox_backward_generic_is_deterministic(be):
    """Generic (inds/wgts) entry: sorted segmented sums, run-to-run bit-identical when no voxel receives more than
    64 contributions (csr.cu kMaxSortedSeg: those segments are put in ascending edge order)."""
    coords, vox, gy, res = cases.devox_case(seed=16, b=2, c=8, n=2000, res=16)
    coords = (rng(17).random(coords.shape) * (res - 1)).astype(np.float32)   # no piles on the corners
    _, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(coords), cu(vox))
    a = host(be.pvcnn.trilinear_devoxelize_backward(cu(gy), inds, wgts, res))
    c = host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res))
    for _ in range(3):
        assert_equal(host(be.pvcnn.trilinear_devoxelize_backward(cu(gy), inds, wgts, res)), a, "run-to-run")
        # cell-plan entry: atomic-free, but cells are visited in arrival order => reproducible to summation order only
        assert_close(host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)), c,
                     "run-to-run (coords)", rtol=1e-6, atol_frac=1e-6)
    assert_close(c, a, "coords entry vs generic entry", rtol=1e-6, atol_frac=1e-6)


def test_pvconv_channels_last_voxel_grids(be):
    """PVConv.channels_last: Conv3d / BatchNorm3d / SE see the voxel grid in channels_last_3d; the block's
    output and the parameter gradients are those of the NCDHW run (TF32 off: only cuDNN's algorithm choice differs)."""
    from pcfm.nn import PVConv
    torch.manual_seed(0)
    blk = PVConv(32, 32, 3, 16, with_se=True, normalize=True, eps=1e-6).cuda().train()
    g = torch.Generator(device="cuda").manual_seed(5)
    f = torch.randn(2, 32, 3000, device="cuda", generator=g, requires_grad=True)
    c = torch.randn(2, 3, 3000, device="cuda", generator=g)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        res = []
        for cl in (False, True):
            PVConv.channels_last = cl
            blk.zero_grad(set_to_none=True)
            f.grad = None
            for m in blk.modules():
                if isinstance(m, torch.nn.BatchNorm3d) or isinstance(m, torch.nn.BatchNorm1d):
                    m.reset_running_stats()
            out, _ = blk((f, c))
            out.square().mean().backward()
            res.append((out.detach().clone(), f.grad.clone(), blk.voxel_layers[0].weight.grad.clone()))
    finally:
        PVConv.channels_last = False
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for a, b, what in zip(res[1], res[0], ("output", "feature gradient", "conv weight gradient")):
        err = float((a - b).abs().max() / b.abs().max())
        assert err <= 1e-4, f"channels_last {what}: {err:.2e}"


# ------------------------------------------------------------------------------------------------ bench shapes
@pytest.mark.parametrize("c,res", [(128, 32), (256, 16), (256, 8)], ids=lambda v: str(v))
def test_pvconv_chain_at_the_bench_shapes_against_the_reference_kernels(be, ref, c, res):
    """BASELINE config 2 exactly (B=8, N=20000, Gaussian clouds through the reference normalisation): the four
    primitives of a PVConv block against the UNMODIFIED reference kernels (oracle/_ref) on the same GPU.  Integer
    outputs bit-exact.  Values: the reference adds with order-free atomics, so it is first compared with ITSELF (two
    runs); this library must agree with it within max(1e-5, that spread) normwise — and the gather-type ops (avg_voxelize
    backward, devoxelize forward), where nothing is order-dependent, within 1e-6 / 1e-5 as usual."""
    R = ref["pvcnn"]
    if R is None:
        pytest.skip("oracle/_ref/pvcnn not built")
    b, n = 8, 20000
    g = torch.Generator(device="cuda").manual_seed(1000 + res)
    feat = torch.randn(b, c, n, device="cuda", generator=g)
    coords = torch.randn(b, 3, n, device="cuda", generator=g)
    nc = coords - coords.mean(2, keepdim=True)
    nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + 1e-6) + 0.5
    nc = torch.clamp(nc * res, 0, res - 1).contiguous()
    vox = torch.round(nc).to(torch.int32).contiguous()
    grid = torch.randn(b, c, res ** 3, device="cuda", generator=g)
    gy = torch.randn(b, c, n, device="cuda", generator=g)

    def spread(x, y):
        return float((x - y).abs().max() / y.abs().max())

    P = be.pvcnn
    out, ind, cnt = P.avg_voxelize_forward(feat, vox, res, nc)
    r_out, r_ind, r_cnt = R.avg_voxelize_forward(feat, vox, res)
    r_out2 = R.avg_voxelize_forward(feat, vox, res)[0]
    assert torch.equal(ind, r_ind) and torch.equal(cnt, r_cnt), "ind / cnt"
    noise = spread(r_out2, r_out)
    assert spread(out, r_out) <= max(1e-5, 2 * noise), f"avg_voxelize: {spread(out, r_out):.2e} (reference vs itself {noise:.2e})"
    report = {"vox_fwd": (spread(out, r_out), noise)}

    outs, inds, wgts = P.trilinear_devoxelize_forward(res, True, nc, grid)
    r_outs, r_inds, r_wgts = R.trilinear_devoxelize_forward(res, True, nc, grid)
    assert torch.equal(inds, r_inds) and torch.equal(wgts, r_wgts), "inds / wgts"
    assert_close(host(outs), host(r_outs), "devoxelize forward")
    report["devox_fwd"] = (spread(outs, r_outs), 0.0)

    gg = P.trilinear_devoxelize_backward_coords(gy, nc, res)
    r_gg = R.trilinear_devoxelize_backward(gy, r_inds, r_wgts, res)
    r_gg2 = R.trilinear_devoxelize_backward(gy, r_inds, r_wgts, res)
    noise = spread(r_gg2, r_gg)
    assert spread(gg, r_gg) <= max(1e-5, 2 * noise), f"devoxelize backward: {spread(gg, r_gg):.2e} (reference vs itself {noise:.2e})"
    report["devox_bwd"] = (spread(gg, r_gg), noise)
    assert torch.equal(gg != 0, r_gg != 0), "sparsity pattern of the devoxelize gradient"

    gf = P.avg_voxelize_backward(grid, ind, cnt)
    r_gf = R.avg_voxelize_backward(grid, r_ind, r_cnt)
    assert_equal(host(gf), host(r_gf), "avg_voxelize backward (a gather: bit-exact)")
    print(f"bench shape C={c} r={res}: " + ", ".join(f"{k} {a:.1e} (ref vs ref {e:.1e})" for k, (a, e) in report.items()))


@pytest.mark.parametrize("c,res", [(128, 32), (256, 16), (256, 8)], ids=lambda v: str(v))
def test_pvconv_primitives_are_adjoint_pairs_at_the_bench_shapes(be, c, res):
    """Size-independent property at BASELINE config 2 (B=8, N=20000): each backward is the ADJOINT of its linear forward,
    <forward(x), y> = <x, backward(y)> for random x, y (inner products in float64), and both forwards are linear.  Needs
    no oracle and no reference, so it holds the full-size kernels to account on any box."""
    b, n = 8, 20000
    g = torch.Generator(device="cuda").manual_seed(2000 + res)
    coords = torch.randn(b, 3, n, device="cuda", generator=g)
    nc = coords - coords.mean(2, keepdim=True)
    nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + 1e-6) + 0.5
    nc = torch.clamp(nc * res, 0, res - 1).contiguous()
    vox = torch.round(nc).to(torch.int32).contiguous()
    feat = torch.randn(b, c, n, device="cuda", generator=g)
    feat2 = torch.randn(b, c, n, device="cuda", generator=g)
    grid = torch.randn(b, c, res ** 3, device="cuda", generator=g)
    gy = torch.randn(b, c, n, device="cuda", generator=g)
    P = be.pvcnn

    def dot(x, y):
        return float((x.double() * y.double()).sum())

    def norm(x):
        return float(x.double().square().sum().sqrt())

    # avg_voxelize: forward F -> V (mean per voxel), backward V -> F
    out, ind, cnt = P.avg_voxelize_forward(feat, vox, res, nc)
    back = P.avg_voxelize_backward(grid, ind, cnt)
    # (the inner products of random fields are sums of ~1e7 signed terms: the identity is held relative to the product
    # of the norms, Cauchy-Schwarz's bound on either side)
    lhs, rhs = dot(out, grid), dot(feat, back)
    assert abs(lhs - rhs) <= 1e-6 * norm(out) * norm(grid), f"avg_voxelize adjoint identity: {lhs} vs {rhs}"
    out2 = P.avg_voxelize_forward(feat2, vox, res, nc)[0]
    out12 = P.avg_voxelize_forward(feat + 0.5 * feat2, vox, res, nc)[0]
    err = float((out12 - (out + 0.5 * out2)).abs().max() / out12.abs().max())
    assert err <= 1e-5, f"avg_voxelize linearity: {err:.2e}"
    assert int(cnt.sum()) == b * n, "every point is counted once"
    # trilinear devoxelize: forward V -> F, backward F -> V
    outs, inds, wgts = P.trilinear_devoxelize_forward(res, True, nc, grid)
    gx = P.trilinear_devoxelize_backward_coords(gy, nc, res)
    lhs, rhs = dot(outs, gy), dot(grid, gx)
    assert abs(lhs - rhs) <= 1e-6 * norm(outs) * norm(gy), f"trilinear devoxelize adjoint identity: {lhs} vs {rhs}"
    gx_generic = P.trilinear_devoxelize_backward(gy, inds, wgts, res)
    assert abs(dot(grid, gx_generic) - lhs) <= 1e-6 * norm(outs) * norm(gy), "adjoint identity, backward from inds / wgts"
    # the 8 weights of a point sum to 1: a constant grid interpolates to the same constant
    ones = P.trilinear_devoxelize_forward(res, False, nc, torch.ones(b, 2, res ** 3, device="cuda"))[0]
    assert float((ones - 1).abs().max()) <= 1e-6, "partition of unity"


# ------------------------------------------------------------------------------------------------ pointnet ops
PN_SHAPES = [
    dict(seed=41, b=2, c=5, n=1024, m=256, kind="rand"),
    dict(seed=42, b=2, c=3, n=600, m=150, kind="lattice"),      # exact ties everywhere
    dict(seed=43, b=1, c=2, n=20, m=7, kind="rand"),             # fewer points than a warp
    dict(seed=44, b=2, c=4, n=5000, m=1250, kind="rand"),
]


@pytest.mark.parametrize("n,m,kind", [(600, 150, "lattice"), (3000, 700, "rand"), (9, 9, "rand"), (3072, 300, "lattice"),
                                      (4100, 500, "rand"), (6144, 300, "lattice"), (20000, 700, "rand"),
                                      (40000, 64, "rand"), (70000, 40, "rand"), (131073, 12, "rand")])
def test_fps_cluster_and_single_cta_kernels_pick_the_reference_winners(be, oracle, n, m, kind):
    """All FPS kernels (one CTA per cloud below 3072 and above 131072 points; 8-CTA cluster with 1..8 points per thread
    in registers up to 65536 points, coordinates in shared memory up to 131072; slots exchanged by st.async pushes
    or, option "fps" = 2, by a cluster barrier and remote reads) must reproduce the reference's choice in every round,
    including its tie-break on lattice clouds."""
    from pcfm import _lib
    pts, _, _ = cases.pointnet_case(seed=45 + n % 7, b=2, c=1, n=n, m=m, kind=kind)
    want = oracle.furthest_point_sampling(pts, m)
    assert_equal(host(be.pvcnn.furthest_point_sampling(cu(pts), m)), want, "fps (default kernel)")
    for opt, what in ((1, "one CTA per cloud"), (2, "cluster kernel with a cluster barrier per round")):
        _lib.set_option("fps", opt)
        try:
            assert_equal(host(be.pvcnn.furthest_point_sampling(cu(pts), m)), want, f"fps ({what})")
        finally:
            _lib.set_option("fps", 0)


@pytest.mark.parametrize("b,c,n,m,u,kind", [(2, 5, 1000, 250, 32, "padded"), (1, 3, 4099, 300, 7, "random"),
                                            (2, 4, 20000, 1250, 32, "padded"), (1, 2, 30001, 100, 64, "padded"),
                                            (1, 2, 60000, 50, 16, "random"), (2, 3, 100000, 2000, 32, "padded"),
                                            (1, 1, 300001, 100, 8, "random"), (1, 1, 9, 1, 1, "random")])
def test_grouping_backward_forms(be, oracle, b, c, n, m, u, kind):
    """grouping backward as a scatter into a shared-memory row (default; rows beyond 55296 points are cut into slices
    with one CTA each, beyond 229376 points the sorted-segment path takes over) and as sorted-segment sums (pcfm_set_option("grouping_bwd", 1)) against the oracle.  "padded": neighbourhoods
    filled up with copies of their first neighbour, as the ball query leaves them (runs of equal targets)."""
    from pcfm import _lib
    g = rng(900 + n + u)
    idx = g.integers(0, n, size=(b, m, u)).astype(np.int32)
    if kind == "padded":
        found = g.integers(1, u + 1, size=(b, m))
        idx = np.where(np.arange(u)[None, None, :] < found[..., None], idx, idx[..., :1]).astype(np.int32)
    gy = g.standard_normal((b, c, m, u)).astype(np.float32)
    want = oracle.grouping_backward(gy, idx, n)
    for opt in (0, 1):
        _lib.set_option("grouping_bwd", opt)
        try:
            got = host(be.pvcnn.grouping_backward(cu(gy), cu(idx), n))
        finally:
            _lib.set_option("grouping_bwd", 0)
        assert_close(got, want, f"grouping backward (grouping_bwd={opt})", atol_frac=1e-6)
        assert_equal(got != 0, want != 0, f"sparsity pattern (grouping_bwd={opt})")


@pytest.mark.parametrize("b,c,n,m", [(2, 5, 3000, 700), (1, 3, 4099, 60000), (1, 2, 50000, 240000), (1, 1, 7, 3)])
def test_three_nn_backward_forms(be, oracle, b, c, n, m):
    """3-NN interpolate backward as a scatter into a shared-memory row of the m centres (default; more than 55296
    centres: slices with one CTA each; more than 229376: sorted segments) and as sorted-segment sums
    (pcfm_set_option("three_nn_bwd", 1)), for arbitrary (idx, weight) triples, against the oracle."""
    from pcfm import _lib
    g = rng(700 + n + m)
    idx = g.integers(0, m, size=(b, 3, n)).astype(np.int32)
    w = g.random((b, 3, n)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    gy = g.standard_normal((b, c, n)).astype(np.float32)
    want = oracle.three_nn_interpolate_backward(gy, idx, w, m)
    for opt in (0, 1):
        _lib.set_option("three_nn_bwd", opt)
        try:
            got = host(be.pvcnn.three_nearest_neighbors_interpolate_backward(cu(gy), cu(idx), cu(w), m))
        finally:
            _lib.set_option("three_nn_bwd", 0)
        assert_close(got, want, f"3-NN interpolate backward (three_nn_bwd={opt})", atol_frac=1e-6)
        assert_equal(got != 0, want != 0, f"sparsity pattern (three_nn_bwd={opt})")


@pytest.mark.parametrize("kw", PN_SHAPES, ids=lambda k: f"{k['kind']}n{k['n']}m{k['m']}")
def test_fps_ballquery_grouping_gather_3nn(be, oracle, ref, kw):
    pts, feat, cfeat = cases.pointnet_case(**kw)
    n, m = kw["n"], kw["m"]
    fidx = be.pvcnn.furthest_point_sampling(cu(pts), m)
    o_fidx = oracle.furthest_point_sampling(pts, m)
    assert_equal(host(fidx), o_fidx, "fps indices (reference tie-break)")
    centers = be.pvcnn.gather_features_forward(cu(pts), fidx)
    o_centers = oracle.gather_forward(pts, o_fidx)
    assert_equal(host(centers), o_centers, "gather")
    for radius, u in [(cases.BALL["radius"], cases.BALL["u"]), (0.05, 3), (10.0, 40)]:
        bq = be.pvcnn.ball_query(centers, cu(pts), radius, u)
        o_bq = oracle.ball_query(o_centers, pts, radius, u)
        assert_equal(host(bq), o_bq, f"ball_query r={radius} u={u}")
    bq = be.pvcnn.ball_query(centers, cu(pts), cases.BALL["radius"], cases.BALL["u"])
    grp = be.pvcnn.grouping_forward(cu(feat), bq)
    assert_equal(host(grp), oracle.grouping_forward(feat, o_bq if False else host(bq)), "grouping")
    ggy = rng(kw["seed"] + 7).standard_normal(tuple(grp.shape)).astype(np.float32)
    ggx = be.pvcnn.grouping_backward(cu(ggy), bq, n)
    assert_close(host(ggx), oracle.grouping_backward(ggy, host(bq), n), "grouping backward")
    gg = be.pvcnn.gather_features_backward(cu(cfeat), fidx, n)
    distinct = all(len(np.unique(row)) == m for row in o_fidx)   # FPS repeats index 0 once a lattice cloud is used up
    check = assert_equal if distinct else assert_close
    check(host(gg), oracle.gather_backward(cfeat, o_fidx, n), "gather backward (distinct indices: exact)")
    from pcfm import _lib
    _lib.set_option("gather_bwd", 1)          # sorted-segment variant
    try:
        check(host(be.pvcnn.gather_features_backward(cu(cfeat), fidx, n)), host(gg), "gather backward variants")
    finally:
        _lib.set_option("gather_bwd", 0)
    dup = fidx.clone(); dup[:, 1::2] = dup[:, ::2][:, :dup[:, 1::2].shape[1]]   # duplicated samples accumulate
    assert_close(host(be.pvcnn.gather_features_backward(cu(cfeat), dup, n)),
                 oracle.gather_backward(cfeat, host(dup), n), "gather backward with repeated indices")
    o3, i3, w3 = be.pvcnn.three_nearest_neighbors_interpolate_forward(cu(pts), centers, cu(cfeat))
    oo3, oi3, ow3 = oracle.three_nn_interpolate_forward(pts, o_centers, cfeat)
    assert_equal(host(i3), oi3, "3nn indices")
    assert_equal(host(w3), ow3, "3nn weights")
    assert_close(host(o3), oo3, "3nn interpolate")
    g3 = be.pvcnn.three_nearest_neighbors_interpolate_backward(cu(feat), i3, w3, m)
    assert_close(host(g3), oracle.three_nn_interpolate_backward(feat, oi3, ow3, m), "3nn backward")
    _lib.set_option("three_nn_bwd", 1)     # sorted-segment sums instead of the shared-memory scatter
    try:
        g3s = be.pvcnn.three_nearest_neighbors_interpolate_backward(cu(feat), i3, w3, m)
    finally:
        _lib.set_option("three_nn_bwd", 0)
    assert_close(host(g3s), oracle.three_nn_interpolate_backward(feat, oi3, ow3, m), "3nn backward (sorted segments)")
    if ref["pvcnn"] is not None:
        R = ref["pvcnn"]
        assert_equal(host(fidx), host(R.furthest_point_sampling(cu(pts), m)), "fps vs reference")
        assert_equal(host(bq), host(R.ball_query(centers, cu(pts), cases.BALL["radius"], cases.BALL["u"])), "ball vs reference")
        assert_equal(host(grp), host(R.grouping_forward(cu(feat), bq)), "grouping vs reference")
        assert_close(host(ggx), host(R.grouping_backward(cu(ggy), bq, n)), "grouping bwd vs reference", atol_frac=1e-5)
        r3, ri3, rw3 = R.three_nearest_neighbors_interpolate_forward(cu(pts), centers, cu(cfeat))
        assert_equal(host(i3), host(ri3), "3nn idx vs reference")
        assert_equal(host(w3), host(rw3), "3nn w vs reference")
        assert_close(host(o3), host(r3), "3nn out vs reference")
        assert_close(host(g3), host(R.three_nearest_neighbors_interpolate_backward(cu(feat), ri3, rw3, m)), "3nn bwd vs reference",
                     atol_frac=1e-5)


def test_ball_query_no_hit_rows_are_zero(be):
    pts = np.zeros((1, 3, 64), np.float32)
    ctr = np.full((1, 3, 5), 100.0, np.float32)
    assert host(be.pvcnn.ball_query(cu(ctr), cu(pts), 0.5, 8)).sum() == 0


# ------------------------------------------------------------------------------------------------ chamfer
CH_SHAPES = [
    dict(seed=21, b=4, n=100, m=200, kind="rand"),          # unit_test.py:15-16
    dict(seed=22, b=2, n=700, m=513, kind="lattice"),
    dict(seed=23, b=1, n=2500, m=2300, kind="randn"),       # crosses the 2048-point chunk
    dict(seed=24, b=2, n=2048, m=2048, kind="randn"),       # README check shape
    dict(seed=25, b=1, n=3, m=1, kind="randn"),
    dict(seed=26, b=32, n=2000, m=1000, kind="rand"),       # the reference's published timing shape
]


def run_chamfer(mod, a, c):
    b, n, m = a.shape[0], a.shape[1], c.shape[1]
    d1 = torch.zeros(b, n, device="cuda"); d2 = torch.zeros(b, m, device="cuda")
    i1 = torch.zeros(b, n, dtype=torch.int32, device="cuda"); i2 = torch.zeros(b, m, dtype=torch.int32, device="cuda")
    assert mod.forward(cu(a), cu(c), d1, d2, i1, i2) == 1
    return d1, d2, i1, i2


@pytest.mark.parametrize("kw", CH_SHAPES, ids=lambda k: f"{k['kind']}b{k['b']}n{k['n']}m{k['m']}")
def test_chamfer(be, oracle, ref, kw):
    a, c = cases.cloud_pair(**kw)
    d1, d2, i1, i2 = run_chamfer(be.chamfer, a, c)
    od1, od2, oi1, oi2 = oracle.chamfer_forward(a, c)
    assert_equal(host(i1), oi1, "idx1"); assert_equal(host(i2), oi2, "idx2")
    assert_equal(host(d1), od1, "dist1"); assert_equal(host(d2), od2, "dist2")
    b, n, m = a.shape[0], a.shape[1], c.shape[1]
    gd1 = rng(kw["seed"] + 1).standard_normal((b, n)).astype(np.float32)
    gd2 = rng(kw["seed"] + 2).standard_normal((b, m)).astype(np.float32)
    g1 = torch.zeros(b, n, 3, device="cuda"); g2 = torch.zeros(b, m, 3, device="cuda")
    assert be.chamfer.backward(cu(a), cu(c), g1, g2, cu(gd1), cu(gd2), i1, i2) == 1
    og1, og2 = oracle.chamfer_backward(a, c, gd1, gd2, oi1, oi2)
    assert_close(host(g1), og1, "gradxyz1"); assert_close(host(g2), og2, "gradxyz2")
    if ref["chamfer"] is not None:
        rd1, rd2, ri1, ri2 = run_chamfer(ref["chamfer"], a, c)
        assert_equal(host(i1), host(ri1), "idx1 vs reference"); assert_equal(host(i2), host(ri2), "idx2 vs reference")
        assert_equal(host(d1), host(rd1), "dist1 vs reference"); assert_equal(host(d2), host(rd2), "dist2 vs reference")
    # chamfer_once: 0 = the two directional passes (pcfm_chamfer3d_forward), 2 = the single pass whatever the size
    # (the default picks by size); same bits either way
    from pcfm import _lib
    for opt in (0, 2):
        _lib.set_option("chamfer_once", opt)
        try:
            t1, t2, j1, j2 = run_chamfer(be.chamfer, a, c)
        finally:
            _lib.set_option("chamfer_once", 1)
        assert_equal(host(j1), oi1, f"idx1 (chamfer_once={opt})"); assert_equal(host(j2), oi2, f"idx2 (chamfer_once={opt})")
        assert_equal(host(t1), od1, f"dist1 (chamfer_once={opt})"); assert_equal(host(t2), od2, f"dist2 (chamfer_once={opt})")


def test_chamfer_single_pass_large_lattice_ties(be, oracle):
    """B=2, N=M=6000 lattice clouds (every distance tied many times): several row tiles AND several column tiles meet
    in the 64-bit atomicMin keys; indices must still be the lowest among the ties."""
    r = rng(77)
    a = (r.integers(0, 9, size=(2, 6000, 3)) * 0.125).astype(np.float32)
    c = (r.integers(0, 9, size=(2, 5000, 3)) * 0.125).astype(np.float32)
    from pcfm import _lib
    _lib.set_option("chamfer_once", 2)
    try:
        d1, d2, i1, i2 = run_chamfer(be.chamfer, a, c)
    finally:
        _lib.set_option("chamfer_once", 1)
    od1, od2, oi1, oi2 = oracle.chamfer_forward(a, c)
    assert_equal(host(i1), oi1, "idx1"); assert_equal(host(i2), oi2, "idx2")
    assert_equal(host(d1), od1, "dist1"); assert_equal(host(d2), od2, "dist2")


def test_chamfer_same_cloud_is_zero(be):
    """README.md:116-134: CD(x, x) == 0 with idx[i] == i."""
    x = rng(0).standard_normal((2, 2048, 3)).astype(np.float32)
    d1, d2, i1, i2 = run_chamfer(be.chamfer, x, x)
    assert float(d1.mean() + d2.mean()) == 0.0
    assert_equal(host(i1), np.tile(np.arange(2048, dtype=np.int32), (2, 1)), "idx")


def test_chamfer_matches_float64_reference_method(be):
    """unit_test.py:14-35: MSE vs the float64 pairwise method < 1e-8 and identical indices."""
    p1 = torch.rand(4, 100, 3, generator=torch.Generator().manual_seed(3))
    p2 = torch.rand(4, 200, 3, generator=torch.Generator().manual_seed(4))
    d1, d2, i1, i2 = run_chamfer(be.chamfer, p1.numpy(), p2.numpy())
    x, y = p1.double(), p2.double()
    P = (x.pow(2).sum(2)[:, :, None] + y.pow(2).sum(2)[:, None, :] - 2 * torch.bmm(x, y.transpose(2, 1)))
    md1, mi1 = P.min(2); md2, mi2 = P.min(1)
    assert float(((d1.cpu() - md1.float()) ** 2).mean() + ((d2.cpu() - md2.float()) ** 2).mean()) < 1e-8
    assert torch.equal(i1.cpu().long(), mi1) and torch.equal(i2.cpu().long(), mi2)


def test_chamfer_autograd_wrapper(be, oracle):
    from pcfm.metrics import chamfer_3DDist
    a, c = cases.cloud_pair(seed=27, b=2, n=300, m=200)
    x1 = cu(a).requires_grad_(True); x2 = cu(c).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(x1, x2)
    (d1.sum() + 0.5 * d2.sum()).backward()
    od1, od2, oi1, oi2 = oracle.chamfer_forward(a, c)
    og1, og2 = oracle.chamfer_backward(a, c, np.ones_like(od1), 0.5 * np.ones_like(od2), oi1, oi2)
    assert_close(host(x1.grad), og1, "grad1"); assert_close(host(x2.grad), og2, "grad2")


# ------------------------------------------------------------------------------------------------ chamfer 2D / 5D / 6D
CHND_SHAPES = [dict(dim=6, **{k: v for k, v in kw.items() if k != "dim"}) for kw in cases.CHAMFER6D.values()] + [
    dict(seed=28, b=2, n=300, m=900, kind="lattice", dim=2),
    dict(seed=29, b=2, n=1200, m=333, kind="randn", dim=5),
    dict(seed=30, b=8, n=20000, m=2048, kind="rand", dim=6),     # two own points per thread
    dict(seed=30, b=1, n=1, m=1, kind="rand", dim=6),
]


@pytest.mark.parametrize("kw", CHND_SHAPES, ids=lambda k: f"d{k['dim']}{k['kind']}b{k['b']}n{k['n']}m{k['m']}")
def test_chamfer_nd(be, oracle, ref, kw):
    """chamfer_2D / 5D / 6D (ChamferDistancePytorch/chamfer6D/chamfer6D.cu): bit-exact distances and arg-mins vs the
    oracle and, for D = 6, vs the compiled reference extension."""
    a, c = cases.cloud_pair(**kw)
    dim = kw["dim"]
    mod = {2: be.chamfer_2d, 5: be.chamfer_5d, 6: be.chamfer_6d}[dim]
    d1, d2, i1, i2 = run_chamfer(mod, a, c)
    od1, od2, oi1, oi2 = oracle.chamfer_nd_forward(a, c)
    assert_equal(host(i1), oi1, "idx1"); assert_equal(host(i2), oi2, "idx2")
    assert_equal(host(d1), od1, "dist1"); assert_equal(host(d2), od2, "dist2")
    b, n, m = a.shape[0], a.shape[1], c.shape[1]
    gd1 = rng(kw["seed"] + 1).standard_normal((b, n)).astype(np.float32)
    gd2 = rng(kw["seed"] + 2).standard_normal((b, m)).astype(np.float32)
    g1 = torch.zeros(b, n, dim, device="cuda"); g2 = torch.zeros(b, m, dim, device="cuda")
    assert mod.backward(cu(a), cu(c), g1, g2, cu(gd1), cu(gd2), i1, i2) == 1
    og1, og2 = oracle.chamfer_nd_backward(a, c, gd1, gd2, oi1, oi2)
    assert_close(host(g1), og1, "gradxyz1"); assert_close(host(g2), og2, "gradxyz2")
    if dim == 6 and ref.get("chamfer6d") is not None and m % 8 not in (1, 3) and n % 8 not in (1, 3):
        rd1, rd2, ri1, ri2 = run_chamfer(ref["chamfer6d"], a, c)
        assert_equal(host(i1), host(ri1), "idx1 vs reference"); assert_equal(host(i2), host(ri2), "idx2 vs reference")
        assert_equal(host(d1), host(rd1), "dist1 vs reference"); assert_equal(host(d2), host(rd2), "dist2 vs reference")


def test_chamfer_6d_autograd_wrapper_and_3d_consistency(be, oracle):
    """dist_chamfer_6D.py:30-82 contract; with the last three coordinates equal in both clouds the 6-D distance is the
    3-D one."""
    from pcfm.metrics import chamfer_3DDist, chamfer_6DDist
    a, c = cases.cloud_pair(seed=27, b=2, n=300, m=200, dim=6)
    x1 = cu(a).requires_grad_(True); x2 = cu(c).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_6DDist()(x1, x2)
    (d1.sum() + 0.5 * d2.sum()).backward()
    od1, od2, oi1, oi2 = oracle.chamfer_nd_forward(a, c)
    og1, og2 = oracle.chamfer_nd_backward(a, c, np.ones_like(od1), 0.5 * np.ones_like(od2), oi1, oi2)
    assert_close(host(x1.grad), og1, "grad1"); assert_close(host(x2.grad), og2, "grad2")
    a[..., 3:] = 0.25; c[..., 3:] = 0.25
    e1, e2, j1, j2 = chamfer_6DDist()(cu(a), cu(c))
    f1, f2, k1, k2 = chamfer_3DDist()(cu(a[..., :3].copy()), cu(c[..., :3].copy()))
    assert torch.equal(e1, f1) and torch.equal(e2, f2) and torch.equal(j1, k1) and torch.equal(j2, k2)
    with pytest.raises(AssertionError):
        chamfer_6DDist()(cu(a[..., :3].copy()), cu(c[..., :3].copy()))


# ------------------------------------------------------------------------------------------------ EMD
EMD_SHAPES = [
    dict(seed=31, b=2, n=256, m=256, kind="rand"),
    dict(seed=32, b=2, n=300, m=150, kind="rand"),     # n = 2m: multiR = 2
    dict(seed=33, b=1, n=100, m=250, kind="rand"),     # multiL = 2 (integer division)
    dict(seed=34, b=2, n=1024, m=1024, kind="rand"),   # README check shape
]


@pytest.mark.parametrize("kw", EMD_SHAPES, ids=lambda k: f"b{k['b']}n{k['n']}m{k['m']}")
def test_emd(be, oracle, ref, kw):
    a, c = cases.cloud_pair(**kw)
    match = be.emd.approxmatch_forward(cu(a), cu(c))
    cost = be.emd.matchcost_forward(cu(a), cu(c), match)
    o_match = oracle.emd_approxmatch(a, c)
    o_cost = oracle.emd_matchcost(a, c, o_match)
    assert_close(host(cost), o_cost, "cost", rtol=1e-5)
    assert_close(host(match).sum(2), o_match.sum(2), "match row sums", rtol=2e-5)
    assert_close(host(match).sum(1), o_match.sum(1), "match col sums", rtol=2e-5)
    gc = np.linspace(0.5, 1.5, a.shape[0]).astype(np.float32)
    g1, g2 = be.emd.matchcost_backward(cu(gc), cu(a), cu(c), match)
    og1, og2 = oracle.emd_matchcost_backward(gc, a, c, host(match))
    assert_close(host(g1), og1, "grad1"); assert_close(host(g2), og2, "grad2")
    if ref["emd"] is not None:
        r_match = ref["emd"].approxmatch_forward(cu(a), cu(c))
        r_cost = ref["emd"].matchcost_forward(cu(a), cu(c), r_match)
        assert_close(host(cost), host(r_cost), "cost vs reference")
        assert_close(host(match), host(r_match), "match vs reference", rtol=2e-5, atol_frac=2e-6)
        rg1, rg2 = ref["emd"].matchcost_backward(cu(gc), cu(a), cu(c), r_match)
        assert_close(host(g1), host(rg1), "grad1 vs reference"); assert_close(host(g2), host(rg2), "grad2 vs reference")
    # the cluster kernel (8 CTAs per batch element, default from 256 points) deals ROWS to the CTAs and leaves every
    # row's summation order alone: bit-identical to the one-CTA kernel
    from pcfm import _lib
    _lib.set_option("emd_cluster", 0)
    try:
        one = be.emd.approxmatch_forward(cu(a), cu(c))
    finally:
        _lib.set_option("emd_cluster", 1)
    assert_equal(host(match), host(one), "match: cluster kernel vs one CTA per batch element")


def test_emd_cluster_kernel_shapes(be):
    """Cluster approxmatch at the README shape (B=2, N=1024), at B=8 x 2048 and on ragged sizes: bit-identical to the
    one-CTA kernel, doubly stochastic up to the auction's tolerance."""
    from pcfm import _lib
    for (b, n, m, seed) in [(2, 1024, 1024, 1), (8, 2048, 2048, 2), (3, 1500, 700, 3), (2, 300, 2100, 4), (1, 9000, 9000, 5)]:
        r = rng(seed)
        a = r.random((b, n, 3)).astype(np.float32); c = r.random((b, m, 3)).astype(np.float32)
        got = be.emd.approxmatch_forward(cu(a), cu(c))
        _lib.set_option("emd_cluster", 0)
        try:
            one = be.emd.approxmatch_forward(cu(a), cu(c))
        finally:
            _lib.set_option("emd_cluster", 1)
        assert torch.equal(got, one), f"B={b} N={n} M={m}"


def test_emd_known_answer(be):
    """test_emd_loss.py:7-43: d = 0.355 per batch element, loss = d0/2 + 2 d1 + d2/3 = 2.011667/2, analytic grads."""
    from pcfm.metrics import earth_mover_distance
    p1 = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]]).repeat(3, 1, 1).cuda().requires_grad_(True)
    p2 = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]]).repeat(3, 1, 1).cuda().requires_grad_(True)
    d = earth_mover_distance(p1, p2, transpose=False)
    assert torch.allclose(d.cpu(), torch.full((3,), 0.355), atol=1e-5)
    loss = d[0] / 2 + d[1] * 2 + d[2] / 3
    assert abs(float(loss) - 2.011667 / 2) < 1e-4   # gt_dist (2.011667) is the cost before the division by N=2
    loss.backward()
    q1 = p1.detach().clone().requires_grad_(True); q2 = p2.detach().clone().requires_grad_(True)
    gt = ((((q1[0, 0] - q2[0, 1]) ** 2).sum() + ((q1[0, 1] - q2[0, 0]) ** 2).sum()) / 2
          + (((q1[1, 0] - q2[1, 1]) ** 2).sum() + ((q1[1, 1] - q2[1, 0]) ** 2).sum()) * 2
          + (((q1[2, 0] - q2[2, 1]) ** 2).sum() + ((q1[2, 1] - q2[2, 0]) ** 2).sum()) / 3) / 2
    gt.backward()
    assert torch.allclose(p1.grad, q1.grad, atol=1e-4) and torch.allclose(p2.grad, q2.grad, atol=1e-4)


# ------------------------------------------------------------------------------------------------ all-pairs
def test_allpairs_chamfer_equals_operator(be, oracle):
    r = rng(51)
    s = (r.standard_normal((5, 700, 3)) * 0.3).astype(np.float32)
    q = (r.standard_normal((4, 900, 3)) * 0.3).astype(np.float32)
    s[1, :50] = s[1, 50:100]   # duplicated points
    cd = host(be.chamfer3d_allpairs(cu(s), cu(q)))
    for i in range(5):
        for j in range(4):
            d1, d2, _, _ = oracle.chamfer_forward(s[i:i + 1], q[j:j + 1])
            want = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
            assert abs(cd[i, j] - want) <= 1e-5 * abs(want), (i, j, cd[i, j], want)


def test_allpairs_chamfer_full_size_properties(be):
    """BASELINE config 4 point count (2048): symmetry cd(S,R) == cd(R,S)^T bit-exact, zero diagonal on cd(S,S)."""
    r = rng(52)
    s = (r.standard_normal((6, 2048, 3)) * 0.3).astype(np.float32)
    q = (r.standard_normal((7, 2048, 3)) * 0.3).astype(np.float32)
    a = host(be.chamfer3d_allpairs(cu(s), cu(q)))
    b = host(be.chamfer3d_allpairs(cu(q), cu(s)))
    assert_close(a, b.T, "cd(S,R) vs cd(R,S)^T", rtol=1e-6)
    assert np.all(np.diag(host(be.chamfer3d_allpairs(cu(s), cu(s)))) == 0.0)


def test_allpairs_emd_equals_operator(be, oracle):
    r = rng(53)
    s = r.random((3, 256, 3)).astype(np.float32)
    q = r.random((2, 256, 3)).astype(np.float32)
    got = host(be.emd_allpairs(cu(s), cu(q)))
    from pcfm.metrics import earth_mover_distance_nograd
    for i in range(3):
        for j in range(2):
            want_op = float(earth_mover_distance_nograd(cu(s[i:i + 1]), cu(q[j:j + 1]), transpose=False))
            want_or = float(oracle.earth_mover_distance(s[i:i + 1], q[j:j + 1])[0])
            assert abs(got[i, j] - want_op) <= 1e-5 * abs(want_op), (i, j, got[i, j], want_op)
            assert abs(got[i, j] - want_or) <= 2e-5 * abs(want_or), (i, j, got[i, j], want_or)


def test_metrics_harness_single_gpu(be):
    from pcfm import evaluation as ev
    r = rng(54)
    s = cu((r.standard_normal((12, 512, 3)) * 0.3).astype(np.float32))
    q = cu((r.standard_normal((10, 512, 3)) * 0.3).astype(np.float32))
    res = ev.compute_metrics(s, q, metrics=("cd", "emd"), block=5)
    for k in ("cd", "emd"):
        assert res[k]["mmd"] > 0 and 0 < res[k]["cov"] <= 1 and 0 <= res[k]["1nna"] <= 1
    full = ev.pairwise_matrix(s, q, "cd", block=100)
    blocked = ev.pairwise_matrix(s, q, "cd", block=5)
    assert torch.equal(full, blocked)


# ------------------------------------------------------------------------------------------------ modules
def test_pvconv_module_forward_backward_and_eval_contract(be, oracle):
    from pcfm.nn import PVConv
    torch.manual_seed(0)
    m = PVConv(16, 16, 3, 8, with_se=True, normalize=True, eps=1e-6).cuda()
    f = torch.randn(2, 16, 500, device="cuda", requires_grad=True)
    c = torch.randn(2, 3, 500, device="cuda")
    out, c2 = m((f, c))
    assert out.shape == (2, 16, 500) and c2 is c
    out.square().mean().backward()
    assert f.grad is not None and torch.isfinite(f.grad).all()
    # the voxel branch alone against the oracle
    vox, nc = m.voxelization(f.detach(), c)
    o_vox, o_nc, _, _, _ = oracle.voxelization_module(host(f), host(c), 8, True, 1e-6)
    assert_close(host(vox).reshape(2, 16, -1), o_vox, "Voxelization module")
    m.eval()
    f2 = f.detach().clone().requires_grad_(True)
    out2, _ = m((f2, c))
    with pytest.raises(Exception):   # nothing saved in eval mode (devoxelization.py:24-26)
        out2.sum().backward()


def _torch_voxel_coords(c, r, normalize, eps):
    """modules/voxelization.py:17-24 as the reference runs it (torch ops on the GPU)."""
    nc = c - c.mean(2, keepdim=True)
    if normalize:
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5
    else:
        nc = (nc + 1) / 2.0
    nc = torch.clamp(nc * r, 0, r - 1)
    return nc, torch.round(nc).to(torch.int32)


@pytest.mark.parametrize("r,normalize,eps", [(32, True, 0), (16, True, 1e-6), (8, False, 0), (7, True, 0)])
def test_voxelization_coords_kernel_is_bit_identical_to_the_torch_ops(be, r, normalize, eps):
    from pcfm.nn import Voxelization
    g = torch.Generator(device="cuda").manual_seed(70 + r)
    c = torch.randn(3, 3, 20001, device="cuda", generator=g) * (0.3 if not normalize else 1.0)
    # adversarial: coordinates that land exactly on .5 voxel boundaries and on the clamp limits
    c[0, :, :64] = 0.0
    c[1, 0, :200] = torch.linspace(-3, 3, 200, device="cuda")
    t_nc, t_vox = _torch_voxel_coords(c, r, normalize, eps)
    _, nc = Voxelization(r, normalize, eps)(torch.zeros(3, 1, 20001, device="cuda"), c)
    assert torch.equal(nc, t_nc), "norm_coords differ from the torch op sequence"
    centred = (c - c.mean(2, keepdim=True)).contiguous()
    scale = (centred.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps).reshape(-1) if normalize else None
    k_nc, k_vox = be.pvcnn.voxel_coords(centred, scale, r)
    assert torch.equal(k_nc, t_nc) and torch.equal(k_vox, t_vox)
    # half-integers round to even, like torch.round
    half = torch.tensor([[[0.5, 1.5, 2.5, 3.5, -0.5, 6.49999, 6.5, 1e9]] * 3], device="cuda")
    k_nc, k_vox = be.pvcnn.voxel_coords((half / r - 0.5).contiguous(), torch.ones(1, device="cuda"), r)
    e_nc = torch.clamp(((half / r - 0.5) / 1.0 + 0.5) * r, 0, r - 1)
    assert torch.equal(k_nc, e_nc) and torch.equal(k_vox, torch.round(e_nc).to(torch.int32))


def test_voxelization_coordinate_memo(be):
    """Layers that receive the SAME coords tensor reuse its voxel coordinates; an in-place write invalidates them."""
    from pcfm import _lib
    from pcfm.nn import Voxelization
    c = torch.randn(2, 3, 4000, device="cuda")
    f = torch.randn(2, 4, 4000, device="cuda")
    v16, v16b, v8 = Voxelization(16, True, 1e-6), Voxelization(16, True, 1e-6), Voxelization(8, True, 1e-6)
    _, nc_a = v16(f, c)
    n0 = _lib.launch_count()
    _, nc_b = v16b(f, c)                    # other instance, same configuration: memo hit
    assert nc_b is nc_a and _lib.launch_count() - n0 == 1      # voxelize kernel only: coords AND cell plan memoised
    _, nc_8 = v8(f, c)                      # other resolution: shares the centred coords / max norm
    assert torch.equal(nc_8, _torch_voxel_coords(c, 8, True, 1e-6)[0])
    c.mul_(0.5)                             # in-place write bumps the version counter
    _, nc_c = v16(f, c)
    assert nc_c is not nc_a and torch.equal(nc_c, _torch_voxel_coords(c, 16, True, 1e-6)[0])
    _, nc_d = v16(f, c.clone())             # a different tensor object never hits
    assert nc_d is not nc_c and torch.equal(nc_d, nc_c)
    Voxelization.cache = False
    try:
        _, nc_e = v16(f, c)
        _, nc_f = v16(f, c)
        assert nc_e is not nc_f and torch.equal(nc_e, nc_f)
    finally:
        Voxelization.cache = True


def test_inputs_are_not_mutated_and_streams_are_respected(be):
    feat, coords, res = cases.voxel_case(seed=61, b=2, c=8, n=3000, res=16)
    vox, _ = gpu_voxel_coords(coords, res)
    f = cu(feat); f0 = f.clone(); v0 = vox.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out, ind, cnt = be.pvcnn.avg_voxelize_forward(f, vox, res)
    s.synchronize()
    ref_out = be.pvcnn.avg_voxelize_forward(f, vox, res)[0]
    torch.cuda.synchronize()
    assert torch.allclose(out, ref_out, rtol=1e-5, atol=1e-6) and torch.equal(f, f0) and torch.equal(vox, v0)


# ------------------------------------------------------------------------------------------------ edge cases
def test_empty_and_degenerate_inputs(be, oracle):
    dev = "cuda"
    # no points: every voxel empty, devoxelize of nothing
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(torch.zeros(2, 4, 0, device=dev), torch.zeros(2, 3, 0, dtype=torch.int32, device=dev), 4)
    assert out.shape == (2, 4, 64) and float(out.abs().sum()) == 0 and int(cnt.sum()) == 0 and ind.shape == (2, 0)
    o, i, w = be.pvcnn.trilinear_devoxelize_forward(4, True, torch.zeros(2, 3, 0, device=dev), torch.randn(2, 4, 64, device=dev))
    assert o.shape == (2, 4, 0) and i.shape == (2, 8, 0)
    g = be.pvcnn.trilinear_devoxelize_backward_coords(torch.zeros(2, 4, 0, device=dev), torch.zeros(2, 3, 0, device=dev), 4)
    assert g.shape == (2, 4, 64) and float(g.abs().sum()) == 0
    # empty batch
    assert be.pvcnn.avg_voxelize_forward(torch.zeros(0, 4, 10, device=dev), torch.zeros(0, 3, 10, dtype=torch.int32, device=dev), 4)[0].shape == (0, 4, 64)
    # chamfer with an empty second cloud: the reference's loops do not run, outputs keep the caller's zeros, returns 1
    x = torch.randn(2, 5, 3, device=dev); y = torch.zeros(2, 0, 3, device=dev)
    d1 = torch.zeros(2, 5, device=dev); d2 = torch.zeros(2, 0, device=dev)
    i1 = torch.zeros(2, 5, dtype=torch.int32, device=dev); i2 = torch.zeros(2, 0, dtype=torch.int32, device=dev)
    assert be.chamfer.forward(x, y, d1, d2, i1, i2) == 1 and float(d1.abs().sum()) == 0
    # ball query over no points / grouping of no centres
    assert int(be.pvcnn.ball_query(torch.randn(1, 3, 4, device=dev), torch.zeros(1, 3, 0, device=dev), 1.0, 3).sum()) == 0
    assert be.pvcnn.grouping_forward(torch.randn(1, 2, 9, device=dev), torch.zeros(1, 0, 3, dtype=torch.int32, device=dev)).shape == (1, 2, 0, 3)


def test_all_points_in_one_voxel_and_long_rows(be, oracle):
    """One cell holding every point (a single 'heavy' cell) and rows too long for shared memory (N = 60000)."""
    r0 = rng(71)
    n, res = 60000, 4
    feat = r0.standard_normal((1, 3, n)).astype(np.float32)
    vox = np.full((1, 3, n), 2, np.int32)
    vox[0, :, : n // 2] = r0.integers(0, res, size=(3, n // 2))
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(cu(feat), cu(vox), res)
    o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, vox, res)
    assert_equal(host(ind), o_ind, "ind"); assert_equal(host(cnt), o_cnt, "cnt")
    assert_close(host(out), o_out, "out", atol_frac=1e-5)
    coords = (r0.random((1, 3, n)) * (res - 1)).astype(np.float32)
    coords[0, :, ::3] = 1.5
    gy = r0.standard_normal((1, 3, n)).astype(np.float32)
    grid = r0.standard_normal((1, 3, res ** 3)).astype(np.float32)
    outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(coords), cu(grid))
    oo, oi, ow = oracle.trilinear_devoxelize_forward(res, coords, grid)
    assert_equal(host(inds), oi, "inds"); assert_equal(host(wgts), ow, "wgts"); assert_close(host(outs), oo, "outs")
    ogx = oracle.trilinear_devoxelize_backward(gy, oi, ow, res)
    assert_close(host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)), ogx, "grad (coords)", atol_frac=1e-5)
    assert_close(host(be.pvcnn.trilinear_devoxelize_backward(cu(gy), inds, wgts, res)), ogx, "grad (generic)", atol_frac=1e-5)


def test_resolution_64_plane_outside_shared_memory(be, oracle):
    """BASELINE config 5 resolution r=64: a voxel plane is 1 MB, read through L2 instead of shared memory."""
    coords, vox, gy, res = cases.devox_case(seed=81, b=1, c=2, n=3000, res=64)
    outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(coords), cu(vox))
    oo, oi, ow = oracle.trilinear_devoxelize_forward(res, coords, vox)
    assert_equal(host(inds), oi, "inds"); assert_equal(host(wgts), ow, "wgts"); assert_close(host(outs), oo, "outs")
    from pcfm.functional import trilinear_devoxelize
    v = cu(vox).view(1, 2, res, res, res).requires_grad_(True)
    trilinear_devoxelize(v, cu(coords), res, True).backward(cu(gy))       # r > 37: cell plan + global atomics
    want = oracle.trilinear_devoxelize_backward(gy, oi, ow, res)
    assert_close(host(v.grad).reshape(1, 2, -1), want, "grad", atol_frac=1e-5)
    assert_close(host(be.pvcnn.trilinear_devoxelize_backward(cu(gy), inds, wgts, res)), want, "grad (generic entry)",
                 atol_frac=1e-5)
    # a cloud too big for shared memory on a grid too big for shared memory: rows read through L2 as well
    cb, vb, gb, _ = cases.devox_case(seed=84, b=1, c=2, n=70000, res=40)
    _, ib, wb = be.pvcnn.trilinear_devoxelize_forward(40, True, cu(cb), cu(vb))
    assert_close(host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(gb), cu(cb), 40)),
                 oracle.trilinear_devoxelize_backward(gb, host(ib), host(wb), 40), "grad r=40 N=70000", atol_frac=1e-5)
    feat, c2, _ = cases.voxel_case(seed=82, b=1, c=2, n=3000, res=64)
    vc, _ = gpu_voxel_coords(c2, 64)
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(cu(feat), vc, 64)
    o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, host(vc), 64)
    assert_equal(host(ind), o_ind, "ind"); assert_equal(host(cnt), o_cnt, "cnt"); assert_close(host(out), o_out, "out")
    gx = be.pvcnn.avg_voxelize_backward(cu(rng(83).standard_normal(o_out.shape).astype(np.float32)), ind, cnt)
    assert gx.shape == (1, 2, 3000)


def test_cuda_graph_capture_of_a_pvconv_primitive_chain(be):
    """Every launcher runs on the caller's stream and never synchronises, so a sampler step can be captured into a
    CUDA graph and replayed (the reference launches 14 of its 20 kernels on the legacy stream and cannot)."""
    feat, coords, res = cases.voxel_case(seed=91, b=2, c=8, n=3000, res=16)
    f = cu(feat)
    vox, nc = gpu_voxel_coords(coords, res)
    nc = nc.contiguous()
    eager = be.pvcnn.trilinear_devoxelize_forward(res, False, nc, be.pvcnn.avg_voxelize_forward(f, vox, res)[0])[0]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):   # warm-up on the side stream (one-time cudaFuncSetAttribute calls happen here)
            be.pvcnn.trilinear_devoxelize_forward(res, False, nc, be.pvcnn.avg_voxelize_forward(f, vox, res)[0])
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        grid = be.pvcnn.avg_voxelize_forward(f, vox, res)[0]
        out = be.pvcnn.trilinear_devoxelize_forward(res, False, nc, grid)[0]
    f.mul_(2.0)          # new input values, same buffers
    g.replay()
    torch.cuda.synchronize()
    assert_close(host(out), 2.0 * host(eager), "graph replay on updated inputs", rtol=1e-6, atol_frac=1e-6)


def test_cuda_graph_capture_of_a_pvconv_stack_in_eval_mode(be):
    """a sampler step (PVConv blocks in eval mode: Voxelization, our kernels with programmatic dependent
    launches, cuDNN Conv3d / BatchNorm between them) is captured once and replayed on new data in the same buffers."""
    from pcfm.nn import PVConv, SharedMLP
    torch.manual_seed(3)
    net = torch.nn.ModuleList([PVConv(8, 16, 3, 8, with_se=True, normalize=True, eps=1e-6), SharedMLP(16, [16]),
                               PVConv(16, 16, 3, 8, with_se=False, normalize=True, eps=1e-6)]).cuda().eval()

    def run(f, c):
        for layer in net:
            f, c = layer((f, c))
        return f

    f = torch.randn(2, 8, 3000, device="cuda")
    c = torch.randn(2, 3, 3000, device="cuda")
    f2, c2 = torch.randn_like(f), torch.randn_like(c)
    with torch.no_grad():
        want = run(f2.clone(), c2.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                run(f, c)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = run(f, c)
        f.copy_(f2); c.copy_(c2)     # new inputs in the captured buffers
        g.replay()
        torch.cuda.synchronize()
    assert_close(host(out), host(want), "graph replay vs eager", rtol=1e-5, atol_frac=1e-5)


# ------------------------------------------------------------------------------------------------ plan reuse
PLAN_SHAPES = [dict(seed=301, b=2, c=8, n=20000, res=32), dict(seed=302, b=3, c=12, n=5000, res=16),
               dict(seed=303, b=2, c=6, n=20000, res=8), dict(seed=304, b=1, c=5, n=777, res=5),
               dict(seed=305, b=1, c=3, n=3000, res=40)]   # r=40: no shared-memory plane (global atomics on a memset)


@pytest.mark.parametrize("kw", PLAN_SHAPES, ids=lambda k: f"b{k['b']}c{k['c']}n{k['n']}r{k['res']}")
def test_plan_reuse_matches_fresh_build(be, oracle, kw):
    """include/pcfm_b200.h "Plan reuse": the second PVConv block of a stage sees the same coords tensor; its
    avg_voxelize forward / devoxelize backward run on the plan the first block's call left behind
    (pcfm_*_planned).  Same results as a fresh build, checked against the oracle as well."""
    from pcfm import _lib
    feat, coords, res = cases.voxel_case(**kw)
    feat2 = rng(kw["seed"] + 50).standard_normal(feat.shape).astype(np.float32)
    vox, nc = gpu_voxel_coords(coords, res)
    nc = nc.contiguous()
    memo = be.plan_memo
    memo.clear(); memo.enabled = True
    h0, m0 = memo.hits, memo.misses
    _lib.reset_launch_count()
    out1, ind1, cnt1 = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
    l1 = _lib.launch_count()
    out2, ind2, cnt2 = be.pvcnn.avg_voxelize_forward(cu(feat2), vox, res)      # planned
    l2 = _lib.launch_count() - l1
    assert (memo.hits - h0, memo.misses - m0) == (1, 1) and l2 == 1 and l2 < l1
    assert ind2 is ind1 and cnt2 is cnt1
    memo.enabled = False
    out2f, ind2f, cnt2f = be.pvcnn.avg_voxelize_forward(cu(feat2), vox, res)   # fresh build
    memo.enabled = True
    o_out2, o_ind, o_cnt = oracle.avg_voxelize_forward(feat2, host(vox), res)
    assert_equal(host(ind2), o_ind, "ind"); assert_equal(host(cnt2), o_cnt, "cnt")
    # cells of more than 16 points are summed in the plan's record order, which differs from build to build
    assert_close(host(out2), host(out2f), "planned vs fresh avg_voxelize", atol_frac=1e-5)
    assert_close(host(out2), o_out2, "planned avg_voxelize vs oracle", atol_frac=1e-5)
    assert_close(host(out1), oracle.avg_voxelize_forward(feat, host(vox), res)[0], "first call", atol_frac=1e-5)

    gy = rng(kw["seed"] + 60).standard_normal(feat.shape).astype(np.float32)
    gy2 = rng(kw["seed"] + 61).standard_normal(feat.shape).astype(np.float32)
    _, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, nc, out1)
    _lib.reset_launch_count()
    g1 = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), nc, res)
    l1 = _lib.launch_count()
    g2 = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy2), nc, res)       # planned
    assert _lib.launch_count() - l1 == 1 and memo.hits - h0 == 2
    want = oracle.trilinear_devoxelize_backward(gy2, host(inds), host(wgts), res)
    assert_close(host(g2), want, "planned devoxelize backward vs oracle", atol_frac=1e-5)
    assert_close(host(g1), oracle.trilinear_devoxelize_backward(gy, host(inds), host(wgts), res), "first", atol_frac=1e-5)


@pytest.mark.parametrize("kw", PLAN_SHAPES, ids=lambda k: f"b{k['b']}c{k['c']}n{k['n']}r{k['res']}")
def test_dual_plan_build(be, oracle, kw):
    """pcfm_avg_voxelize_forward_dual: the first avg_voxelize of a training stage is told where the block will
    devoxelize and builds the trilinear plan of the backward pass beside its own (one launch); the backward then runs
    on that plan.  Same results as the separate builds / the oracle."""
    from pcfm import _lib
    feat, coords, res = cases.voxel_case(**kw)
    vox, nc = gpu_voxel_coords(coords, res)
    nc = nc.contiguous()
    memo = be.plan_memo
    memo.clear(); memo.enabled = True
    _lib.reset_launch_count()
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res, nc)
    assert _lib.launch_count() == 2                      # dual plan kernel + row kernel (fused build: all PLAN_SHAPES)
    o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, host(vox), res)
    assert_equal(host(ind), o_ind, "ind"); assert_equal(host(cnt), o_cnt, "cnt")
    assert_close(host(out), o_out, "out", atol_frac=1e-5)
    gy = rng(kw["seed"] + 70).standard_normal(feat.shape).astype(np.float32)
    _, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, nc, out)
    h0 = memo.hits
    _lib.reset_launch_count()
    g = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), nc, res)
    assert memo.hits == h0 + 1 and _lib.launch_count() == 1     # the row kernel only
    assert_close(host(g), oracle.trilinear_devoxelize_backward(gy, host(inds), host(wgts), res),
                 "devoxelize backward on the plan built in the forward pass", atol_frac=1e-5)
    _lib.set_option("plan", 1)                           # multi-kernel build: dual falls back to two builds
    try:
        memo.clear()
        out2, ind2, cnt2 = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res, nc)
        g2 = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), nc, res)
    finally:
        _lib.set_option("plan", 0)
    assert_equal(host(ind2), o_ind, "ind (multi-kernel)"); assert_close(host(out2), o_out, "out (multi-kernel)", atol_frac=1e-5)
    assert_close(host(g2), host(g), "grad (multi-kernel)", rtol=1e-5, atol_frac=1e-5)


def test_plan_memo_invalidation(be, oracle):
    """A plan is valid for one coords tensor OBJECT at one version: an in-place write, another tensor with the same
    values, another resolution or a cleared memo all rebuild."""
    feat, coords, res = cases.voxel_case(seed=311, b=2, c=4, n=4000, res=16)
    vox, _ = gpu_voxel_coords(coords, res)
    memo = be.plan_memo
    memo.clear(); memo.enabled = True
    f = cu(feat)
    be.pvcnn.avg_voxelize_forward(f, vox, res)
    h = memo.hits
    be.pvcnn.avg_voxelize_forward(f, vox.clone(), res); assert memo.hits == h       # other object
    be.pvcnn.avg_voxelize_forward(f, vox, res); assert memo.hits == h + 1           # same object
    vox.sub_(1).clamp_(0, res - 1)                                                  # in-place: version bump
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(f, vox, res); assert memo.hits == h + 1
    o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, host(vox), res)
    assert_equal(host(ind), o_ind, "ind after in-place edit"); assert_equal(host(cnt), o_cnt, "cnt")
    assert_close(host(out), o_out, "out after in-place edit", atol_frac=1e-5)
    out, ind, cnt = be.pvcnn.avg_voxelize_forward(f, vox, res); assert memo.hits == h + 2
    assert_close(host(out), o_out, "planned after in-place edit", atol_frac=1e-5)
    from pcfm import _lib
    _lib.set_option("vox_fwd", 1)                                                   # generic variant: no cell plan kept
    try:
        be.pvcnn.avg_voxelize_forward(f, vox, res)
        assert len(memo.entries) == 0
        be.pvcnn.avg_voxelize_forward(f, vox, res); assert memo.hits == h + 2
    finally:
        _lib.set_option("vox_fwd", 0)


def test_plan_reuse_through_autograd_two_blocks(be):
    """Two PVConv-style blocks on one coords tensor, forward + backward through the autograd wrappers: identical
    gradients with and without the plan memo, and the memo serves both the second forward and the second backward
    (the backward runs on autograd's device thread, which keeps its own memo)."""
    from pcfm import functional as F
    from pcfm.nn import Voxelization
    torch.manual_seed(5)
    f0 = torch.randn(2, 8, 6000, device="cuda")
    coords = torch.randn(2, 3, 6000, device="cuda")
    w = torch.randn(2, 8, 6000, device="cuda")

    def run(enabled):
        be.plan_memo.clear(); be.plan_memo.enabled = enabled
        f = f0.clone().requires_grad_(True)
        vm = Voxelization(16, normalize=True, eps=1e-6)
        x = f
        for _ in range(2):
            grid, nc = vm(x, coords)
            x = F.trilinear_devoxelize(torch.tanh(grid), nc, 16, True) + x
        h0 = be.plan_memo.hits
        (x * w).sum().backward()
        torch.cuda.synchronize()
        if enabled:   # both devoxelize backwards run on the trilinear plan the forward pass built (other thread)
            assert be.plan_memo.hits - h0 == 2
        return host(x), host(f.grad)

    try:
        y_on, g_on = run(True)
        y_off, g_off = run(False)
    finally:
        be.plan_memo.enabled = True
    assert_close(y_on, y_off, "forward with / without plan reuse", rtol=1e-5, atol_frac=1e-5)
    assert_close(g_on, g_off, "gradient with / without plan reuse", rtol=1e-5, atol_frac=1e-5)


# ------------------------------------------------------------------------------------------------ randomized shapes
def test_random_small_shapes_pvconv_primitives(be, oracle):
    """40 random (B, C, N, r) combinations, odd sizes included: ragged tails, unaligned rows (no TMA), r not a power
    of two, fewer points than a warp, more voxels than points and vice versa."""
    r0 = rng(1001)
    for it in range(40):
        b = int(r0.integers(1, 4)); c = int(r0.integers(1, 20)); n = int(r0.integers(1, 700)); res = int(r0.integers(2, 21))
        feat, coords, _ = cases.voxel_case(seed=2000 + it, b=b, c=c, n=n, res=res, kind="randn" if it % 2 else "clustered")
        vox, nc = gpu_voxel_coords(coords, res)
        tag = f"[it={it} b={b} c={c} n={n} r={res}]"
        out, ind, cnt = be.pvcnn.avg_voxelize_forward(cu(feat), vox, res)
        o_out, o_ind, o_cnt = oracle.avg_voxelize_forward(feat, host(vox), res)
        assert_equal(host(ind), o_ind, "ind " + tag); assert_equal(host(cnt), o_cnt, "cnt " + tag)
        assert_close(host(out), o_out, "vox out " + tag)
        gy = r0.standard_normal(o_out.shape).astype(np.float32)
        assert_equal(host(be.pvcnn.avg_voxelize_backward(cu(gy), ind, cnt)), oracle.avg_voxelize_backward(gy, o_ind, o_cnt), "vox grad " + tag)
        ncf = host(nc).astype(np.float32)
        outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, cu(ncf), out)
        oo, oi, ow = oracle.trilinear_devoxelize_forward(res, ncf, host(out))
        assert_equal(host(inds), oi, "inds " + tag); assert_equal(host(wgts), ow, "wgts " + tag)
        assert_close(host(outs), oo, "devox out " + tag)
        g2 = r0.standard_normal(oo.shape).astype(np.float32)
        ogx = oracle.trilinear_devoxelize_backward(g2, oi, ow, res)
        assert_close(host(be.pvcnn.trilinear_devoxelize_backward_coords(cu(g2), cu(ncf), res)), ogx, "devox grad (coords) " + tag, atol_frac=1e-5)
        assert_close(host(be.pvcnn.trilinear_devoxelize_backward(cu(g2), inds, wgts, res)), ogx, "devox grad (generic) " + tag, atol_frac=1e-5)


def test_random_small_shapes_metrics_and_pointnet_ops(be, oracle):
    r0 = rng(1002)
    for it in range(25):
        b = int(r0.integers(1, 4)); n = int(r0.integers(1, 600)); m = int(r0.integers(1, 600))
        kind = ["randn", "rand", "lattice"][it % 3]
        a, c = cases.cloud_pair(seed=3000 + it, b=b, n=n, m=m, kind=kind)
        tag = f"[it={it} b={b} n={n} m={m} {kind}]"
        d1, d2, i1, i2 = run_chamfer(be.chamfer, a, c)
        od1, od2, oi1, oi2 = oracle.chamfer_forward(a, c)
        assert_equal(host(i1), oi1, "idx1 " + tag); assert_equal(host(i2), oi2, "idx2 " + tag)
        assert_equal(host(d1), od1, "dist1 " + tag); assert_equal(host(d2), od2, "dist2 " + tag)
        cd = host(be.chamfer3d_allpairs(cu(a), cu(c)))
        want = od1.astype(np.float64).mean(1)[:, None] * 0  # filled below
        for i in range(b):
            for j in range(b):
                e1, e2, _, _ = oracle.chamfer_forward(a[i:i + 1], c[j:j + 1])
                w = e1.astype(np.float64).mean() + e2.astype(np.float64).mean()
                assert abs(cd[i, j] - w) <= 1e-5 * max(abs(w), 1e-12), ("allpairs " + tag, i, j, cd[i, j], w)
    for it in range(12):
        b = int(r0.integers(1, 3)); n = int(r0.integers(3, 500)); m = int(r0.integers(1, min(n, 120) + 1)); c = int(r0.integers(1, 9))
        pts, feat, cfeat = cases.pointnet_case(seed=4000 + it, b=b, c=c, n=n, m=m, kind="lattice" if it % 2 else "rand")
        tag = f"[it={it} b={b} c={c} n={n} m={m}]"
        fidx = be.pvcnn.furthest_point_sampling(cu(pts), m)
        assert_equal(host(fidx), oracle.furthest_point_sampling(pts, m), "fps " + tag)
        ctr = be.pvcnn.gather_features_forward(cu(pts), fidx)
        u = int(r0.integers(1, 20)); rad = float(r0.uniform(0.05, 0.6))
        bq = be.pvcnn.ball_query(ctr, cu(pts), rad, u)
        assert_equal(host(bq), oracle.ball_query(host(ctr), pts, rad, u), "ball query " + tag)
        o3, i3, w3 = be.pvcnn.three_nearest_neighbors_interpolate_forward(cu(pts), ctr, cu(cfeat))
        oo3, oi3, ow3 = oracle.three_nn_interpolate_forward(pts, host(ctr), cfeat)
        assert_equal(host(i3), oi3, "3nn idx " + tag); assert_equal(host(w3), ow3, "3nn w " + tag)
        assert_close(host(o3), oo3, "3nn out " + tag)


def test_unaligned_workspace_pointer(be):
    """The C ABI takes any workspace pointer: blocks are carved on 256-byte boundaries inside it (the *_workspace()
    sizes include the slack)."""
    from pcfm import _lib
    feat, coords, res = cases.voxel_case(seed=96, b=2, c=4, n=3000, res=16)
    vox, nc = gpu_voxel_coords(coords, res)
    f = cu(feat)
    want = be.pvcnn.avg_voxelize_forward(f, vox, res)
    b, c, n = f.shape
    s = res ** 3
    out = torch.empty(b, c, s, device="cuda")
    ind = torch.empty(b, n, dtype=torch.int32, device="cuda")
    cnt = torch.empty(b, s, dtype=torch.int32, device="cuda")
    nb = _lib.lib.pcfm_avg_voxelize_forward_workspace(b, c, n, res)
    ws = torch.empty(nb + 8, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = _lib.lib.pcfm_avg_voxelize_forward(f.data_ptr(), vox.data_ptr(), b, c, n, res, out.data_ptr(), ind.data_ptr(),
                                            cnt.data_ptr(), ws.data_ptr() + 4, nb, st)
    assert rc == 0
    assert torch.equal(ind, want[1]) and torch.equal(cnt, want[2])
    assert_close(host(out), host(want[0]), "out with an unaligned workspace")
    # a workspace that is too small is reported, not overrun
    rc = _lib.lib.pcfm_avg_voxelize_forward(f.data_ptr(), vox.data_ptr(), b, c, n, res, out.data_ptr(), ind.data_ptr(),
                                            cnt.data_ptr(), ws.data_ptr(), 1024, st)
    assert rc != 0


def test_two_devices_in_one_process(be, oracle):
    """Kernel attributes (dynamic shared memory opt-in) are per device: the same process must be able to run the
    big-shared-memory kernels on a second GPU after it has used the first."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    feat, coords, res = cases.voxel_case(seed=95, b=2, c=4, n=20000, res=32)
    want = None
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(dev):
            c = torch.from_numpy(coords).to(dev)
            nc = c - c.mean(2, keepdim=True)
            nc = torch.clamp((nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5) * res, 0, res - 1)
            vox = torch.round(nc).to(torch.int32).contiguous()
            f = torch.from_numpy(feat).to(dev)
            out, ind, cnt = be.pvcnn.avg_voxelize_forward(f, vox, res)
            outs, inds, wgts = be.pvcnn.trilinear_devoxelize_forward(res, True, nc.contiguous(), out)
            g = be.pvcnn.trilinear_devoxelize_backward_coords(outs, nc.contiguous(), res)
            gf = be.pvcnn.avg_voxelize_backward(g, ind, cnt)
            torch.cuda.synchronize(dev)
            got = [t.cpu().numpy() for t in (out, outs, g, gf)]
        if want is None:
            want = got
            o_out, _, _ = oracle.avg_voxelize_forward(feat, vox.cpu().numpy(), res)
            assert_close(got[0], o_out, "voxelize on cuda:0")
        else:
            for a, b_, name in zip(got, want, ("vox", "devox", "devox bwd", "vox bwd")):
                assert_close(a, b_, f"{name} on {dev} vs cuda:0", rtol=1e-6, atol_frac=1e-6)
