"""Shared helpers for the parity tests: seeded inputs and the acceptance rule for fp32 tensors."""
import numpy as np

# north_star: "fp32 features, distances and gradients within 1e-5 relative".  The reference itself accumulates
# with order-nondeterministic fp32 atomics, so the bar is applied as in SURVEY.md 7.4(3):
#   normwise   max|a-b| <= RTOL * max|b|      and
#   elementwise |a-b| <= RTOL*|b| + ATOL_FRAC * max|b|
RTOL = 1e-5
ATOL_FRAC = 1e-6


def assert_close(a, b, what, rtol=RTOL, atol_frac=ATOL_FRAC):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    scale = max(np.abs(b).max(), 1e-30)
    diff = np.abs(a - b)
    assert diff.max() <= rtol * scale, f"{what}: normwise {diff.max() / scale:.3e} > {rtol}"
    bad = diff > rtol * np.abs(b) + atol_frac * scale
    assert not bad.any(), f"{what}: {int(bad.sum())} elements outside |a-b|<=rtol|b|+{atol_frac}*max|b|"


def assert_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        nz = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(nz)} mismatches, first at {tuple(nz[0])}: {a[tuple(nz[0])]} vs {b[tuple(nz[0])]}")


def rng(seed):
    return np.random.default_rng(seed)


def cloud_cf(r, b, n, kind="randn"):
    """channel-first coords [B,3,N]"""
    if kind == "randn":
        return r.standard_normal((b, 3, n)).astype(np.float32)
    if kind == "lattice":  # many exact ties / duplicated points
        return (r.integers(0, 5, size=(b, 3, n)) * 0.25).astype(np.float32)
    return r.random((b, 3, n)).astype(np.float32)


def vox_float_coords(r, b, n, res, kind="uniform"):
    """float voxel coords in [0,res-1] incl. exact lattice hits and the upper boundary"""
    c = (r.random((b, 3, n)) * (res - 1)).astype(np.float32)
    if kind == "edges":
        k = max(1, n // 8)
        c[:, :, :k] = np.round(c[:, :, :k])          # exactly on lattice planes (frac == 0)
        c[:, :, k:2 * k] = res - 1                     # upper corner
        c[:, :, 2 * k:3 * k] = 0                       # lower corner
    return np.clip(c, 0, res - 1).astype(np.float32)
