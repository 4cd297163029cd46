"""Seeded input generators shared by make_golden.py (runs the real reference on a GPU), the CPU oracle tests and
the GPU parity tests.  Small on purpose: the fixtures are committed."""
import numpy as np


def _r(seed):
    return np.random.default_rng(seed)


def voxel_case(seed, b, c, n, res, kind="randn"):
    r = _r(seed)
    feat = r.standard_normal((b, c, n)).astype(np.float32)
    if kind == "randn":
        coords = r.standard_normal((b, 3, n)).astype(np.float32)
    else:  # clustered: many points per voxel
        coords = (r.standard_normal((b, 3, n)) * 0.05 + r.integers(0, 3, size=(b, 3, n))).astype(np.float32)
    return feat, coords, res


def devox_case(seed, b, c, n, res):
    r = _r(seed)
    coords = (r.random((b, 3, n)) * (res - 1)).astype(np.float32)
    k = max(1, n // 8)
    coords[:, :, :k] = np.round(coords[:, :, :k])
    coords[:, :, k:2 * k] = res - 1
    coords[:, :, 2 * k:3 * k] = 0
    vox = r.standard_normal((b, c, res ** 3)).astype(np.float32)
    gy = r.standard_normal((b, c, n)).astype(np.float32)
    return np.clip(coords, 0, res - 1).astype(np.float32), vox, gy, res


def cloud_pair(seed, b, n, m, kind="randn", dim=3):
    r = _r(seed)
    if kind == "lattice":  # exact ties and duplicated points
        a = (r.integers(0, 6, size=(b, n, dim)) * 0.125).astype(np.float32)
        c = (r.integers(0, 6, size=(b, m, dim)) * 0.125).astype(np.float32)
    elif kind == "rand":
        a = r.random((b, n, dim)).astype(np.float32)
        c = r.random((b, m, dim)).astype(np.float32)
    else:
        a = r.standard_normal((b, n, dim)).astype(np.float32)
        c = r.standard_normal((b, m, dim)).astype(np.float32)
    return a, c


def pointnet_case(seed, b, c, n, m, kind="rand"):
    r = _r(seed)
    if kind == "lattice":
        pts = (r.integers(0, 5, size=(b, 3, n)) * 0.25).astype(np.float32)
    else:
        pts = r.random((b, 3, n)).astype(np.float32)
    feat = r.standard_normal((b, c, n)).astype(np.float32)
    cfeat = r.standard_normal((b, c, m)).astype(np.float32)
    return pts, feat, cfeat


# name -> (generator, kwargs); make_golden.py stores the reference's outputs under the same names
VOXEL = {
    "vox_c1small": dict(seed=1, b=2, c=8, n=2048, res=32),
    "vox_odd": dict(seed=2, b=2, c=3, n=1001, res=7),
    "vox_dense": dict(seed=3, b=1, c=4, n=4096, res=4, kind="clustered"),
}
DEVOX = {
    "devox_r16": dict(seed=11, b=2, c=6, n=1500, res=16),
    "devox_r5": dict(seed=12, b=1, c=3, n=333, res=5),
}
# 6-D clouds (xyz + rgb).  Chunk lengths (cloud size mod 2048) avoid 1 and 3 (mod 8): there the reference's unrolled
# loop bound (chamfer6D.cu:28 `end_k-(end_k&6)`, `&3` in chamfer3D.cu) runs two / one points past the end of the chunk
# and can return an out-of-range index (observed on a B200: idx 513 for m = 513)
CHAMFER6D = {
    "ch6_unit": dict(seed=24, b=3, n=100, m=200, kind="rand", dim=6),
    "ch6_lattice": dict(seed=25, b=2, n=700, m=516, kind="lattice", dim=6),
    "ch6_big": dict(seed=26, b=1, n=2500, m=2300, kind="randn", dim=6),   # two chunks of the 2048-point tile
}
CHAMFER = {
    "ch_unit": dict(seed=21, b=4, n=100, m=200, kind="rand"),       # unit_test.py:15-16 shapes
    "ch_lattice": dict(seed=22, b=2, n=700, m=513, kind="lattice"),
    "ch_big": dict(seed=23, b=1, n=2500, m=2300, kind="randn"),
}
EMD = {
    "emd_small": dict(seed=31, b=2, n=256, m=256, kind="rand"),
    "emd_ragged": dict(seed=32, b=2, n=300, m=150, kind="rand"),
}
POINTNET = {
    "pn_rand": dict(seed=41, b=2, c=5, n=1024, m=256, kind="rand"),
    "pn_lattice": dict(seed=42, b=2, c=3, n=600, m=150, kind="lattice"),
}
BALL = dict(radius=0.2, u=16)
