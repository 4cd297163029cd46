#!/usr/bin/env python
"""How much of the voxel grid does a cloud use after the reference normalisation (modules/voxelization.py:18-22:
centre, divide by twice the largest norm, +0.5)?  The farthest point lands on the sphere of radius r/2, so real clouds
fill a fraction of the r^3 grid.  Input for the core-box variant of the r=32 kernels planned in DESIGN.md section 7
(stage only the central box of every plane; points outside it go through L2).  CPU only (numpy):

    python tests/study_occupancy.py > profiles/r01_study_occupancy.json
"""
import json

import numpy as np


def normalised(c, r):
    n = c - c.mean(1, keepdims=True)
    n = n / (np.linalg.norm(n, axis=0).max() * 2 + 1e-6) + 0.5
    return np.clip(n * r, 0, r - 1)


def clouds(rng, n):
    yield "gaussian", rng.standard_normal((3, n))
    yield "uniform cube", rng.random((3, n)) * 2 - 1
    u = rng.standard_normal((3, n))
    yield "sphere surface", u / np.linalg.norm(u, axis=0)
    # "Scissors-like" (SURVEY 8d): two thin boxes with random orientations, points uniform on them
    parts = []
    for k in range(2):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        p = (rng.random((3, n // 2)) * 2 - 1) * np.array([[1.0], [0.12], [0.03]])
        parts.append(q @ p + rng.standard_normal((3, 1)) * 0.1)
    yield "two thin boxes", np.concatenate(parts, axis=1)


def main():
    rng = np.random.default_rng(0)
    n = 20000
    rows = []
    for name, c in clouds(rng, n):
        for r in (32, 16):
            f = np.floor(normalised(c, r)).astype(int)
            occupied = len(np.unique(f[0] * r * r + f[1] * r + f[2]))
            row = {"cloud": name, "N": n, "r": r, "occupied_cells": occupied, "cells": r ** 3,
                   "bbox_cells": [int(f[a].max() - f[a].min() + 1) for a in range(3)], "inside_central_box": {}}
            for side in (13, 17, 19, 21, 23, 25):
                if side >= r:
                    continue
                lo = r // 2 - side // 2
                inside = np.all((f >= lo) & (f < lo + side), axis=0).mean()
                row["inside_central_box"][str(side)] = round(float(inside), 4)
            rows.append(row)
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
