#!/usr/bin/env python
"""BASELINE config 3 (flow-matching sampling, 20000-point clouds, batch 32, 100 velocity evaluations): what one
velocity evaluation costs on the PVConv side in eval mode, eager and as a replayed CUDA graph.

    python tests/perf_sampler.py [--batch 32] [--nfe 100]

  primitives : the 6 voxelize -> devoxelize(eval) chains of ContextNet (models.py:402-404 shapes), our kernels only
  modules    : SharedMLP projections + 6 PVConv blocks (cuDNN Conv3d / BatchNorm3d between our kernels), forward only
The reference cannot capture this path (14 of its 20 kernels launch on the legacy stream).  Developer
tool, not part of the test suite; prints one JSON line per measurement."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "point-cloud-flow-matching_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from pcfm import _lib  # noqa: E402
from pcfm import functional as F  # noqa: E402
from pcfm.nn import PVConv, SharedMLP, Voxelization  # noqa: E402

STAGES = [(128, 32), (128, 32), (256, 16), (256, 16), (256, 8), (256, 8)]


def timed(fn, n, warm=2):
    for _ in range(warm):   # allocator growth, one-time function attributes
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def capture(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--points", type=int, default=20000)
    ap.add_argument("--nfe", type=int, default=100)
    args = ap.parse_args()
    B, N = args.batch, args.points
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    coords = torch.randn(B, 3, N, device=dev, generator=g)
    feats = {c: torch.randn(B, c, N, device=dev, generator=g) for c in (128, 256)}
    grids = {(c, r): torch.randn(B, c, r, r, r, device=dev, generator=g) for (c, r) in set(STAGES)}
    vms = [Voxelization(r, normalize=True, eps=1e-6) for (c, r) in STAGES]

    def primitives():
        out = None
        with torch.no_grad():
            for (c, r), vm in zip(STAGES, vms):
                grid, nc = vm(feats[c], coords)                                  # Voxelization + avg_voxelize
                out = F.trilinear_devoxelize(grids[(c, r)], nc, r, False)       # what the conv stack would hand back
        return out

    rows = []
    for name, fn in (("primitives", primitives),):
        ms = timed(fn, 10)
        _lib.reset_launch_count()
        fn()
        launches = _lib.launch_count()
        gr, _ = capture(fn)

        def replay():
            coords.add_(0.0)      # a sampler step moves the points: bumps the version, the graph recomputes the plans
            gr.replay()
        ms_g = timed(replay, args.nfe)
        rows.append({"what": name, "B": B, "N": N, "ms_eager": round(ms, 3), "ms_graph_replay": round(ms_g, 3),
                     "pcfm_launches_per_eval": launches, "mpts_per_s_graph": round(6 * B * N / ms_g / 1e3, 1),
                     "nfe": args.nfe, "ms_per_sample_batch_graph": round(ms_g * args.nfe, 1)})
        print(json.dumps(rows[-1]), flush=True)
        del gr

    torch.manual_seed(0)
    layers, cin = [], 64
    for (c, r) in STAGES:
        if c != cin:
            layers.append(SharedMLP(cin, [c]))
            cin = c
        layers.append(PVConv(c, c, 3, r, with_se=True, normalize=True, eps=1e-6))
    net = torch.nn.ModuleList(layers).to(dev).eval()
    f0 = torch.randn(B, 64, N, device=dev, generator=g)

    def modules():
        with torch.no_grad():
            f, c = f0, coords
            for layer in net:
                f, c = layer((f, c))
        return f

    try:
        ms = timed(modules, 3)
        gr, _ = capture(modules)
        ms_g = timed(gr.replay, 10)
        rows.append({"what": "modules (cuDNN Conv3d/BN3d included), forward, eval", "B": B, "N": N, "ms_eager": round(ms, 2),
                     "ms_graph_replay": round(ms_g, 2), "mpts_per_s_graph": round(6 * B * N / ms_g / 1e3, 1)})
        print(json.dumps(rows[-1]), flush=True)
    except RuntimeError as e:   # e.g. out of memory at a larger batch
        print(json.dumps({"what": "modules", "error": str(e)[:200]}), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "perf_sampler.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
