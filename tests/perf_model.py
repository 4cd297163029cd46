#!/usr/bin/env python
"""BASELINE configs 2 and 3 on the reference's OWN model (unmodified models.py staged under baseline/_ref/, its PVConv /
SharedMLP imports served by dropin/):

  config 3  flow-matching sampling: HybridMLP, 20000-point RGB clouds, batch 32, 100 velocity evaluations = 50 Heun
            predictor-corrector steps (train.py:333-341), eager and with ONE captured step replayed 50 times
  config 2  one training step at B=8 (forward, MSE flow-matching loss, backward, AdamW)

    python tests/perf_model.py [--batch 32] [--steps 50]

Measurement script (not part of the test suite); writes gpurun_out/perf_model.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "point-cloud-flow-matching_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import refmodel  # noqa: E402
from pcfm import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--points", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=50, help="Heun steps (2 velocity evaluations each)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    refmodel.load_models()
    net = refmodel.make_hybrid(dev).eval()
    B, N = args.batch, args.points
    g = torch.Generator(device=dev).manual_seed(0)
    x0 = torch.randn(B, N, 6, device=dev, generator=g)
    cond = torch.randn(B, 129, device=dev, generator=g)
    rows = []

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.reset_launch_count()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), _lib.launch_count(), out

    refmodel.heun_sample(net, x0, cond, 2)          # warm-up: cuDNN algorithm choice, allocator
    ms_e, launches, xe = timed(lambda: refmodel.heun_sample(net, x0, cond, args.steps))
    nfe = 2 * args.steps
    rows.append({"what": "Heun sampler, eager (train.py:333-341)", "B": B, "N": N, "nfe": nfe, "ms_total": round(ms_e, 1),
                 "ms_per_nfe": round(ms_e / nfe, 3), "mpts_per_s": round(B * N * nfe / ms_e / 1e3, 1),
                 "pcfm_launches_per_nfe": launches / nfe})
    print(json.dumps(rows[-1]), flush=True)
    hg = refmodel.HeunGraph(net, x0, cond, args.steps)
    hg.sample(x0)
    ms_g, launches, xg = timed(lambda: hg.sample(x0))
    err = float((xg - xe).abs().max() / xe.abs().max())
    rows.append({"what": "Heun sampler, one captured step replayed", "B": B, "N": N, "nfe": nfe, "ms_total": round(ms_g, 1),
                 "ms_per_nfe": round(ms_g / nfe, 3), "mpts_per_s": round(B * N * nfe / ms_g / 1e3, 1),
                 "max_rel_diff_vs_eager": err, "note": "TF32 convolutions on (PyTorch default): eager and captured runs "
                 "may pick different cuDNN algorithms; tests/test_gpu_refmodel.py checks 1e-4 with TF32 off"})
    print(json.dumps(rows[-1]), flush=True)
    del hg
    torch.cuda.empty_cache()

    # config 2: one training step at B = 8
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    Bt = 8
    data = torch.rand(Bt, N, 6, device=dev, generator=g) * 2 - 1
    c8 = cond[:Bt]

    def train_step():
        z = torch.randn(Bt, N, 6, device=dev, generator=g)
        t = torch.rand(Bt, device=dev, generator=g)
        xt = (1 - t)[:, None, None] * z + t[:, None, None] * data
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(net(xt, t, c8), data - z)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        train_step()
    ms_t, launches, _ = timed(lambda: [train_step() for _ in range(5)])
    rows.append({"what": "training step (forward + backward + AdamW), HybridMLP", "B": Bt, "N": N, "ms_per_step": round(ms_t / 5, 2),
                 "mpts_per_s": round(Bt * N * 5 / ms_t / 1e3, 2), "pcfm_launches_per_step": launches / 5})
    print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "perf_model.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
