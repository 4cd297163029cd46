#!/usr/bin/env python
"""bench.py — headline benchmark of the point-cloud hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config 2 of BASELINE.json, named in `config.workload`): the PVConv point-voxel primitive chain of one
HybridMLP velocity-net step — B=8 clouds of N=20000 points through the 6 PVConv instances of ContextNet
(models.py:402-404: 2x(C=128,r=32), 2x(C=256,r=16), 2x(C=256,r=8)); per instance avg_voxelize forward,
trilinear_devoxelize forward (training mode), trilinear_devoxelize backward, avg_voxelize backward.
One step = all 6 chains, forward + backward.  value = 6*B*N point-passes / step time, in Mpts/s.
Conv3d / BatchNorm between the primitives are cuDNN library calls outside the path and are not
timed; each primitive reads tensors that no kernel of the same step has just written (nothing is L2-hot).

Also measured and reported in the same JSON line:
  roofline      dominant primitive: algorithmic bytes (SURVEY.md 8d formulas) / its CUDA-event time vs measured HBM peak
  e2e           the same chains through the autograd operator API with HOST (pinned) inputs and outputs
  cpu_baseline  the same ops in stock PyTorch on the host cores (oracle/torch_cpu.py), bounded sample
  allpairs      all-pairs Chamfer / EMD cloud-pairs/s (config 4 shapes), rows sharded over the N ranks
  pvconv_modules  the same 6 PVConv blocks as nn.Modules (cuDNN Conv3d / BatchNorm included), forward + backward

--impl reference runs only the PyTorch-CPU arm (rank 0) and prints the same line shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "point-cloud-flow-matching_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "pvconv_primitives_fwd_bwd_mpts_per_s"
UNIT = "Mpts/s"
B, NPTS = 8, 20000
STAGES = [(128, 32), (128, 32), (256, 16), (256, 16), (256, 8), (256, 8)]  # (C, r) of the 6 PVConv instances
CONFIG = {
    "workload": "BASELINE.json configs[1]: PVConv primitive chain of one velocity-net step, B=8, N=20000, "
                "(C,r) = 2x(128,32) 2x(256,16) 2x(256,8); vox fwd + devox fwd(train) + devox bwd + vox bwd per instance",
    "batch": B, "points": NPTS, "stages": STAGES,
    "plans": "the two PVConv blocks of a stage share one coords tensor (models.py:385-389), so the second block's "
             "avg_voxelize forward / devoxelize forward / devoxelize backward reuse the first block's cell plans "
             "(pcfm_*_planned, pcfm_trilinear_devoxelize_forward_cells); the first avg_voxelize of a stage builds the "
             "mean AND the trilinear plan in one launch (pcfm_avg_voxelize_forward_dual, what pcfm.nn.Voxelization "
             "calls); the plan memo is cleared at the start of EVERY step (new point clouds every training step): "
             "3 dual builds per step",
    "l2": "inputs larger than L2: ~4.8 GB touched per step, every primitive reads tensors not written in that step",
    "parallelism": "one process per GPU, replicas (data parallel), no data-path collective",
}


def algorithmic_bytes(op, c, r, b=B, n=NPTS):
    s = r ** 3
    return {  # SURVEY.md 8(d); zero-fill of outputs not counted
        "vox_fwd": 4 * b * (c * n + 3 * n + c * s + n + s),
        "vox_bwd": 4 * b * (c * s + n + s + c * n),
        "devox_fwd": 4 * b * (3 * n + c * s + c * n + 16 * n),
        "devox_bwd": 4 * b * (c * n + 16 * n + c * s),
    }[op]


# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the warm-up tail and the timed region: an NVML thread (10 ms
    period; `nvidia-smi -lms` needs ~0.5 s to produce its first line and the timed region is ~0.1 s), `nvidia-smi`
    one-shot as the fall-back when pynvml cannot initialise."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        import threading
        self.gpu_index = gpu_index
        self.sm, self.bits, self.mx, self.src = [], 0, None, None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.src = "nvml"
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        except Exception:  # noqa: BLE001
            self.nv = None

    def _reasons(self):
        for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            f = getattr(self.nv, fn, None)
            if f is not None:
                try:
                    return int(f(self.h))
                except Exception:  # noqa: BLE001
                    continue
        return 0

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.bits |= self._reasons()
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            o = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={q}",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            r = [c.strip() for c in o.strip().splitlines()[0].split(",")]
            self.sm.append(float(r[0])); self.mx = float(r[1]); self.src = "nvidia-smi (one sample, end of timed region)"
            for (name, bit), v in zip(self.REASONS, r[2:6]):
                if v.lower().startswith("active"):
                    self.bits |= bit
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join(timeout=2)
        if not self.sm:
            self._smi_once()
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "source": "unavailable"}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.mx,
                "reasons": sorted(n for n, b in self.REASONS if self.bits & b), "samples": len(self.sm),
                "source": self.src}


class gpu_local_cpus:
    """Context: the calling thread runs on the CPUs NVML reports as local to the GPU, so that pinned buffers allocated
    inside are placed on the GPU's NUMA node (a cross-socket pinned buffer halves the achievable PCIe rate).  No-op
    when NVML or the mask is unavailable.  `.cpus` = the CPUs used (None = unchanged)."""

    def __init__(self, gpu_index):
        self.gpu_index, self.cpus, self.saved = gpu_index, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                    phys = int(ids[self.gpu_index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
            cpus = {64 * w + k for w, m in enumerate(words) for k in range(64) if (int(m) >> k) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus and cpus != allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus)
                self.cpus = sorted(cpus)
        except Exception:  # noqa: BLE001
            pass
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
def cpu_chain_step(torch, tc, feats, coords, batch):
    """One bounded sample of the workload on the host: one chain per distinct stage shape."""
    pts = 0
    for (c, r), f, x in zip([(128, 32), (256, 16), (256, 8)], feats, coords):
        tc.pvconv_primitives_fwd_bwd(f[:batch], x[:batch], r)
        pts += batch * NPTS
    return pts


def make_cpu_inputs(torch):
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(B, c, NPTS, generator=g) for c in (128, 256, 256)]
    coords = [torch.randn(B, 3, NPTS, generator=g) for _ in range(3)]
    return feats, coords


def run_reference_arm(args):
    """PyTorch-CPU statement of the path on the host cores (the reference ships no CPU implementation of its CUDA
    kernels; north_star names this arm: scatter_add voxelize, grid_sample devoxelize, autograd backward)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import torch_cpu as tc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feats, coords = make_cpu_inputs(torch)
    t0 = time.perf_counter()
    cpu_chain_step(torch, tc, feats, coords, 2)            # probe at B=2
    probe = time.perf_counter() - t0
    budget = 150.0                                          # seconds for the whole K+W run
    batch = B
    while batch > 1 and probe * (batch / 2) * (args.steps + args.warmup) > budget:
        batch //= 2
    for _ in range(args.warmup + 2):   # + 2: the first passes of a fresh process page in the thread pool and the
        cpu_chain_step(torch, tc, feats, coords, batch)   # allocator arenas (an outlier of 4x was seen without them)
    t0 = time.perf_counter()
    pts = 0
    for _ in range(args.steps):
        pts += cpu_chain_step(torch, tc, feats, coords, batch)
    dt = time.perf_counter() - t0
    val = pts / dt / 1e6
    sample = f"per step: one chain per stage shape (3 of the 6), batch {batch} of {B}, N={NPTS}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# --------------------------------------------------------------------------------------------------------------
def check_against_oracle(torch, P, be, chains, nch=4):
    """Warm-up check of the timed kernels at the FULL bench shapes: the first `nch` channels of one chain per stage shape
    against the C oracle (oracle/pcfm_oracle.c) — every (batch, channel) row is an independent problem, so a channel
    slice of the full-size call is checked exactly.  Asserts indices bit-exact and values within 1e-5 (normwise)."""
    import numpy as np
    from oracle import cpu_ref
    cpu_ref.build()
    worst, rows = 0.0, []

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

    for ch in chains[::2]:
        c, r = ch["c"], ch["r"]
        be.plan_memo.clear()
        out, ind, cnt = P.avg_voxelize_forward(ch["feat"], ch["vox"], r, ch["nc"])
        outs, inds, wgts = P.trilinear_devoxelize_forward(r, True, ch["nc"], ch["conv_out"])
        ggrid = P.trilinear_devoxelize_backward_coords(ch["gy"], ch["nc"], r)
        gfeat = P.avg_voxelize_backward(ch["gconv"], ind, cnt)
        h = lambda t: t.detach().cpu().numpy()  # noqa: E731
        o_out, o_ind, o_cnt = cpu_ref.avg_voxelize_forward(h(ch["feat"][:, :nch].contiguous()), h(ch["vox"]), r)
        o_outs, o_inds, o_wgts = cpu_ref.trilinear_devoxelize_forward(r, h(ch["nc"]), h(ch["conv_out"][:, :nch].contiguous()))
        o_gg = cpu_ref.trilinear_devoxelize_backward(h(ch["gy"][:, :nch].contiguous()), o_inds, o_wgts, r)
        o_gf = cpu_ref.avg_voxelize_backward(h(ch["gconv"][:, :nch].contiguous()), o_ind, o_cnt)
        assert np.array_equal(h(ind), o_ind) and np.array_equal(h(cnt), o_cnt), "voxel indices / counts"
        assert np.array_equal(h(inds), o_inds) and np.array_equal(h(wgts), o_wgts), "trilinear indices / weights"
        errs = {"vox_fwd": rel(h(out[:, :nch]), o_out), "devox_fwd": rel(h(outs[:, :nch]), o_outs),
                "devox_bwd": rel(h(ggrid[:, :nch]), o_gg), "vox_bwd": rel(h(gfeat[:, :nch]), o_gf)}
        for k, e in errs.items():
            assert e <= 1e-5, f"bench warm-up: {k} C={c} r={r} differs from the oracle by {e:.3e}"
        worst = max(worst, max(errs.values()))
        rows.append({"C": c, "r": r, **{k: float(f"{e:.2e}") for k, e in errs.items()}})
    return {"checked": "first %d channels of one chain per stage shape at the full bench sizes vs the C oracle: voxel / "
                       "trilinear indices, counts, weights bit-exact; values normwise" % nch,
            "max_rel_err": float(f"{worst:.2e}"), "tolerance": 1e-5, "per_shape": rows}


def dp_training_step(torch, dist, dev, rank, world, barrier, max_over_ranks, _lib, steps=6):
    """BASELINE config 2 as train.py runs it: the reference's HybridMLP (models.py:604-694, imported UNMODIFIED from
    baseline/_ref/models.py; its PVConv / SharedMLP imports are served by dropin/) wrapped in DistributedDataParallel
    exactly like train.py:241-244, per-GPU batch 8 x 20000 RGB points, cond = latent 128 + 1; one step = forward, MSE
    flow-matching loss (train.py:604-640), backward with the bucketed gradient all-reduce, AdamW step."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refmodel
    if not refmodel.staged():
        return {"unavailable": "baseline/_ref/models.py not staged (python __graft_entry__.py build)"}
    net = refmodel.make_hybrid(dev).train()
    nparam = sum(p.numel() for p in net.parameters())
    model = net
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model = DDP(net, device_ids=[dev.index], output_device=dev.index, broadcast_buffers=False,
                    find_unused_parameters=False)                                   # train.py:241-244
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0.0)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    data = torch.rand(B, NPTS, 6, device=dev, generator=g) * 2 - 1
    cond = torch.randn(B, 129, device=dev, generator=g)

    def one_step():
        z = torch.randn(B, NPTS, 6, device=dev, generator=g)
        t = torch.rand(B, device=dev, generator=g)
        x_t = (1.0 - t)[:, None, None] * z + t[:, None, None] * data
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(x_t, t, cond), data - z)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        one_step()
    barrier()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one_step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    out = {"model": "HybridMLP from the unmodified reference models.py (defaults: stages (128,256,256) x 2 blocks, "
                    "r = (32,16,8)), point_dim 6, cond_dim 129",
           "parallelism": f"DistributedDataParallel x{world}" if world > 1 else "single GPU (no all-reduce)",
           "per_gpu_batch": B, "points": NPTS, "parameters": nparam,
           "grad_allreduce_bytes_per_step": nparam * 4 if world > 1 else 0,
           "ms_per_step": ms, "mpts_per_s": B * NPTS * world / ms / 1e3, "scaling": "weak",
           "pcfm_launches_per_step": _lib.launch_count() / steps, "loss": float(loss.detach()), "steps": steps}
    del model, net, opt
    return out


def time_reference_cuda(torch, chains):
    """BASELINE.md column 2: the UNMODIFIED reference CUDA kernels (oracle/_ref/pvcnn) on the same chains, same GPU."""
    from oracle import build_ref
    R = build_ref.load_ref("pvcnn")
    if R is None:
        return {"unavailable": "oracle/_ref/pvcnn not built"}

    def ref_step():
        for ch in chains:
            r = ch["r"]
            out, ind, cnt = R.avg_voxelize_forward(ch["feat"], ch["vox"], r)
            outs, inds, wgts = R.trilinear_devoxelize_forward(r, True, ch["nc"], ch["conv_out"])
            R.trilinear_devoxelize_backward(ch["gy"], inds, wgts, r)
            R.avg_voxelize_backward(ch["gconv"], ind, cnt)

    ref_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the reference launches on the legacy default stream: events recorded there bracket its kernels
    legacy = torch.cuda.default_stream()
    e0.record(legacy)
    n_it = 2
    for _ in range(n_it):
        ref_step()
    e1.record(legacy)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_it
    return {"ms_per_step": ms, "value": len(chains) * B * NPTS / ms / 1e3, "unit": UNIT, "steps": n_it,
            "what": "third_party/pvcnn kernels compiled unmodified for sm_100a (oracle/build_ref.py), same chains, same "
                    "resident inputs, this GPU"}


# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pcfm  # noqa: F401
    from pcfm import _lib, backend as be
    from pcfm import functional as F
    from pcfm.nn import Voxelization

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- resident inputs: distinct buffers per PVConv instance ----------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    chains = []
    for k, (c, r) in enumerate(STAGES):
        feat = torch.randn(B, c, NPTS, device=dev, generator=g)
        if k % 2 == 1 and STAGES[k - 1] == (c, r):      # second block of a stage: the SAME coords tensor objects
            nc, vox = chains[-1]["nc"], chains[-1]["vox"]
        else:
            coords = torch.randn(B, 3, NPTS, device=dev, generator=g)
            nc = coords - coords.mean(2, keepdim=True)                         # modules/voxelization.py:16-24
            nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + 1e-6) + 0.5
            nc = torch.clamp(nc * r, 0, r - 1).contiguous()
            vox = torch.round(nc).to(torch.int32).contiguous()
        chains.append(dict(c=c, r=r, feat=feat, nc=nc, vox=vox,
                           conv_out=torch.randn(B, c, r ** 3, device=dev, generator=g),     # what devox fwd reads
                           gy=torch.randn(B, c, NPTS, device=dev, generator=g),             # upstream grad of devox
                           gconv=torch.randn(B, c, r ** 3, device=dev, generator=g)))       # what vox bwd reads
    P = be.pvcnn

    def step(rec=None):
        be.plan_memo.clear()        # a training step sees new clouds: no plan survives from the previous step
        for ch in chains:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if rec is not None else None
            if ev: ev[0].record()
            out, ind, cnt = P.avg_voxelize_forward(ch["feat"], ch["vox"], ch["r"], ch["nc"])  # as pcfm.nn.Voxelization
            if ev: ev[1].record()
            outs, inds, wgts = P.trilinear_devoxelize_forward(ch["r"], True, ch["nc"], ch["conv_out"])
            if ev: ev[2].record()
            ggrid = P.trilinear_devoxelize_backward_coords(ch["gy"], ch["nc"], ch["r"])  # what the autograd op calls
            if ev: ev[3].record()
            gfeat = P.avg_voxelize_backward(ch["gconv"], ind, cnt)
            if ev:
                ev[4].record()
                rec.append((ch["c"], ch["r"], ev))
        return gfeat

    step()
    barrier()
    oracle_check = None
    if rank == 0 and not args.no_oracle_check:
        oracle_check = check_against_oracle(torch, P, be, chains)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3) + 60):   # >= W warm-up steps (~0.2 s), under the clock sampler
        step()
    barrier()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _lib.launch_count()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    pts_per_step = len(STAGES) * B * NPTS * world
    value = pts_per_step / (ms_per_step * 1e-3) / 1e6

    # ---------------- per-primitive CUDA-event times (same launches, instrumented pass) ----------------
    rec = []
    for _ in range(3):
        step(rec)
    torch.cuda.synchronize()
    per = {}
    for c, r, ev in rec:
        for k, name in enumerate(["vox_fwd", "devox_fwd", "devox_bwd", "vox_bwd"]):
            per.setdefault((name, c, r), []).append(ev[k].elapsed_time(ev[k + 1]))
    peak, peak_src = measured_peak_gbs()
    prim = []
    for (name, c, r), ts in sorted(per.items()):
        t = statistics.mean(ts)
        by = algorithmic_bytes(name, c, r)
        prim.append({"op": name, "C": c, "r": r, "ms": round(t, 4), "GBps": round(by / t / 1e6, 1),
                     "frac": round(by / t / 1e6 / peak, 4)})
    top = max(prim, key=lambda d: d["ms"])
    top_bytes = algorithmic_bytes(top["op"], top["C"], top["r"])
    tot_bytes = sum(algorithmic_bytes(n_, c, r) * len(ts) / 3 for (n_, c, r), ts in per.items())
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of that op's main kernel, one launch, from the committed capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        traffic = tj["dram_bytes_per_launch"].get(f"{top['op']},{top['C']},{top['r']}")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": f"{top['op']} C={top['C']} r={top['r']} (op = row kernel + its share of the stage's plan build: "
                          "mean over the two blocks of the stage)",
                "achieved": top["GBps"], "peak": peak, "unit": "GB/s", "frac": round(top["GBps"] / peak, 4),
                "traffic": traffic, "traffic_source": "profiles/r02_ncu_traffic.json (ncu --set full, same shapes, main kernel of the op)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": top_bytes,
                "step_aggregate": {"algorithmic_GB_per_step": round(tot_bytes / 1e9, 3),
                                   "achieved_GBps": round(tot_bytes / (ms_per_step * 1e-3) / 1e9 * 1.0, 1),
                                   "frac": round(tot_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4)},
                "per_primitive": prim}

    # ---------------- e2e: autograd operator API, HOST (pinned) inputs and outputs ----------------
    h_src = [(torch.randn(B, c, NPTS), torch.randn(B, 3, NPTS)) for (c, r) in STAGES]
    with gpu_local_cpus(local) as numa:   # pinned pages are placed when they are locked: do that next to the GPU
        h_in = [(hf.pin_memory(), hc.pin_memory()) for hf, hc in h_src]
        h_out2 = [[torch.empty(B, c, NPTS).pin_memory() for (c, r) in STAGES] for _ in range(2)]  # double-buffered results
    del h_src
    h_out = h_out2[0]
    voxmods = [Voxelization(r, normalize=True, eps=1e-6) for (c, r) in STAGES]

    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    in_flight = []

    def e2e_step(k=0):
        """Every step copies all inputs host->device and all results device->host.  Three streams: uploads of the next
        PVConv instance and downloads of the previous one overlap the kernels of the current one (full-duplex PCIe).
        Two steps may be in flight (results double-buffered in pinned memory): the host waits for step k-2 before it
        issues step k, like a training loop that reads its loss one step late."""
        main = torch.cuda.current_stream(dev)
        if len(in_flight) >= 2:
            in_flight.pop(0).synchronize()
        h_out = h_out2[k & 1]
        staged = []
        with torch.cuda.stream(s_in):
            for hf, hc in h_in:
                f = hf.to(dev, non_blocking=True)
                x = hc.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                staged.append((f, x, ev))
        for (c, r), (f, x, ev), ho, vm in zip(STAGES, staged, h_out, voxmods):
            main.wait_event(ev)
            f.record_stream(main)
            x.record_stream(main)
            f.requires_grad_(True)
            grid, nc = vm(f, x)
            o = F.trilinear_devoxelize(grid, nc, r, True)
            o.sum().backward()
            g = f.grad
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                g.record_stream(s_out)
                ho.copy_(g, non_blocking=True)
        end = torch.cuda.Event()
        end.record(s_out)
        in_flight.append(end)

    e2e_steps = max(3, args.steps // 4)
    for k in range(2):
        e2e_step(k)
    barrier()
    in_flight.clear()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    barrier()   # includes torch.cuda.synchronize(): every copy of the timed steps has completed
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    h2d = sum(hf.numel() * 4 + hc.numel() * 4 for hf, hc in h_in)
    d2h = sum(ho.numel() * 4 for ho in h_out)
    e2e = {"value": pts_per_step / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "pcie_GBps_each_way": {"h2d": h2d / e2e_s / 1e9, "d2h": d2h / e2e_s / 1e9},
           "pinned_on_cpus": ("GPU-local " + str(len(numa.cpus)) + " of " + str(os.cpu_count())) if numa.cpus else "all",
           "api": "pcfm.nn.Voxelization + pcfm.functional.trilinear_devoxelize + autograd backward; pinned host tensors "
                  "in and out every step, uploads / kernels / downloads on three streams, two steps in flight"}
    del h_in, h_out, h_out2

    # ---------------- module level: the PVConv pyramid of ContextNet, forward + backward, cuDNN included ----------------
    module = None
    if not args.skip_module:
        try:
            from pcfm.nn import PVConv, SharedMLP
            torch.manual_seed(0)
            layers, cin = [], 64
            for (c, r) in STAGES:                      # models.py:372-386: SharedMLP projection, then PVConv blocks
                if c != cin:
                    layers.append(SharedMLP(cin, [c]))
                    cin = c
                layers.append(PVConv(c, c, 3, r, with_se=True, normalize=True, eps=1e-6))
            net = torch.nn.ModuleList(layers).to(dev).train()
            mf = torch.randn(B, 64, NPTS, device=dev, generator=g).requires_grad_(True)
            mc = torch.randn(B, 3, NPTS, device=dev, generator=g)

            def module_step():
                f, c = mf, mc.clone()              # a new coords tensor every step: no memo hit across steps
                for layer in net:
                    f, c = layer((f, c))
                f.square().mean().backward()

            def time_module(n_it):
                module_step(); module_step()
                torch.cuda.synchronize()
                m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                _lib.reset_launch_count()
                m0.record()
                for _ in range(n_it):
                    module_step()
                m1.record()
                torch.cuda.synchronize()
                return m0.elapsed_time(m1) / n_it, _lib.launch_count() / n_it

            ms_mod, launches_mod = time_module(5)
            # the same pyramid with the voxel grids in channels_last_3d (NDHWC) for the cuDNN layers
            PVConv.channels_last = True
            try:
                net_cl = net.to(memory_format=torch.channels_last_3d)
                ms_cl, _ = time_module(5)
            finally:
                PVConv.channels_last = False
                net.to(memory_format=torch.contiguous_format)
            module = {"what": "SharedMLP projections + 6 PVConv blocks (Conv3d/BatchNorm3d/SE through cuDNN/cuBLAS, TF32 "
                              "conv default) of ContextNet on ONE shared coords tensor, forward + backward, B=8, N=20000",
                      "ms_fwd_bwd": ms_mod, "mpts_per_s": B * NPTS * len(STAGES) / ms_mod / 1e3,
                      "pcfm_launches_per_step": launches_mod,
                      "primitive_share": round(ms_per_step / ms_mod, 3),
                      "ms_fwd_bwd_channels_last": ms_cl,
                      "channels_last": "PVConv.channels_last = True: voxel grids in torch.channels_last_3d for Conv3d / "
                                       "BatchNorm3d / SE (cuDNN NDHWC kernels), NCDHW for the point-voxel primitives"}
            del net, mf, mc
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            module = {"error": repr(e)}

    # ---------------- all-pairs CD / EMD (config 4 shapes), rows sharded over ranks ----------------
    # Timed region = what an evaluation run does: the reference set arrives SHARDED over the ranks and is
    # all-gathered (24.6 MB), every rank computes its block of rows with the all-pairs kernels, and the row blocks are
    # all-gathered into the full matrix on every rank (4 MB) - both collectives inside the CUDA-event window.
    allpairs = None
    try:
        if args.skip_allpairs:
            raise RuntimeError("skipped (--skip-allpairs)")
        from pcfm import evaluation as ev
        gg = torch.Generator(device=dev).manual_seed(7)
        S = torch.randn(1000, 2048, 3, device=dev, generator=gg) * 0.3
        R = torch.randn(1000, 2048, 3, device=dev, generator=gg) * 0.3
        rlo, rhi = ev.shard_range(1000, rank, world)
        R_local = R[rlo:rhi].contiguous()          # this rank's shard of the reference set
        r_counts = [h - l for l, h in (ev.shard_range(1000, k, world) for k in range(world))]
        del R
        be.chamfer3d_allpairs(S[:8], S[:64])
        ev.all_gather_clouds(R_local[:2].contiguous(), [2] * world); barrier()
        a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a0.record()
        R = ev.all_gather_clouds(R_local, r_counts)                          # collective 1: reference set
        a1.record()
        cd = ev.pairwise_matrix_sharded(S, R, "cd", block=250, gather=True)  # kernels + collective 2: row blocks
        a2.record(); barrier()
        cd_ms = max_over_ranks(a0.elapsed_time(a2))
        gather_ms = max_over_ranks(a0.elapsed_time(a1))
        full_emd = world >= 8 or args.full_emd
        nref_emd = 1000 if full_emd else 16                     # N < 8: bounded EMD sample (1000 x 16 pairs)
        be.emd_allpairs(S[:2], R[:2]); barrier()
        a0.record()
        emd = ev.pairwise_matrix_sharded(S, R[:nref_emd].contiguous(), "emd", block=250, gather=True)
        a1.record(); barrier()
        emd_ms = max_over_ranks(a0.elapsed_time(a1))
        # the complete metric set on the CD matrices (MMD / COV / 1-NNA: S x R, and S x S, R x R as symmetric halves)
        a0.record()
        metrics = ev.compute_metrics(S, R, metrics=("cd",), block=250)
        a1.record(); barrier()
        metrics_ms = max_over_ranks(a0.elapsed_time(a1))
        allpairs = {"scaling": "strong", "clouds": "1000 x 1000 x 2048 pts (config 4)",
                    "timed": "all-gather of the sharded reference set + all-pairs kernels on this rank's row block + "
                             "all-gather of the row blocks (full matrix on every rank); CUDA events, max over ranks",
                    "cd_pairs_per_s": 1000 * 1000 / (cd_ms * 1e-3), "cd_ms": cd_ms, "ref_all_gather_ms": gather_ms,
                    "ref_all_gather_bytes": 1000 * 2048 * 3 * 4, "matrix_all_gather_bytes": 1000 * 1000 * 4,
                    "cd_evals_per_s": 1000 * 1000 * 2048 * 2048 / (cd_ms * 1e-3),
                    "emd_pairs_per_s": 1000 * nref_emd / (emd_ms * 1e-3), "emd_ms": emd_ms,
                    "emd_sample": f"1000 x {nref_emd} pairs" + (" (the full matrix)" if full_emd else ""),
                    "cd_checksum": float(cd.double().sum()), "emd_checksum": float(emd.double().sum()),
                    "cd_metrics": metrics["cd"], "cd_metrics_ms": metrics_ms,
                    "cd_metrics_pairs": 1000 * 1000 + 2 * (1000 * 1001 // 2),
                    "cd_metrics_note": "S x R in full, S x S and R x R as symmetric matrices (upper-triangle blocks, "
                                       "half the pairs), blocks of 250 rows"}
        del S, R, cd, emd
    except Exception as e:  # noqa: BLE001
        allpairs = {"error": repr(e)}

    # ---------------- data-parallel training step of the reference's own model (config 2) ----------------
    dp = None
    if not args.skip_dp:
        try:
            dp = dp_training_step(torch, dist, dev, rank, world, barrier, max_over_ranks, _lib)
        except Exception as e:  # noqa: BLE001
            dp = {"error": repr(e)}
        torch.cuda.empty_cache()

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        try:
            ref_cuda = time_reference_cuda(torch, chains)
        except Exception as e:  # noqa: BLE001
            ref_cuda = {"error": repr(e)}

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import torch_cpu as tc
        torch.set_num_threads(os.cpu_count() or 1)
        feats, coords = make_cpu_inputs(torch)
        cpu_chain_step(torch, tc, feats, coords, 2)
        t0 = time.perf_counter()
        pts = cpu_chain_step(torch, tc, feats, coords, B)
        if time.perf_counter() - t0 < 8:
            pts += cpu_chain_step(torch, tc, feats, coords, B)
        dt = time.perf_counter() - t0
        cpu = {"value": pts / dt / 1e6, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"one chain per stage shape (3 of the 6) at B={B}, N={NPTS}: {pts} point-passes in {dt:.1f} s; "
                         "stock PyTorch CPU ops (scatter_add voxelize, grid_sample devoxelize, autograd backward)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG, "clocks": clocks,
                "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "allpairs": allpairs, "pvconv_modules": module,
                "dp_step": dp, "reference_cuda": ref_cuda, "oracle_check": oracle_check, "lib": _lib.build_info()}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = sys.stdout


def _reserve_stdout():
    """stdout carries exactly ONE JSON line: everything else that libraries print there (NCCL's version banner,
    NCCL_DEBUG output, cpp_extension chatter) is sent to stderr by pointing fd 1 at fd 2 for the rest of the run."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-allpairs", action="store_true", help="profiling runs: PVConv primitives only")
    ap.add_argument("--skip-module", action="store_true", help="skip the module-level (cuDNN included) measurement")
    ap.add_argument("--skip-dp", action="store_true", help="skip the data-parallel training step of the reference model")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip timing the reference CUDA kernels (oracle/_ref)")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the warm-up check against the CPU oracle")
    ap.add_argument("--full-emd", action="store_true", help="the full 1000 x 1000 EMD matrix also below 8 GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
