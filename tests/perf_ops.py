"""Per-op timing table (CUDA events) of the new kernels and, when oracle/_ref is built, of the UNMODIFIED reference
CUDA kernels on the same GPU and inputs.  Test/measurement infrastructure:
    python tests/perf_ops.py [--ref] [--only s1,chamfer] [--kind randn|uniform] [--ops vox_fwd,...] [--opt key=value,...]
Every op is a fresh call (the cell-plan memo is off unless --memo 1); `vox_fwd_planned` / `devox_bwd_planned` time the
same ops on a plan left by an earlier call, i.e. the second PVConv block of a stage.  --opt forwards pcfm_set_option
pairs for A/B runs (e.g. devox_cl=1); PCFM_B200_LIB selects a variant library built by tests/build_variant.sh.
Writes gpurun_out/perf_ops_<kind>.json.  See also tests/perf_sampler.py (config 3)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "point-cloud-flow-matching_b200"))
from pcfm import _lib, backend as be  # noqa: E402


GRAPH = False


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if GRAPH:   # the calls captured once and replayed: device time without the Python / ctypes launch path
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            fn()
        torch.cuda.current_stream().wait_stream(st)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * iters)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--kind", default="randn")
    ap.add_argument("--ops", default="")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3, help="untimed calls before the timed ones (0 for ncu captures)")
    ap.add_argument("--opt", default="", help="extra pcfm_set_option pairs, e.g. devox_rep=0,stage=1")
    ap.add_argument("--plan", type=int, default=0, help="pcfm_set_option('plan', .): 1 = multi-kernel cell-plan build")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5: r x N x C sweep")
    ap.add_argument("--sweep-batch", type=int, default=8, help="batch size of the config-5 sweep")
    ap.add_argument("--memo", type=int, default=0, help="1: leave the cell-plan memo on for every op (repeated calls on "
                    "the same coords then time the *_planned entries); default: off, planned entries timed as own ops")
    ap.add_argument("--graph", action="store_true", help="time every op as a replayed CUDA graph of --iters calls "
                    "(device time without the Python launch path; a *_planned op then includes 1/iters of a plan build)")
    args = ap.parse_args()
    global GRAPH
    GRAPH = args.graph
    only = set(args.only.split(",")) if args.only else None
    ref = None
    if args.ref:
        from oracle import build_ref
        ref = {k: build_ref.load_ref(k) for k in ("pvcnn", "chamfer", "emd", "chamfer6d")}
    dev = "cuda"
    _lib.set_option("plan", args.plan)
    be.plan_memo.enabled = bool(args.memo)
    for kv in filter(None, args.opt.split(",")):
        _lib.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    g = torch.Generator(device=dev).manual_seed(0)
    B, N = 8, 20000
    rows = []
    shapes = [("s1", 128, 32, B, N), ("s2", 256, 16, B, N), ("s3", 256, 8, B, N), ("c1", 64, 32, 2, 2048)]
    if args.sweep:
        shapes = [(f"c5_r{r}_n{n}_c{c}", c, r, args.sweep_batch, n)
                  for r in (16, 32, 64) for n in (10000, 20000, 100000) for c in (64, 256)]
    for name, c, r, b, n in shapes:
        if only and name not in only:
            continue
        feat = torch.randn(b, c, n, device=dev, generator=g)
        if args.kind == "randn":
            coords = torch.randn(b, 3, n, device=dev, generator=g)
        else:  # uniform cube: flat occupancy
            coords = torch.rand(b, 3, n, device=dev, generator=g) * 2 - 1
        nc = coords - coords.mean(2, keepdim=True)
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + 1e-6) + 0.5
        nc = torch.clamp(nc * r, 0, r - 1).contiguous()
        vox = torch.round(nc).to(torch.int32).contiguous()
        grid = torch.randn(b, c, r ** 3, device=dev, generator=g)
        gy = torch.randn(b, c, n, device=dev, generator=g)
        P = be.pvcnn
        out, ind, cnt = P.avg_voxelize_forward(feat, vox, r)
        outs, inds, wgts = P.trilinear_devoxelize_forward(r, True, nc, grid)
        s = r ** 3
        byt = {"vox_fwd": 4 * b * (c * n + 3 * n + c * s + n + s), "vox_bwd": 4 * b * (c * s + n + s + c * n),
               "devox_fwd": 4 * b * (3 * n + c * s + c * n + 16 * n), "devox_bwd": 4 * b * (c * n + 16 * n + c * s)}
        # sparse-aware denominators ("when N < r^3 the grid term dominates ... report both"): the gather-type
        # ops only READ the voxels that hold points (avg_voxelize backward) / the corners around them (devoxelize
        # forward); grids that are WRITTEN (avg_voxelize forward, devoxelize backward) stay dense, zeros included
        occ = int((cnt > 0).sum())                                   # occupied voxels, summed over the batch
        corners = sum(int(torch.unique(inds[i]).numel()) for i in range(b))
        byt_sparse = {"vox_bwd": 4 * (c * occ + b * n + occ + b * c * n),
                      "devox_fwd": 4 * (b * 3 * n + c * corners + b * c * n + b * 16 * n)}
        for k in ("vox_bwd", "devox_fwd"):
            byt_sparse[k + "_planned"] = byt_sparse[k]
        feat0 = feat[:, :0].contiguous()
        ops = {
            "plan_mean": lambda: P.avg_voxelize_forward(feat0, vox, r),     # C == 0: the cell-plan build alone
            "vox_fwd": lambda: P.avg_voxelize_forward(feat, vox, r),
            "vox_bwd": lambda: P.avg_voxelize_backward(grid, ind, cnt),
            "devox_fwd": lambda: P.trilinear_devoxelize_forward(r, True, nc, grid),
            "devox_fwd_eval": lambda: P.trilinear_devoxelize_forward(r, False, nc, grid),
            "devox_bwd": (lambda: P.trilinear_devoxelize_backward_coords(gy, nc, r)) if r <= P.MAX_FAST_DEVOX_BWD_R
            else (lambda: P.trilinear_devoxelize_backward(gy, inds, wgts, r)),
            "devox_bwd_generic": lambda: P.trilinear_devoxelize_backward(gy, inds, wgts, r),
        }
        if not args.memo and r <= P.MAX_FAST_DEVOX_BWD_R:   # the second PVConv block of a stage: plan left by the first

            def planned(fn):
                def run():
                    be.plan_memo.enabled = True
                    fn()
                    be.plan_memo.enabled = False
                return run
            be.plan_memo.enabled = True
            be.plan_memo.clear()
            ops["vox_fwd"](); ops["devox_bwd"]()      # builds the two plans the *_planned ops reuse
            be.plan_memo.enabled = False
            ops["vox_fwd_planned"] = planned(ops["vox_fwd"])
            ops["devox_bwd_planned"] = planned(ops["devox_bwd"])
            ops["devox_fwd_planned"] = planned(ops["devox_fwd"])
            ops["devox_fwd_eval_planned"] = planned(ops["devox_fwd_eval"])
            byt["vox_fwd_planned"], byt["devox_bwd_planned"] = byt["vox_fwd"], byt["devox_bwd"]
            byt["devox_fwd_planned"] = byt["devox_fwd"]
        if args.ops:
            ops = {k: v for k, v in ops.items() if k in args.ops.split(",")}
        for op, fn in ops.items():
            ms = timeit(fn, iters=args.iters, warm=args.warm)
            row = {"shape": name, "C": c, "r": r, "B": b, "N": n, "op": op, "ms": round(ms, 4),
                   "max_cnt": int(cnt.max()), "occupied": int((cnt > 0).sum())}
            if op in byt:
                row["GBps"] = round(byt[op] / ms / 1e6, 1)
                row["frac_of_6545"] = round(byt[op] / ms / 1e6 / 6545.3, 4)
            if op in byt_sparse:
                row["frac_of_6545_sparse_aware"] = round(byt_sparse[op] / ms / 1e6 / 6545.3, 4)
            if not args.no_variants:
                if op == "vox_fwd":
                    _lib.set_option("vox_fwd", 1); row["ms_atomic_variant"] = round(timeit(fn), 4)
                    _lib.set_option("vox_fwd", 2); row["ms_csr_variant"] = round(timeit(fn), 4); _lib.set_option("vox_fwd", 0)
                if op == "devox_bwd_generic":
                    _lib.set_option("devox_bwd", 1); row["ms_atomic_variant"] = round(timeit(fn), 4); _lib.set_option("devox_bwd", 0)
                if op in ("vox_fwd", "devox_fwd"):
                    _lib.set_option("stage", 1); row["ms_plain_stage"] = round(timeit(fn), 4); _lib.set_option("stage", 0)
            if ref and ref["pvcnn"] is not None and op in ("vox_fwd", "vox_bwd", "devox_fwd", "devox_fwd_eval", "devox_bwd",
                                                           "devox_bwd_generic"):
                R = ref["pvcnn"]
                rf = {"vox_fwd": lambda: R.avg_voxelize_forward(feat, vox, r),
                      "vox_bwd": lambda: R.avg_voxelize_backward(grid, ind, cnt),
                      "devox_fwd": lambda: R.trilinear_devoxelize_forward(r, True, nc, grid),
                      "devox_fwd_eval": lambda: R.trilinear_devoxelize_forward(r, False, nc, grid),
                      "devox_bwd": lambda: R.trilinear_devoxelize_backward(gy, inds, wgts, r),
                      "devox_bwd_generic": lambda: R.trilinear_devoxelize_backward(gy, inds, wgts, r)}[op]
                row["ms_reference_cuda"] = round(timeit(rf, iters=3, warm=1), 3)
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.sweep:
        only = {"none"}
    if only and "pointnet" in only:
        # SURVEY 8(a) rows a8-a11 at the config-5 sizes: N points, M = N/4 centres (FPS), U = 32 neighbours, C = 64
        for (b, n, c) in [(8, 4096, 64), (8, 10000, 64), (8, 20000, 64), (8, 100000, 64)]:
            m, u = n // 4, 32
            pts = torch.rand(b, 3, n, device=dev, generator=g) * 2 - 1
            feat = torch.randn(b, c, n, device=dev, generator=g)
            P = be.pvcnn
            R = ref["pvcnn"] if ref else None
            idx = P.furthest_point_sampling(pts, m)
            ctr = P.gather_features_forward(pts, idx)
            radius = 0.1 * 2.0
            nb = P.ball_query(ctr, pts, radius, u)
            grp = P.grouping_forward(feat, nb)
            cfeat = torch.randn(b, c, m, device=dev, generator=g)
            _, i3, w3 = P.three_nearest_neighbors_interpolate_forward(pts, ctr, cfeat)
            gint = torch.randn(b, c, n, device=dev, generator=g)
            cases = {
                "fps": (lambda: P.furthest_point_sampling(pts, m), (lambda: R.furthest_point_sampling(pts, m)) if R else None),
                "ball_query": (lambda: P.ball_query(ctr, pts, radius, u), (lambda: R.ball_query(ctr, pts, radius, u)) if R else None),
                "grouping_fwd": (lambda: P.grouping_forward(feat, nb), (lambda: R.grouping_forward(feat, nb)) if R else None),
                "grouping_bwd": (lambda: P.grouping_backward(grp, nb, n), (lambda: R.grouping_backward(grp, nb, n)) if R else None),
                "gather_fwd": (lambda: P.gather_features_forward(feat, idx), (lambda: R.gather_features_forward(feat, idx)) if R else None),
                "gather_bwd": (lambda: P.gather_features_backward(cfeat, idx, n), (lambda: R.gather_features_backward(cfeat, idx, n)) if R else None),
                "three_nn_fwd": (lambda: P.three_nearest_neighbors_interpolate_forward(pts, ctr, cfeat),
                                 (lambda: R.three_nearest_neighbors_interpolate_forward(pts, ctr, cfeat)) if R else None),
                "three_nn_bwd": (lambda: P.three_nearest_neighbors_interpolate_backward(gint, i3, w3, m),
                                 (lambda: R.three_nearest_neighbors_interpolate_backward(gint, i3, w3, m)) if R else None),
            }
            for op, (fn, rfn) in cases.items():
                if args.ops and op not in args.ops.split(","):
                    continue
                it = 2 if op == "fps" else args.iters
                row = {"op": op, "B": b, "N": n, "M": m, "U": u, "C": c, "ms": round(timeit(fn, iters=it, warm=1), 4)}
                if op == "fps":
                    row["rounds_per_s"] = round(b * m / row["ms"] * 1e3)
                if op == "ball_query":
                    row["neighbours_found_mean"] = round(float((nb != nb[..., :1]).float().sum(-1).mean()) + 1, 2)
                    row["Gevals_per_s_worst_case"] = round(b * m * n / row["ms"] / 1e6, 1)
                    # early exit: a centre stops scanning at its U-th hit (ball_query.cu:36); with fewer hits the tail
                    # of the row repeats the first one and the whole cloud was scanned
                    full = nb[..., -1] != nb[..., 0]
                    scanned = torch.where(full, nb[..., -1].float() + 1, torch.full_like(nb[..., -1], n).float())
                    row["scanned_fraction"] = round(float(scanned.mean()) / n, 4)
                    row["Gevals_per_s_actual"] = round(b * m * n * row["scanned_fraction"] / row["ms"] / 1e6, 1)
                    row["radius"] = radius
                if op == "three_nn_fwd":
                    row["Gevals_per_s"] = round(b * m * n / row["ms"] / 1e6, 1)
                if rfn is not None:
                    row["ms_reference_cuda"] = round(timeit(rfn, iters=2, warm=1), 3)
                rows.append(row); print(json.dumps(row), flush=True)
    if only and "allpairs" in only:
        # config 4 kernels on a block that fills the GPU several times: 296 x 64 cloud pairs (CD), 296 x 8 (EMD)
        S = torch.randn(296, 2048, 3, device=dev, generator=g) * 0.3
        Rr = torch.randn(64, 2048, 3, device=dev, generator=g) * 0.3
        row = {"op": "chamfer_allpairs", "pairs": 296 * 64, "N": 2048,
               "ms": round(timeit(lambda: be.chamfer3d_allpairs(S, Rr), iters=args.iters, warm=args.warm), 3)}
        row["Mpairs_per_s"] = round(row["pairs"] / row["ms"] / 1e3, 4)
        row["Tevals_per_s"] = round(row["pairs"] * 2048 * 2048 / row["ms"] / 1e9, 3)
        rows.append(row); print(json.dumps(row), flush=True)
        row = {"op": "emd_allpairs", "pairs": 296 * 8, "N": 2048,
               "ms": round(timeit(lambda: be.emd_allpairs(S, Rr[:8]), iters=max(1, args.iters // 3), warm=min(args.warm, 1)), 3)}
        row["kpairs_per_s"] = round(row["pairs"] / row["ms"], 3)
        rows.append(row); print(json.dumps(row), flush=True)
    if not only or "chamfer" in only:
        for (b, n, m) in [(32, 2000, 1000), (2, 2048, 2048), (8, 20000, 20000)]:
            a = torch.rand(b, n, 3, device=dev, generator=g); c2 = torch.rand(b, m, 3, device=dev, generator=g)
            d1 = torch.zeros(b, n, device=dev); d2 = torch.zeros(b, m, device=dev)
            i1 = torch.zeros(b, n, dtype=torch.int32, device=dev); i2 = torch.zeros(b, m, dtype=torch.int32, device=dev)
            row = {"op": "chamfer_fwd", "B": b, "N": n, "M": m,
                   "ms": round(timeit(lambda: be.chamfer.forward(a, c2, d1, d2, i1, i2)), 4)}
            row["Gevals_per_s"] = round(2 * b * n * m / row["ms"] / 1e6, 1)
            if ref and ref["chamfer"] is not None:
                row["ms_reference_cuda"] = round(timeit(lambda: ref["chamfer"].forward(a, c2, d1, d2, i1, i2), iters=3, warm=1), 3)
            rows.append(row); print(json.dumps(row), flush=True)
    if not only or "chamfer" in only:
        for (b, n, m) in [(32, 2000, 1000), (2, 2048, 2048), (8, 20000, 20000)]:
            a = torch.rand(b, n, 6, device=dev, generator=g); c2 = torch.rand(b, m, 6, device=dev, generator=g)
            d1 = torch.zeros(b, n, device=dev); d2 = torch.zeros(b, m, device=dev)
            i1 = torch.zeros(b, n, dtype=torch.int32, device=dev); i2 = torch.zeros(b, m, dtype=torch.int32, device=dev)
            row = {"op": "chamfer6d_fwd", "B": b, "N": n, "M": m,
                   "ms": round(timeit(lambda: be.chamfer_6d.forward(a, c2, d1, d2, i1, i2)), 4)}
            row["Gevals_per_s"] = round(2 * b * n * m / row["ms"] / 1e6, 1)
            if ref and ref.get("chamfer6d") is not None:
                row["ms_reference_cuda"] = round(timeit(lambda: ref["chamfer6d"].forward(a, c2, d1, d2, i1, i2), iters=3, warm=1), 3)
            rows.append(row); print(json.dumps(row), flush=True)
    if not only or "emd" in only:
        for (b, n) in [(8, 1024), (32, 2048)]:
            a = torch.rand(b, n, 3, device=dev, generator=g); c2 = torch.rand(b, n, 3, device=dev, generator=g)
            row = {"op": "emd_approxmatch+matchcost", "B": b, "N": n}
            row["ms"] = round(timeit(lambda: be.emd.matchcost_forward(a, c2, be.emd.approxmatch_forward(a, c2)), iters=3, warm=1), 3)
            if ref and ref["emd"] is not None:
                E = ref["emd"]
                row["ms_reference_cuda"] = round(timeit(lambda: E.matchcost_forward(a, c2, E.approxmatch_forward(a, c2)), iters=3, warm=1), 3)
            rows.append(row); print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "sweep_" + args.kind if args.sweep else args.kind
    if args.only and not args.sweep:
        tag += "_" + args.only.replace(",", "-")
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"perf_ops_{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
