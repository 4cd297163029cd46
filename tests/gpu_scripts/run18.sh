mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "emd or channels_last" 2>&1 | tail -4
python - <<'PY'
import sys, json, torch
sys.path[:0] = ["point-cloud-flow-matching_b200", "tests", "."]
from pcfm import backend as be, _lib
from oracle import build_ref
E = build_ref.load_ref("emd")
def timeit(fn, it=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
g = torch.Generator(device="cuda").manual_seed(0)
rows = []
for (b, n) in [(2, 1024), (8, 1024), (8, 2048), (32, 2048)]:
    a = torch.rand(b, n, 3, device="cuda", generator=g); c = torch.rand(b, n, 3, device="cuda", generator=g)
    row = {"op": "emd approxmatch + matchcost", "B": b, "N": n}
    f = lambda: be.emd.matchcost_forward(a, c, be.emd.approxmatch_forward(a, c))
    row["ms_cluster"] = round(timeit(f), 3)
    _lib.set_option("emd_cluster", 0); row["ms_one_cta"] = round(timeit(f), 3); _lib.set_option("emd_cluster", 1)
    if E is not None:
        row["ms_reference_cuda"] = round(timeit(lambda: E.matchcost_forward(a, c, E.approxmatch_forward(a, c)), 3), 3)
    row["speedup_vs_one_cta"] = round(row["ms_one_cta"] / row["ms_cluster"], 2)
    rows.append(row); print(json.dumps(row), flush=True)
json.dump(rows, open("gpurun_out/perf_emd_cluster.json", "w"), indent=1)
PY
