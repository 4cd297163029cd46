"""FPS: one CTA per cloud against the 8-CTA cluster kernel at small clouds (where the crossover is)."""
import os, sys, json, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "point-cloud-flow-matching_b200"))
from pcfm import _lib, backend as be
dev = torch.device("cuda")
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for n in (512, 1024, 2048, 4096, 6144, 8192):
    pts = torch.rand(8, 3, n, device=dev) * 2 - 1
    m = n // 4
    row = {"N": n, "M": m}
    for opt, name in ((0, "cluster_or_default"), (1, "one_cta")):
        _lib.set_option("fps", opt)
        row[name] = round(timeit(lambda: be.pvcnn.furthest_point_sampling(pts, m)), 4)
    _lib.set_option("fps", 0)
    print(json.dumps(row), flush=True)
