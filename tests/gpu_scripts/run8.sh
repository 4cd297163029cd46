mkdir -p gpurun_out
for i in 1 2 3; do python -m pytest tests/test_gpu_parity.py -x -q -k "test_trilinear_devoxelize or test_trilinear_walks" 2>&1 | tail -3; done
python - <<'PY'
import os, sys
sys.path[:0] = ["point-cloud-flow-matching_b200", "tests", "tests/golden", "."]
import numpy as np, torch, cases
from pcfm import backend as be, _lib
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
coords, vox, gy, res = cases.devox_case(seed=13, b=2, c=16, n=20000, res=16)
for pdl in (2, 0):
    _lib.set_option("pdl", pdl)
    bad = 0
    for it in range(30):
        try:
            g = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)
            torch.cuda.synchronize()
        except Exception as e:
            print("pdl", pdl, "iter", it, "FAILED", str(e)[:60]); bad += 1; break
    print("pdl", pdl, "bad", bad, flush=True)
    if bad: break
PY
