import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "point-cloud-flow-matching_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import cases
from pcfm import backend as be, _lib
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for kw in [dict(seed=13, b=2, c=16, n=20000, res=16), dict(seed=14, b=2, c=16, n=20000, res=8), dict(seed=11, b=2, c=64, n=2048, res=32)]:
    coords, vox, gy, res = cases.devox_case(**kw)
    for step in ("fwd_fresh", "fwd_planned", "bwd_planned", "bwd_fresh"):
        try:
            cc = cu(coords)
            if step == "fwd_fresh":
                be.pvcnn.trilinear_devoxelize_forward(res, True, cc, cu(vox))
            elif step == "fwd_planned":
                be.pvcnn.trilinear_devoxelize_forward(res, True, cc, cu(vox)); torch.cuda.synchronize()
                be.pvcnn.trilinear_devoxelize_forward(res, False, cc, cu(vox))
            elif step == "bwd_planned":
                be.pvcnn.trilinear_devoxelize_forward(res, True, cc, cu(vox)); torch.cuda.synchronize()
                be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cc, res)
            else:
                be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cc, res)
            torch.cuda.synchronize()
            print(kw, step, "ok", flush=True)
        except Exception as e:
            print(kw, step, "FAILED", str(e)[:80], flush=True)
            sys.exit(1)
