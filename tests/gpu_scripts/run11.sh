python - <<'PY' 2>&1 | tail -30
import os, sys, subprocess
code = r'''
import os, sys
sys.path[:0] = ["point-cloud-flow-matching_b200", "tests", "tests/golden", "."]
import numpy as np, torch, cases
from pcfm import backend as be, _lib
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
tri2, pdl, which = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
_lib.set_option("tri2", tri2); _lib.set_option("pdl", pdl)
coords, vox, gy, res = cases.devox_case(seed=13, b=2, c=16, n=20000, res=16)
for it in range(6):
    junk = torch.randint(-2**31, 2**31 - 1, (64 << 20,), dtype=torch.int32, device="cuda"); del junk
    if which == "bwd":
        g = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)
    else:
        g = be.pvcnn.trilinear_devoxelize_forward(res, False, cu(coords), cu(vox))
    torch.cuda.synchronize()
print("ok")
'''
open("/tmp/m.py", "w").write(code)
for tri2 in (0, 1):
    for pdl in (2, 0):
        for which in ("bwd", "fwd"):
            r = subprocess.run([sys.executable, "/tmp/m.py", str(tri2), str(pdl), which], capture_output=True, text=True)
            last = (r.stdout.strip().splitlines() or [""])[-1]
            err = [l for l in r.stderr.splitlines() if "Error" in l][-1:] 
            print("tri2", tri2, "pdl", pdl, which, "->", last, err)
PY
