PCFM_B200_LIB=point-cloud-flow-matching_b200/lib/var/libpcfm_chk.so python - <<'PY' 2>&1 | sort | uniq -c | sort -rn | head -30
import os, sys
sys.path[:0] = ["point-cloud-flow-matching_b200", "tests", "tests/golden", "."]
import numpy as np, torch, cases
from pcfm import backend as be, _lib
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
coords, vox, gy, res = cases.devox_case(seed=13, b=2, c=16, n=20000, res=16)
for it in range(3):
    junk = torch.randint(-2**31, 2**31 - 1, (64 << 20,), dtype=torch.int32, device="cuda")
    del junk
    g = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)
    torch.cuda.synchronize()
print("done")
PY
