mkdir -p gpurun_out
for t in 2 3; do echo "tri2=$t"; python - <<PY 2>&1 | tail -4
import sys, pytest
sys.path[:0] = ["point-cloud-flow-matching_b200", "tests", "tests/golden", "."]
from pcfm import _lib
_lib.set_option("tri2", $t)
import pcfm.backend
orig = _lib.set_option
def patched(k, v):
    orig(k, $t if k == "tri2" else v)
_lib.set_option = patched
sys.exit(pytest.main(["tests/test_gpu_parity.py", "-x", "-q", "-k", "test_trilinear_devoxelize and b2c16n20000r16"]))
PY
done
python - <<'PY'
import os, sys
sys.path[:0] = ["point-cloud-flow-matching_b200", "tests", "tests/golden", "."]
import numpy as np, torch, cases
from pcfm import backend as be, _lib
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
coords, vox, gy, res = cases.devox_case(seed=13, b=2, c=16, n=20000, res=16)
bad = 0
for it in range(10):
    junk = torch.randint(-2**31, 2**31 - 1, (64 << 20,), dtype=torch.int32, device="cuda")   # 256 MB of garbage
    del junk
    try:
        g = be.pvcnn.trilinear_devoxelize_backward_coords(cu(gy), cu(coords), res)
        torch.cuda.synchronize()
    except Exception as e:
        print("garbage-memory iter", it, "FAILED", str(e)[:60]); bad += 1; break
print("garbage-memory bad", bad, flush=True)
PY
