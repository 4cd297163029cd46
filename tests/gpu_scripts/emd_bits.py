"""Dumps EMD results (operator: match + cost; all-pairs block: costs) for fixed seeded inputs, to compare two builds of
the library bit for bit:  PCFM_B200_LIB=<lib> python tests/gpu_scripts/emd_bits.py <out.npz>"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "point-cloud-flow-matching_b200"))
from pcfm import _lib, backend as be
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(5)
out = {}
for name, (b, n, m) in {"op_small": (2, 1024, 1024), "op_odd": (3, 777, 1500), "op_big": (4, 2048, 2048)}.items():
    x = torch.rand(b, n, 3, device=dev, generator=g)
    y = torch.rand(b, m, 3, device=dev, generator=g)
    for cl in (1, 0):
        _lib.set_option("emd_cluster", cl)
        match = be.emd.approxmatch_forward(x, y)
        cost = be.emd.matchcost_forward(x, y, match)
        out[f"{name}_cl{cl}_match"] = match.cpu().numpy()
        out[f"{name}_cl{cl}_cost"] = cost.cpu().numpy()
_lib.set_option("emd_cluster", 1)
S = torch.rand(24, 2048, 3, device=dev, generator=g)
R = torch.rand(16, 2048, 3, device=dev, generator=g)
from pcfm import evaluation as ev
out["allpairs"] = ev.pairwise_matrix(S, R, metric="emd").cpu().numpy() if hasattr(ev, "pairwise_matrix") else np.zeros(1)
np.savez(sys.argv[1], **out)
print({k: float(np.asarray(v, dtype=np.float64).sum()) for k, v in out.items()})
