import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "point-cloud-flow-matching_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import refmodel
import test_gpu_refmodel as T
from oracle import build_ref
import pcfm.functional as PF
from pcfm import backend
ref = build_ref.load_ref("pvcnn")
dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cudnn.deterministic = True; torch.backends.cuda.matmul.allow_tf32 = False
refmodel.load_models()
net = refmodel.make_hybrid(dev).train()
x, t, cond = T._inputs(2, 4096, dev, seed=7)
names = [n for n, _ in net.named_parameters()]
params = [p for _, p in net.named_parameters()]
def run():
    net.zero_grad(set_to_none=True)
    for m in net.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm): m.reset_running_stats()
    v = net(x, t, cond)
    v.square().mean().backward()
    return v.detach().clone(), [None if p.grad is None else p.grad.detach().clone() for p in params]
v1, g1 = run()
v1b, g1b = run()
ours = PF._backend
PF._backend = T._ReferenceBackend(ref, ours); backend.plan_memo.clear()
v2, g2 = run()
v2b, g2b = run()
PF._backend = ours
def rel(a, b): return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
print("velocity ours vs ref", rel(v1, v2), "ours run-to-run", rel(v1, v1b), "ref run-to-run", rel(v2, v2b))
rows = []
for n, a, a2, b, b2 in zip(names, g1, g1b, g2, g2b):
    if a is None: continue
    rows.append((rel(a, b), rel(a, a2), rel(b, b2), n, float(b.abs().max())))
rows.sort(reverse=True)
for r in rows[:25]: print("%.3e ours-vs-ref | %.3e ours r2r | %.3e ref r2r | %s  max|g| %.3e" % r)
print("median", sorted(r[0] for r in rows)[len(rows)//2])
