"""The reference's OWN `models.py` (unmodified, staged by `__graft_entry__.build()` under baseline/_ref/, git-ignored)
running on this repository's kernels: `dropin/` answers its `from modules.pvconv import PVConv` /
`from modules.shared_mlp import SharedMLP` (models.py:9-15).  Test / measurement infrastructure.

    load_models()                     -> the imported reference module (HybridMLP, ContextNet ...)
    heun_sample(net, x, cond, steps)  -> the Heun predictor-corrector loop of train.py:333-341 (2 velocity
                                         evaluations per step), eager
    HeunGraph(net, x, cond, steps)    -> the same loop with ONE captured step replayed `steps` times
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "point-cloud-flow-matching_b200")
STAGED = os.path.join(ROOT, "baseline", "_ref", "models.py")


def staged():
    return os.path.exists(STAGED)


def load_models():
    """Imports baseline/_ref/models.py with dropin/ first on sys.path.  models.py:10-13 would put its own
    third_party/pvcnn first if that directory existed beside it; it is deliberately not staged."""
    if "pcfm_ref_models" in sys.modules:
        return sys.modules["pcfm_ref_models"]
    if not staged():
        raise FileNotFoundError(f"{STAGED} is missing: run `python __graft_entry__.py build` where /root/reference exists")
    for p in (PKG, os.path.join(PKG, "dropin")):
        if p not in sys.path:
            sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("pcfm_ref_models", STAGED)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pcfm_ref_models"] = mod
    spec.loader.exec_module(mod)
    import pcfm.nn
    assert mod.PVConv is pcfm.nn.PVConv and mod.SharedMLP is pcfm.nn.SharedMLP, "models.py did not bind the drop-in modules"
    return mod


def make_hybrid(device, point_dim=6, cond_dim=129, seed=0, **kw):
    """HybridMLP with the defaults of models.py:613-636 (stage_channels (128,256,256), stage_res (32,16,8), 2 blocks
    per stage): point_dim 6 (xyz + rgb), cond = latent 128 + 1 joint value."""
    m = load_models()
    torch.manual_seed(seed)
    net = m.HybridMLP(cond_dim=cond_dim, point_dim=point_dim, **kw).to(device)
    # head_out / the velocity head are zero-initialised (models.py:456, 571-577): give them weights so that outputs
    # and gradients depend on the PVConv path
    with torch.no_grad():
        for p in net.parameters():
            if p.abs().max() == 0 and p.dim() > 1:
                p.normal_(0, 0.02)
    return net


def heun_step(net, x, t0, t1, cond, dt, guidance_scale=0.0):
    """train.py:336-341"""
    v1 = net.guided_velocity(x, t0, cond, guidance_scale=guidance_scale)
    x_hat = x + v1 * dt
    v2 = net.guided_velocity(x_hat, t1, cond, guidance_scale=guidance_scale)
    return x + 0.5 * dt * (v1 + v2)


@torch.no_grad()
def heun_sample(net, x, cond, steps, guidance_scale=0.0):
    """train.py:333-341: t0 = k/steps -> t1 = (k+1)/steps, 2 NFE per step."""
    b = x.shape[0]
    dt = 1.0 / steps
    for k in range(steps):
        t0 = torch.full((b,), k * dt, device=x.device, dtype=x.dtype)
        t1 = torch.full((b,), (k + 1) * dt, device=x.device, dtype=x.dtype)
        x = heun_step(net, x, t0, t1, cond, dt, guidance_scale)
    return x


class HeunGraph:
    """One Heun step captured into a CUDA graph (static x / t0 / t1 buffers), replayed once per step.  Every launcher
    of the library runs on the capturing stream and the host-side memos key on the capture (pcfm_stream_capture_id),
    so the replay rebuilds the cell plans for the moved points."""

    def __init__(self, net, x, cond, steps, guidance_scale=0.0):
        self.steps, self.dt = steps, 1.0 / steps
        self.x = x.clone()
        self.cond = cond
        b = x.shape[0]
        self.t0 = torch.zeros(b, device=x.device, dtype=x.dtype)
        self.t1 = torch.zeros(b, device=x.device, dtype=x.dtype)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):
                heun_step(net, self.x, self.t0, self.t1, cond, self.dt, guidance_scale)
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.x_next = heun_step(net, self.x, self.t0, self.t1, cond, self.dt, guidance_scale)

    @torch.no_grad()
    def sample(self, x):
        self.x.copy_(x)
        for k in range(self.steps):
            self.t0.fill_(k * self.dt)
            self.t1.fill_((k + 1) * self.dt)
            self.graph.replay()
            self.x.copy_(self.x_next)
        return self.x.clone()
