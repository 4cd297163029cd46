"""Drop-in replacements for the reference's three pybind11 extension modules, on top of the C ABI.

    pvcnn   <->  `_pvcnn_backend`   third_party/pvcnn/modules/functional/src/bindings.cpp:10-37
    chamfer <->  `chamfer_3D`       third_party/ChamferDistancePytorch/chamfer3D/chamfer_cuda.cpp:30-33
    emd     <->  `emd_ext`          third_party/PyTorchEMD/cuda/emd.cpp:23-27

Same function names, argument order, return structure and error behaviour (argument checks raise
RuntimeError exactly where the reference's TORCH_CHECKs do, utils.hpp:7-18).  torch is used for device
memory and the current stream only; all arithmetic happens in libpcfm_b200.so.
"""
import threading
import weakref

import torch

from . import _lib
from ._lib import lib, check


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _check_cuda(x, name):
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _check_contig(x, name):
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")


def _check_float(x, name):
    if x.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")


def _check_int(x, name):
    if x.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")


def _chk(x, name, kind):
    _check_cuda(x, name)
    _check_contig(x, name)
    (_check_float if kind == "f" else _check_int)(x, name)


def _version(t):
    """Autograd version counter of a tensor, or None when it has none (tensors created under torch.inference_mode()
    do not track versions): such tensors are simply not memoised."""
    try:
        return t._version
    except RuntimeError:
        return None


class _PlanMemo:
    """Memo of cell plans (include/pcfm_b200.h, "Plan reuse"), keyed on the coordinate tensor.

    The PVConv blocks of one ContextNet stage voxelize / devoxelize at the SAME coords tensor and resolution
    (models.py:385-389; pcfm.nn.Voxelization hands every block of a stage the same (norm_coords, vox_coords) tensor
    objects), so the plan built by the first avg_voxelize_forward serves the second, and the trilinear plan — built
    beside the mean plan when the forward pass announces the float coordinates, else by the first
    trilinear_devoxelize_backward_coords — serves both backward calls.  An entry is valid only for the same tensor
    OBJECT (held by weak reference, so a recycled address cannot alias) with the same autograd version counter, shape,
    resolution, stream and CUDA-graph capture.  One memo for all threads (the backward pass runs on autograd's device
    thread and uses plans the forward thread built), guarded by a lock.  `plan_memo.enabled = False` switches it off;
    `plan_memo.clear()` drops every plan (a write through `.data` or a raw pointer bypasses the version counter).
    """

    CAPACITY = 8  # 3 stages x (mean, trilinear) per step and some slack; ~6 MB per entry at B=8, N=20000

    def __init__(self):
        self.enabled = True
        self.entries = {}  # key -> (weakref(coords), version, payload); insertion order = age
        self.hits = self.misses = 0
        self._lock = threading.Lock()

    @staticmethod
    def _key(kind, coords, r):
        st = torch.cuda.current_stream(coords.device)
        return (kind, id(coords), tuple(coords.shape), int(r), coords.device.index, st.cuda_stream,
                _lib.stream_capture_id(st.cuda_stream))

    def get(self, kind, coords, r):
        if not self.enabled:
            return None
        ver = _version(coords)
        if ver is None:
            return None
        key = self._key(kind, coords, r)
        with self._lock:
            e = self.entries.get(key)
            if e is not None and e[0]() is coords and e[1] == ver:
                self.hits += 1
                return e[2]
            if e is not None:
                del self.entries[key]
            self.misses += 1
        return None

    def put(self, kind, coords, r, payload):
        if not self.enabled:
            return
        ver = _version(coords)
        if ver is None:
            return
        key = self._key(kind, coords, r)
        with self._lock:
            for k in [k for k, e in self.entries.items() if e[0]() is None]:  # coords tensor died
                del self.entries[k]
            self.entries.pop(key, None)
            while len(self.entries) >= self.CAPACITY:
                del self.entries[next(iter(self.entries))]
            self.entries[key] = (weakref.ref(coords), ver, payload)

    def clear(self):
        with self._lock:
            self.entries.clear()


plan_memo = _PlanMemo()
_lib.option_listeners.append(plan_memo.clear)  # an A/B option may change what a workspace holds


class _PvcnnBackend:
    """`_pvcnn_backend` (bindings.cpp:10-37)."""

    # ---- voxelization (vox.cpp:17-43, 54-76) ----
    @staticmethod
    def avg_voxelize_forward(features, coords, resolution, float_coords=None):
        """vox.cpp:17-43.  `float_coords` (extension, optional): the float voxel coordinates the block will devoxelize
        at; when given and no plan is memoised yet, the trilinear plan of the backward pass is built by the same
        kernel as the mean plan (pcfm_avg_voxelize_forward_dual) and memoised for trilinear_devoxelize_backward_coords."""
        _chk(features, "features", "f")
        _chk(coords, "coords", "i")
        b, c, n = features.shape
        r = int(resolution)
        s = r * r * r
        dev = features.device
        out = torch.empty((b, c, s), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            hit = plan_memo.get("mean", coords, r) if n > 0 and c > 0 else None
            if hit is not None:  # same coords tensor as an earlier call: reuse its cell plan, ind and cnt
                ws, ind, cnt = hit
                check(lib.pcfm_avg_voxelize_forward_planned(_ptr(features), b, c, n, r, _ptr(out), _ptr(ws),
                                                            ws.numel(), _stream(features)),
                      "avg_voxelize_forward_planned")
                return [out, ind, cnt]
            ind = torch.empty((b, n), dtype=torch.int32, device=dev)
            cnt = torch.empty((b, s), dtype=torch.int32, device=dev)
            nb = lib.pcfm_avg_voxelize_forward_workspace(b, c, n, r)
            ws = _ws(nb, dev)
            memo_ok = n > 0 and c > 0 and _lib.options.get("vox_fwd", 0) == 0  # generic variants leave no cell plan
            if float_coords is not None:
                _chk(float_coords, "float_coords", "f")
                if tuple(float_coords.shape) != (b, 3, n):
                    raise RuntimeError("float_coords must be [B,3,N]")
            if (float_coords is not None and memo_ok and plan_memo.enabled
                    and plan_memo.get("tri", float_coords, r) is None):
                wst = _ws(lib.pcfm_trilinear_devoxelize_backward_coords_workspace(b, c, n, r), dev)
                check(lib.pcfm_avg_voxelize_forward_dual(_ptr(features), _ptr(coords), _ptr(float_coords), b, c, n, r,
                                                         _ptr(out), _ptr(ind), _ptr(cnt), _ptr(ws), ws.numel(),
                                                         _ptr(wst), wst.numel(), _stream(features)),
                      "avg_voxelize_forward_dual")
                plan_memo.put("tri", float_coords, r, wst)
            else:
                check(lib.pcfm_avg_voxelize_forward(_ptr(features), _ptr(coords), b, c, n, r, _ptr(out), _ptr(ind),
                                                    _ptr(cnt), _ptr(ws), ws.numel(), _stream(features)),
                      "avg_voxelize_forward")
            if memo_ok:
                plan_memo.put("mean", coords, r, (ws, ind, cnt))
        return [out, ind, cnt]

    @staticmethod
    def voxel_coords(centred, scale, resolution):
        """Fused tail of Voxelization.forward (modules/voxelization.py:19-24): centred f32[B,3,N] (coords minus their
        mean), scale f32[B] or None -> (norm_coords f32[B,3,N] in [0,r-1], vox_coords i32[B,3,N])."""
        _chk(centred, "centred", "f")
        b, three, n = centred.shape
        if three != 3:
            raise RuntimeError("centred must be [B,3,N]")
        if scale is not None:
            _chk(scale, "scale", "f")
            if scale.numel() != b:
                raise RuntimeError("scale must have B elements")
        nc = torch.empty_like(centred)
        vox = torch.empty(centred.shape, dtype=torch.int32, device=centred.device)
        with torch.cuda.device(centred.device):
            check(lib.pcfm_voxel_coords(_ptr(centred), _ptr(scale) if scale is not None else None, b, n,
                                        int(resolution), _ptr(nc), _ptr(vox), _stream(centred)), "voxel_coords")
        return nc, vox

    @staticmethod
    def avg_voxelize_backward(grad_y, indices, cnt):
        _chk(grad_y, "grad_y", "f")
        _chk(indices, "indices", "i")
        _chk(cnt, "cnt", "i")
        b, c, s = grad_y.shape
        n = indices.shape[1]
        dev = grad_y.device
        grad_x = torch.empty((b, c, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_avg_voxelize_backward_workspace(b, c, n, s), dev)
            check(lib.pcfm_avg_voxelize_backward(_ptr(grad_y), _ptr(indices), _ptr(cnt), b, c, n, s, _ptr(grad_x),
                                                 _ptr(ws), ws.numel(), _stream(grad_y)), "avg_voxelize_backward")
        return grad_x

    # ---- trilinear devoxelization (trilinear_devox.cpp:18-55, 67-91) ----
    @staticmethod
    def trilinear_devoxelize_forward(r, is_training, coords, features):
        _chk(features, "features", "f")
        _chk(coords, "coords", "f")
        b, c = features.shape[0], features.shape[1]
        n = coords.shape[2]
        r = int(r)
        dev = features.device
        outs = torch.empty((b, c, n), dtype=torch.float32, device=dev)
        if is_training:
            inds = torch.empty((b, 8, n), dtype=torch.int32, device=dev)
            wgts = torch.empty((b, 8, n), dtype=torch.float32, device=dev)
        else:  # trilinear_devox.cpp:45-49: 1-element dummies
            inds = torch.zeros((1,), dtype=torch.int32, device=dev)
            wgts = torch.zeros((1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            # Walk over the trilinear cell plan (pcfm_trilinear_devoxelize_forward_cells): the plan is the one the
            # block's avg_voxelize built beside its mean plan (pcfm.nn.Voxelization announces the float coordinates),
            # and it serves the backward pass as well.  Without a memoised plan the walk pays for a build, which is
            # only worth it for more than a few channels.
            ws = None
            if (c > 0 and n > 0 and _lib.options.get("devox_fwd", 0) == 0
                    and lib.pcfm_trilinear_devoxelize_forward_cells_ok(b, c, n, r)):
                ws = plan_memo.get("tri", coords, r)
                planned = ws is not None
                if not planned and c >= 16:
                    ws = _ws(lib.pcfm_trilinear_devoxelize_backward_coords_workspace(b, c, n, r), dev)
            if ws is not None:
                check(lib.pcfm_trilinear_devoxelize_forward_cells(_ptr(coords), _ptr(features), b, c, n, r,
                                                                  1 if is_training else 0, _ptr(outs),
                                                                  _ptr(inds) if is_training else None,
                                                                  _ptr(wgts) if is_training else None, _ptr(ws),
                                                                  ws.numel(), 1 if planned else 0, _stream(features)),
                      "trilinear_devoxelize_forward_cells")
                if not planned:
                    plan_memo.put("tri", coords, r, ws)
                return [outs, inds, wgts]
            check(lib.pcfm_trilinear_devoxelize_forward(_ptr(coords), _ptr(features), b, c, n, r,
                                                        1 if is_training else 0, _ptr(outs),
                                                        _ptr(inds) if is_training else None,
                                                        _ptr(wgts) if is_training else None, _stream(features)),
                  "trilinear_devoxelize_forward")
        return [outs, inds, wgts]

    @staticmethod
    def trilinear_devoxelize_backward(grad_y, indices, weights, r):
        _chk(grad_y, "grad_y", "f")
        _chk(weights, "weights", "f")
        _chk(indices, "indices", "i")
        b, c, n = grad_y.shape
        r = int(r)
        dev = grad_y.device
        grad_x = torch.empty((b, c, r * r * r), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_trilinear_devoxelize_backward_workspace(b, c, n, r), dev)
            check(lib.pcfm_trilinear_devoxelize_backward(_ptr(grad_y), _ptr(indices), _ptr(weights), b, c, n, r,
                                                         _ptr(grad_x), _ptr(ws), ws.numel(), _stream(grad_y)),
                  "trilinear_devoxelize_backward")
        return grad_x

    MAX_FAST_DEVOX_BWD_R = 256  # 24-bit cell ids in the cell plan (r <= 37: plane in shared memory, above: global atomics)

    @staticmethod
    def trilinear_devoxelize_backward_coords(grad_y, coords, r):
        """Extension (not in the reference module): backward from the float coordinates instead of inds/wgts.
        Same values as trilinear_devoxelize_backward(grad_y, *forward(coords)[1:], r); every point is visited once."""
        _chk(grad_y, "grad_y", "f")
        _chk(coords, "coords", "f")
        b, c, n = grad_y.shape
        r = int(r)
        dev = grad_y.device
        grad_x = torch.empty((b, c, r * r * r), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = plan_memo.get("tri", coords, r) if n > 0 and c > 0 else None
            if ws is not None:
                check(lib.pcfm_trilinear_devoxelize_backward_planned(_ptr(grad_y), b, c, n, r, _ptr(grad_x), _ptr(ws),
                                                                     ws.numel(), _stream(grad_y)),
                      "trilinear_devoxelize_backward_planned")
                return grad_x
            ws = _ws(lib.pcfm_trilinear_devoxelize_backward_coords_workspace(b, c, n, r), dev)
            check(lib.pcfm_trilinear_devoxelize_backward_coords(_ptr(grad_y), _ptr(coords), b, c, n, r, _ptr(grad_x),
                                                                _ptr(ws), ws.numel(), _stream(grad_y)),
                  "trilinear_devoxelize_backward_coords")
            if n > 0 and c > 0:
                plan_memo.put("tri", coords, r, ws)
        return grad_x

    # ---- ball query (ball_query.cpp:6-30) ----
    @staticmethod
    def ball_query(centers_coords, points_coords, radius, num_neighbors):
        _chk(centers_coords, "centers_coords", "f")
        _chk(points_coords, "points_coords", "f")
        b, _, m = centers_coords.shape
        n = points_coords.shape[2]
        u = int(num_neighbors)
        dev = centers_coords.device
        idx = torch.empty((b, m, u), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pcfm_ball_query(_ptr(centers_coords), _ptr(points_coords), b, n, m, float(radius), u,
                                      _ptr(idx), _stream(centers_coords)), "ball_query")
        return idx

    # ---- grouping (grouping.cpp:6-44) ----
    @staticmethod
    def grouping_forward(features, indices):
        _chk(features, "features", "f")
        _chk(indices, "indices", "i")
        b, c, n = features.shape
        m, u = indices.shape[1], indices.shape[2]
        dev = features.device
        out = torch.empty((b, c, m, u), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pcfm_grouping_forward(_ptr(features), _ptr(indices), b, c, n, m, u, _ptr(out),
                                            _stream(features)), "grouping_forward")
        return out

    @staticmethod
    def grouping_backward(grad_y, indices, n):
        _chk(grad_y, "grad_y", "f")
        _chk(indices, "indices", "i")
        b, c = grad_y.shape[0], grad_y.shape[1]
        m, u = indices.shape[1], indices.shape[2]
        n = int(n)
        dev = grad_y.device
        grad_x = torch.empty((b, c, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_grouping_backward_workspace(b, c, n, m, u), dev)
            check(lib.pcfm_grouping_backward(_ptr(grad_y), _ptr(indices), b, c, n, m, u, _ptr(grad_x), _ptr(ws),
                                             ws.numel(), _stream(grad_y)), "grouping_backward")
        return grad_x

    # ---- gather / FPS (sampling.cpp:6-58) ----
    @staticmethod
    def gather_features_forward(features, indices):
        _chk(features, "features", "f")
        _chk(indices, "indices", "i")
        b, c, n = features.shape
        m = indices.shape[1]
        dev = features.device
        out = torch.empty((b, c, m), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pcfm_gather_features_forward(_ptr(features), _ptr(indices), b, c, n, m, _ptr(out),
                                                   _stream(features)), "gather_features_forward")
        return out

    @staticmethod
    def gather_features_backward(grad_y, indices, n):
        _chk(grad_y, "grad_y", "f")
        _chk(indices, "indices", "i")
        b, c = grad_y.shape[0], grad_y.shape[1]
        m = indices.shape[1]
        n = int(n)
        dev = grad_y.device
        grad_x = torch.empty((b, c, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_gather_features_backward_workspace(b, c, n, m), dev)
            check(lib.pcfm_gather_features_backward(_ptr(grad_y), _ptr(indices), b, c, n, m, _ptr(grad_x),
                                                    _ptr(ws), ws.numel(), _stream(grad_y)),
                  "gather_features_backward")
        return grad_x

    @staticmethod
    def furthest_point_sampling(coords, num_samples):
        _chk(coords, "coords", "f")
        b, _, n = coords.shape
        m = int(num_samples)
        dev = coords.device
        idx = torch.zeros((b, m), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_furthest_point_sampling_workspace(b, n, m), dev)
            check(lib.pcfm_furthest_point_sampling(_ptr(coords), b, n, m, _ptr(idx), _ptr(ws), ws.numel(),
                                                   _stream(coords)), "furthest_point_sampling")
        return idx

    # ---- 3-NN interpolation (neighbor_interpolate.cpp:6-65) ----
    @staticmethod
    def three_nearest_neighbors_interpolate_forward(points_coords, centers_coords, centers_features):
        _chk(points_coords, "points_coords", "f")
        _chk(centers_coords, "centers_coords", "f")
        _chk(centers_features, "centers_features", "f")
        b, c, m = centers_features.shape
        n = points_coords.shape[2]
        dev = points_coords.device
        idx = torch.empty((b, 3, n), dtype=torch.int32, device=dev)
        wgt = torch.empty((b, 3, n), dtype=torch.float32, device=dev)
        out = torch.empty((b, c, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pcfm_three_nn_interpolate_forward(_ptr(points_coords), _ptr(centers_coords),
                                                        _ptr(centers_features), b, c, n, m, _ptr(out), _ptr(idx),
                                                        _ptr(wgt), _stream(points_coords)),
                  "three_nearest_neighbors_interpolate_forward")
        return [out, idx, wgt]

    @staticmethod
    def three_nearest_neighbors_interpolate_backward(grad_y, indices, weights, m):
        _chk(grad_y, "grad_y", "f")
        _chk(indices, "indices", "i")
        _chk(weights, "weights", "f")
        b, c, n = grad_y.shape
        m = int(m)
        dev = grad_y.device
        grad_x = torch.empty((b, c, m), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_three_nn_interpolate_backward_workspace(b, c, n, m), dev)
            check(lib.pcfm_three_nn_interpolate_backward(_ptr(grad_y), _ptr(indices), _ptr(weights), b, c, n, m,
                                                         _ptr(grad_x), _ptr(ws), ws.numel(), _stream(grad_y)),
                  "three_nearest_neighbors_interpolate_backward")
        return grad_x


def _chk_chamfer(dim, xyz, per_point, idx, grads=()):
    """The reference reads every tensor through `.data<float>()` / `.data<int>()` (chamfer3D.cu:139-143), which raises
    on a dtype mismatch; shapes are taken from xyz1 / xyz2 only.  Here the raw pointers go to the C ABI, so dtype,
    device, contiguity and the shapes the kernels assume are checked first."""
    x1, x2 = xyz
    _chk(x1, "xyz1", "f")
    _chk(x2, "xyz2", "f")
    if x1.dim() != 3 or x2.dim() != 3 or x1.shape[2] != dim or x2.shape[2] != dim or x1.shape[0] != x2.shape[0]:
        raise RuntimeError(f"chamfer_{dim}D: expected xyz1 [B,N,{dim}] and xyz2 [B,M,{dim}], got "
                           f"{tuple(x1.shape)} {tuple(x2.shape)}")
    for t, x, name in zip(per_point, xyz, ("dist1", "dist2")):
        _chk(t, name, "f")
        if t.numel() != x.shape[0] * x.shape[1]:
            raise RuntimeError(f"chamfer_{dim}D: {name} must have B*N elements")
    for t, x, name in zip(idx, xyz, ("idx1", "idx2")):
        _chk(t, name, "i")
        if t.numel() != x.shape[0] * x.shape[1]:
            raise RuntimeError(f"chamfer_{dim}D: {name} must have B*N elements")
    for t, x, name in zip(grads, xyz, ("gradxyz1", "gradxyz2")):
        _chk(t, name, "f")
        if t.numel() != x.numel():
            raise RuntimeError(f"chamfer_{dim}D: {name} must match its cloud")


class _Chamfer3D:
    """`chamfer_3D` (chamfer_cuda.cpp:30-33): caller-allocated outputs, returns 1 on success / 0 on CUDA error."""

    @staticmethod
    def forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
        _chk_chamfer(3, (xyz1, xyz2), (dist1, dist2), (idx1, idx2))
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        with torch.cuda.device(xyz1.device):
            if _lib.options.get("chamfer_once", 1):   # from ~3e8 evaluations on: every distance evaluated once for both directions
                ws = _ws(lib.pcfm_chamfer3d_forward_workspace(b, n, m), xyz1.device)
                rc = lib.pcfm_chamfer3d_forward_ws(_ptr(xyz1), _ptr(xyz2), b, n, m, _ptr(dist1), _ptr(dist2),
                                                   _ptr(idx1), _ptr(idx2), _ptr(ws), ws.numel(), _stream(xyz1))
            else:
                rc = lib.pcfm_chamfer3d_forward(_ptr(xyz1), _ptr(xyz2), b, n, m, _ptr(dist1), _ptr(dist2), _ptr(idx1),
                                                _ptr(idx2), _stream(xyz1))
        if rc != 0:
            print(f"error in nnd updateOutput: pcfm rc={rc}")
            return 0
        return 1

    @staticmethod
    def backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
        _chk_chamfer(3, (xyz1, xyz2), (graddist1, graddist2), (idx1, idx2), (gradxyz1, gradxyz2))
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        with torch.cuda.device(xyz1.device):
            rc = lib.pcfm_chamfer3d_backward(_ptr(xyz1), _ptr(xyz2), _ptr(graddist1), _ptr(graddist2), _ptr(idx1),
                                             _ptr(idx2), b, n, m, _ptr(gradxyz1), _ptr(gradxyz2), _stream(xyz1))
        if rc != 0:
            print(f"error in nnd get grad: pcfm rc={rc}")
            return 0
        return 1


class _ChamferND:
    """`chamfer_2D` / `chamfer_5D` / `chamfer_6D` (chamfer6D/chamfer_cuda.cpp:30-33): same contract as `chamfer_3D` for
    points with `dim` coordinates."""

    def __init__(self, dim):
        self.dim = dim

    def forward(self, xyz1, xyz2, dist1, dist2, idx1, idx2):
        _chk_chamfer(self.dim, (xyz1, xyz2), (dist1, dist2), (idx1, idx2))
        b, n, d = xyz1.shape
        m = xyz2.shape[1]
        with torch.cuda.device(xyz1.device):
            rc = lib.pcfm_chamfer_nd_forward(_ptr(xyz1), _ptr(xyz2), b, n, m, d, _ptr(dist1), _ptr(dist2),
                                             _ptr(idx1), _ptr(idx2), _stream(xyz1))
        if rc != 0:
            print(f"error in nnd updateOutput: pcfm rc={rc}")
            return 0
        return 1

    def backward(self, xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
        _chk_chamfer(self.dim, (xyz1, xyz2), (graddist1, graddist2), (idx1, idx2), (gradxyz1, gradxyz2))
        b, n, d = xyz1.shape
        m = xyz2.shape[1]
        with torch.cuda.device(xyz1.device):
            rc = lib.pcfm_chamfer_nd_backward(_ptr(xyz1), _ptr(xyz2), _ptr(graddist1), _ptr(graddist2), _ptr(idx1),
                                              _ptr(idx2), b, n, m, d, _ptr(gradxyz1), _ptr(gradxyz2),
                                              _stream(xyz1))
        if rc != 0:
            print(f"error in nnd get grad: pcfm rc={rc}")
            return 0
        return 1


class _EmdExt:
    """`emd_ext` / `emd_cuda_dynamic` (emd.cpp:23-27)."""

    @staticmethod
    def _shapes(xyz1, xyz2):
        b, n, d1 = xyz1.shape
        b2, m, d2 = xyz2.shape
        if b2 != b or d1 != 3 or d2 != 3:  # CHECK_EQ in emd_kernel.cu:176-178
            raise RuntimeError(f"emd: expected xyz1 [B,N,3] and xyz2 [B,M,3], got {tuple(xyz1.shape)} {tuple(xyz2.shape)}")
        if xyz1.dtype != torch.float32 or xyz2.dtype != torch.float32:
            raise RuntimeError("emd: only float32 is supported (the reference's Python wrappers cast to float32)")
        if not (xyz1.is_cuda and xyz2.is_cuda and xyz1.is_contiguous() and xyz2.is_contiguous()):
            raise RuntimeError("emd: inputs must be contiguous CUDA tensors")
        return b, n, m

    @staticmethod
    def approxmatch_forward(xyz1, xyz2):
        b, n, m = _EmdExt._shapes(xyz1, xyz2)
        dev = xyz1.device
        match = torch.empty((b, m, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_emd_approxmatch_workspace(b, n, m), dev)
            check(lib.pcfm_emd_approxmatch_forward(_ptr(xyz1), _ptr(xyz2), b, n, m, _ptr(match), _ptr(ws),
                                                   ws.numel(), _stream(xyz1)), "approxmatch_forward")
        return match

    @staticmethod
    def matchcost_forward(xyz1, xyz2, match):
        b, n, m = _EmdExt._shapes(xyz1, xyz2)
        dev = xyz1.device
        cost = torch.empty((b,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(lib.pcfm_emd_matchcost_workspace(b, n, m), dev)
            check(lib.pcfm_emd_matchcost_forward(_ptr(xyz1), _ptr(xyz2), _ptr(match.contiguous()), b, n, m,
                                                 _ptr(cost), _ptr(ws), ws.numel(), _stream(xyz1)),
                  "matchcost_forward")
        return cost

    @staticmethod
    def matchcost_backward(grad_cost, xyz1, xyz2, match):
        b, n, m = _EmdExt._shapes(xyz1, xyz2)
        dev = xyz1.device
        grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
        grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pcfm_emd_matchcost_backward(_ptr(grad_cost.contiguous()), _ptr(xyz1), _ptr(xyz2),
                                                  _ptr(match.contiguous()), b, n, m, _ptr(grad1), _ptr(grad2),
                                                  _stream(xyz1)), "matchcost_backward")
        return [grad1, grad2]


pvcnn = _PvcnnBackend()
chamfer = _Chamfer3D()
chamfer_2d, chamfer_5d, chamfer_6d = _ChamferND(2), _ChamferND(5), _ChamferND(6)
emd = _EmdExt()


# ---- new capability: all-pairs blocks for the evaluation harness ----
def chamfer3d_allpairs(s, r):
    """cd[i,j] = dist1(S_i,R_j).mean() + dist2(S_i,R_j).mean();  s [ns,N,3], r [nr,M,3] -> [ns,nr]."""
    if not (s.is_cuda and r.is_cuda and s.dtype == torch.float32 and r.dtype == torch.float32):
        raise RuntimeError("chamfer3d_allpairs: float32 CUDA tensors required")
    s, r = s.contiguous(), r.contiguous()
    ns, n, _ = s.shape
    nr, m, _ = r.shape
    out = torch.empty((ns, nr), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        check(lib.pcfm_chamfer3d_allpairs(_ptr(s), _ptr(r), ns, nr, n, m, _ptr(out), _stream(s)),
              "chamfer3d_allpairs")
    return out


def emd_allpairs(s, r):
    """emd[i,j] = earth_mover_distance(S_i, R_j, transpose=False) (emd.py:27-51) without materialising match."""
    if not (s.is_cuda and r.is_cuda and s.dtype == torch.float32 and r.dtype == torch.float32):
        raise RuntimeError("emd_allpairs: float32 CUDA tensors required")
    s, r = s.contiguous(), r.contiguous()
    ns, n, _ = s.shape
    nr, m, _ = r.shape
    out = torch.empty((ns, nr), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        ws = _ws(lib.pcfm_emd_allpairs_workspace(ns, nr, n, m), s.device)
        check(lib.pcfm_emd_allpairs(_ptr(s), _ptr(r), ns, nr, n, m, _ptr(out), _ptr(ws), ws.numel(), _stream(s)),
              "emd_allpairs")
    return out


__all__ = ["pvcnn", "chamfer", "emd", "chamfer3d_allpairs", "emd_allpairs", "_lib"]
