"""Build the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

The reference's hot path is three pybind11/ATen CUDA extensions (plus chamfer_6D, a "next" row of SURVEY 8f).  They are compiled here from the
sources where they lie under /root/reference (nothing is copied into this repo) with torch's
cpp_extension builder for sm_100a, outputs only under oracle/_ref/ (git-ignored, shipped by gpurun):

  oracle/_ref/pvcnn/_pvcnn_backend.so   <- third_party/pvcnn/modules/functional/src/**   (backend.py:6-23)
  oracle/_ref/chamfer/chamfer_3D.so     <- third_party/ChamferDistancePytorch/chamfer3D/ (dist_chamfer_3D.py:12-16)
  oracle/_ref/emd/emd_ext.so            <- third_party/PyTorchEMD/cuda/                  (backend.py:11-26)
  oracle/_ref/chamfer6d/chamfer_6D.so   <- third_party/ChamferDistancePytorch/chamfer6D/ (dist_chamfer_6D.py:16-21)

Flags mirror the reference loaders (pvcnn: -O3 host only; chamfer: defaults; EMD: -O3 --use_fast_math).
Only tests/, __graft_entry__.smoke() and tests/perf scripts may load these; the product never does.
Run:  python oracle/build_ref.py   (needs /root/reference; no GPU needed, ~5 min on 8 cores)
"""
import os
import sys
import time

REF = os.environ.get("PCFM_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(which=None):
    if not os.path.isdir(REF):
        print(f"[oracle/_ref] {REF} not present: skipping (prebuilt files are used as they are)")
        return False
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    jobs = {}
    pv = os.path.join(REF, "third_party/pvcnn/modules/functional/src")
    jobs["pvcnn"] = dict(
        name="_pvcnn_backend",
        extra_cflags=["-O3", "-std=c++17"],
        sources=[os.path.join(pv, f) for f in [
            "ball_query/ball_query.cpp", "ball_query/ball_query.cu",
            "grouping/grouping.cpp", "grouping/grouping.cu",
            "interpolate/neighbor_interpolate.cpp", "interpolate/neighbor_interpolate.cu",
            "interpolate/trilinear_devox.cpp", "interpolate/trilinear_devox.cu",
            "sampling/sampling.cpp", "sampling/sampling.cu",
            "voxelization/vox.cpp", "voxelization/vox.cu",
            "bindings.cpp"]],
    )
    ch = os.path.join(REF, "third_party/ChamferDistancePytorch/chamfer3D")
    jobs["chamfer"] = dict(
        name="chamfer_3D",
        sources=[os.path.join(ch, "chamfer_cuda.cpp"), os.path.join(ch, "chamfer3D.cu")],
    )
    c6 = os.path.join(REF, "third_party/ChamferDistancePytorch/chamfer6D")
    jobs["chamfer6d"] = dict(
        name="chamfer_6D",
        sources=[os.path.join(c6, "chamfer_cuda.cpp"), os.path.join(c6, "chamfer6D.cu")],
    )
    em = os.path.join(REF, "third_party/PyTorchEMD/cuda")
    jobs["emd"] = dict(
        name="emd_ext",
        sources=[os.path.join(em, "emd.cpp"), os.path.join(em, "emd_kernel.cu")],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=["-O3", "--use_fast_math", "-std=c++17",
                           "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "--expt-relaxed-constexpr", "--expt-extended-lambda"],
        extra_include_paths=[em],
    )
    ok = True
    for key, kw in jobs.items():
        if which and key not in which:
            continue
        bdir = os.path.join(OUT, key)
        so = os.path.join(bdir, kw["name"] + ".so")
        if os.path.exists(so):
            print(f"[oracle/_ref] {so} exists")
            continue
        os.makedirs(bdir, exist_ok=True)
        t = time.time()
        try:
            load(build_directory=bdir, verbose=False, is_python_module=False, **kw)
            print(f"[oracle/_ref] built {so} in {time.time() - t:.0f}s")
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[oracle/_ref] FAILED {key}: {e}")
    return ok


def load_ref(key):
    """Import a prebuilt reference extension by file path (tests only). Returns module or None."""
    import importlib.util
    names = {"pvcnn": "_pvcnn_backend", "chamfer": "chamfer_3D", "emd": "emd_ext", "chamfer6d": "chamfer_6D"}
    so = os.path.join(OUT, key, names[key] + ".so")
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(names[key], so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if build(sys.argv[1:] or None) else 1)
