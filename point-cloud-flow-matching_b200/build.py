"""Build libpcfm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

  python point-cloud-flow-matching_b200/build.py [--force] [--verbose]

Output: point-cloud-flow-matching_b200/lib/libpcfm_b200.so  (git-ignored; shipped to the GPU box by gpurun).
No torch dependency: the library exposes only the extern "C" symbols declared in include/pcfm_b200.h.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libpcfm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "pcfm_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


# emd.cu mirrors a reference translation unit that is built with --use_fast_math (PyTorchEMD/backend.py:20):
# denormals are flushed there, so they are flushed here (approximate exp / division are explicit intrinsics).
EXTRA = {"emd.cu": ["-ftz=true"]}


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + EXTRA.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    hm = _headers_mtime()
    todo, objs = [], []
    for s in _sources():
        obj = os.path.join(OBJDIR, s[:-3] + ".o")
        objs.append(obj)
        sm = max(os.path.getmtime(os.path.join(CSRC, s)), hm)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, obj, rc, log in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose or rc:
                    print(f"--- {src}\n{log}")
                if rc:
                    raise RuntimeError(f"nvcc failed on {src}")
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-Xcompiler", "-fPIC", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
