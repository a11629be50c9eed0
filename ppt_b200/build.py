"""In-tree build of libppt_b200.so (sm_100a only).

    python -m ppt_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so
it travels with the source tree (it is git-ignored, not pip-installed).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libppt_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# Geometry kernels reproduce the reference's fp32 rounding with explicit _rn
# intrinsics; -fmad=false is belt and braces for anything written with plain operators.
PER_FILE = {
    "fps.cu": ["-fmad=false"],
    "knn.cu": ["-fmad=false"],
    "ball_query.cu": ["-fmad=false"],
    "gather.cu": ["-fmad=false"],
    "interp.cu": ["-fmad=false"],
    "graph_feature.cu": ["-fmad=false"],
    "edge_conv.cu": [],
}


def _nvcc():
    return os.environ.get("NVCC", "nvcc")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "ppt_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), _deps_mtime()):
        return obj, False
    cmd = [_nvcc()] + ARCH + COMMON + PER_FILE.get(src, []) + ["-c", path, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), _sources()))
    objs = [o for o, _ in results]
    if force or any(changed for _, changed in results) or not os.path.exists(LIB):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fvisibility=hidden"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
