"""Build libpercnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The library is several translation units (csrc/*.cu) compiled in parallel and linked into one shared object; every
TU carries its own copy of the __constant__ parameter block (csrc/plan.h)."""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libpercnn_b200.so")
OBJ = os.path.join(_HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    inc = os.path.join(os.path.dirname(_HERE), "include", "percnn_b200.h")
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + [inc, os.path.abspath(__file__)]
    return max(os.path.getmtime(f) for f in files)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    hdr = _headers_mtime()
    os.makedirs(OBJ, exist_ok=True)
    todo, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr, os.path.getmtime(path)):
            todo.append((path, obj))
    if not todo and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(o) for o in objs):
        return OUT
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libpercnn_b200.so")

    def compile_one(job):
        path, obj = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
        return path, subprocess.run(cmd, capture_output=True, text=True)

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for path, res in ex.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {path}:\n" + res.stdout + res.stderr)
            if verbose:
                sys.stderr.write(res.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", OUT] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
