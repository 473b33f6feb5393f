"""Build libmsed_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

The library is several translation units -- msed.cu (host side, C ABI, controllers, helper kernels) and one
msed_tu_*.cu per family of stepping kernels -- compiled in parallel and linked into one shared object."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "..", "build", "obj")
LIB = os.path.join(HERE, "libmsed_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    hs.append(os.path.join(HERE, "..", "include", "msed.h"))
    return hs


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources() + _headers())


def _compile(src: str, extra, verbose: bool, log: dict) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(f) for f in [src] + _headers())
    if os.path.exists(obj) and os.path.getmtime(obj) > newest and not extra and not verbose:
        return obj
    cmd = [NVCC, *FLAGS, *extra, "-c", "-o", obj, src]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log[src] = res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """``defines``: extra -D switches (kernel variants for the sweeps in tools/); ``out``: library path."""
    lib = out or LIB
    if not force and not defines and out is None and not needs_build():
        return lib
    os.makedirs(OBJ, exist_ok=True)
    if force or defines:
        for f in os.listdir(OBJ):
            os.unlink(os.path.join(OBJ, f))
    extra = [f"-D{d}" for d in defines]
    log: dict = {}
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, extra, verbose, log), sources()))
    res = subprocess.run([NVCC, *FLAGS[:2], "-shared", "-o", lib, *objs, "-ldl"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if defines:   # objects of a variant build must not be mistaken for the default ones
        for f in os.listdir(OBJ):
            os.unlink(os.path.join(OBJ, f))
    if verbose:
        print("\n".join(log.values()))
    return lib
