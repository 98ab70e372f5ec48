"""Build ``liblhvi.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python lifted-hybrid-variational-inference_b200/build.py [--force] [--verbose]

Every ``csrc/*.cu`` is compiled to an object file in parallel, then linked into
``liblhvi.so`` next to this file; ``csrc/host/lhvi_lift.cpp`` (host-side lifting passes, no CUDA)
is compiled with g++ into ``liblhvi_lift.so``.  nvcc cross-compiles without a GPU; the resulting ``.so``
is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "csrc", "_obj")
LIB_PATH = os.path.join(PKG_DIR, "liblhvi.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIFT_SRC = os.path.join(CSRC, "host", "lhvi_lift.cpp")
LIFT_LIB = os.path.join(PKG_DIR, "liblhvi_lift.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
] + os.environ.get("LHVI_NVCC_EXTRA", "").split()      # e.g. -DLHVI_RUN_BLOCKS=3 for a tuning experiment (then --force)


def nvcc_path() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; liblhvi.so cannot be built")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def dependencies():
    return sources() + headers() + [os.path.abspath(__file__)]


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) \
        + [h for h in glob.glob(os.path.join(INCLUDE, "*.h")) if os.path.basename(h) != "lhvi_lift.h"]


_INCLUDE_RE = None


def includes_of(path, seen=None):
    """Transitive closure of the quoted #include files of ``path`` (csrc/ and include/), so that a
    translation unit is rebuilt only when a header it actually uses changed."""
    import re
    global _INCLUDE_RE
    if _INCLUDE_RE is None:
        _INCLUDE_RE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
    seen = set() if seen is None else seen
    try:
        text = open(path).read()
    except OSError:
        return seen
    for name in _INCLUDE_RE.findall(text):
        for base in (os.path.dirname(path), CSRC, INCLUDE):
            cand = os.path.normpath(os.path.join(base, name))
            if os.path.exists(cand):
                if cand not in seen:
                    seen.add(cand)
                    includes_of(cand, seen)
                break
    return seen


def up_to_date() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    built = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(p) <= built for p in dependencies())


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    if os.path.exists(obj):
        newest = max(os.path.getmtime(p) for p in [src] + sorted(includes_of(src)) + [os.path.abspath(__file__)])
        if os.path.getmtime(obj) >= newest:
            return obj
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build_lift(force: bool = False) -> str:
    """``liblhvi_lift.so``: the host-side C++ lifting passes (include/lhvi_lift.h), plain g++."""
    deps = [LIFT_SRC, os.path.join(INCLUDE, "lhvi_lift.h")]
    if not force and os.path.exists(LIFT_LIB) and all(os.path.getmtime(p) <= os.path.getmtime(LIFT_LIB) for p in deps):
        return LIFT_LIB
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        raise RuntimeError("g++ not found; liblhvi_lift.so cannot be built")
    cmd = [cxx, "-O3", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", INCLUDE,
           LIFT_SRC, "-o", LIFT_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed on {LIFT_SRC}:\n{res.stdout}\n{res.stderr}")
    return LIFT_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    build_lift(force)
    if not force and up_to_date():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ_DIR, "*.o")):
            os.remove(f)
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose), sources()))
    cmd = [nvcc_path(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
