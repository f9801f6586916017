"""
In-tree build of the CUDA library ``heybuddy_b200/_lib/libheybuddy_b200.so`` for sm_100a.

``nvcc`` cross-compiles without a GPU; the resulting ``.so`` is git-ignored but travels to
the GPU box with the repo snapshot.  Invoked by ``__graft_entry__.build()`` and, as a
convenience, by ``python -m heybuddy_b200.build``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "_lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libheybuddy_b200.so")
INCLUDE = os.path.abspath(os.path.join(PKG_DIR, "..", "include"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
] + os.environ.get("HB_NVCC_EXTRA", "").split()      # e.g. -DHB_TCG_PLANE_SKEW=32 for the experiments in embed_tcg.cu


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the heybuddy_b200 CUDA library cannot be built")
    return exe


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(INCLUDE, "heybuddy_b200.h")
    ]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha256"
    digest = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj
    cmd = [nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(f"[heybuddy_b200.build] compiled {os.path.basename(src)}\n")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return obj


def build(verbose: bool = True, force: bool = False) -> str:
    """Compile every ``csrc/*.cu`` for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcuda"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(f"[heybuddy_b200.build] linked {LIB_PATH}\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
