"""Build libmte.so (hand-written sm_100a CUDA behind the C ABI of include/mte.h).

    python -m mindtheedge_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(mindtheedge_b200/libmte.so) so it travels to the GPU box with the repo
snapshot; it links cudart statically and has no torch dependency.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.environ.get("MTE_LIB_OUT", os.path.join(HERE, "libmte.so"))  # alternate output for A/B experiments
if "MTE_LIB_OUT" in os.environ:  # one object directory per alternate build (its defines differ)
    OBJ = OBJ + "_" + os.path.splitext(os.path.basename(LIB))[0]
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + os.environ.get("MTE_NVCC_DEFS", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "mte.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr = _newest_header()
    todo, objs = [], []
    for s in sources():
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        src_m = max(os.path.getmtime(os.path.join(CSRC, s)), hdr)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_m:
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
