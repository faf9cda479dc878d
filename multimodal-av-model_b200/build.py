"""Build libavctc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python multimodal-av-model_b200/build.py [--force]

One object per .cu (compiled in parallel), linked into csrc/libavctc_b200.so.  The .so is git-ignored
but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libavctc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-I", os.path.join(os.path.dirname(HERE), "include")]
# --use_fast_math would also flush fp32 denormals and relax division; the CTC kernels use explicit
# ex2/lg2.approx and need IEEE division nowhere, but keep it OFF for bit-stable host/device parity.
FLAGS.remove("--use_fast_math")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "avctc_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hm = _deps_mtime()
    todo = []
    for s in srcs:
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(os.path.join(CSRC, s)), hm):
            todo.append(s)
    if todo:
        if verbose:
            print("nvcc:", " ".join(todo), flush=True)
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(_compile, todo))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
