"""Builds libjalla_b200.so (the C-ABI product, nvcc, sm_100a only) and the CPU oracle.

In-tree build so the .so travels to the GPU box with the repo snapshot.
Usage: python -m jalla_b200.build [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "jalla_b200", "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(ROOT, "jalla_b200", "libjalla_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("JB_NVCC_EXTRA", "").split()     # e.g. -DJB_LU32_ABLATIONS for tools/time_lu32.py's profiling variants


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path: str) -> str:
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + ["../../include/jalla_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str) -> tuple[str, str]:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OBJ, src[:-3] + ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    return obj, r.stderr


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp(CSRC)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        results = list(ex.map(_compile, _sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xlinker", "-Bsymbolic", "-Xlinker", "--no-undefined", "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp_file, "w").write(stamp)
    return LIB


def build_oracle() -> str:
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    return os.path.join(ROOT, "oracle", "libjalla_oracle.so")


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    print(build_oracle())
