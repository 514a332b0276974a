"""Build recipe of the CUDA extension: nvcc, sm_100a only, in-tree output.

    python -m dd4ml_b200.build [--force] [-v]

Each ``csrc/k_*.cu`` is compiled to an object in parallel and linked into
``dd4ml_b200/libdd4ml_b200.so`` (git-ignored, but it travels to the GPU box).
No ``--use_fast_math``: parity with the reference needs IEEE division / sqrt.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "libdd4ml_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "k_*.cu")))


def headers() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(HERE, "..", "include", "dd4ml_b200.h")]


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("dd4ml_b200: nvcc not found; the CUDA extension cannot be built")


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    deps = [src] + headers()
    if os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
        return obj
    cmd = [nvcc_path(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return obj


def up_to_date() -> bool:
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in sources() + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(f)
    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), sources()))
    subprocess.check_call([nvcc_path(), *ARCH, "--shared", "-Xcompiler", "-fPIC", "-o", OUT, *objs])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
