"""Build ``libb200mp.so`` (the C-ABI + CUDA kernels) for sm_100a with nvcc, in-tree.

    python -m python_motionplanning_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting ``.so`` is git-ignored but travels to the GPU box
with the repository snapshot.  The library links the shared CUDA runtime so that it shares one runtime
(current device, streams) with the process's PyTorch, which is used only for buffers and streams.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libb200mp.so")
SOURCES = ["b200mp_api.cu", "rollout_kernels_f64.cu", "rollout_kernels_f32.cu", "collision_kernels.cu", "misc_kernels.cu", "tracking_kernels.cu", "lattice_kernels.cu"]
HEADERS = ["b200mp_internal.h", "b200mp_math.cuh", "vehicle_rhs.cuh", "rollout_kernels.cuh", "slice_sched.cuh", os.path.join("..", "..", "include", "b200mp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libb200mp.so cannot be built")


def _newest_input() -> float:
    paths = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return max(os.path.getmtime(p) for p in paths + [os.path.abspath(__file__)])


def _compile(nvcc, src, verbose):
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return obj, res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_input():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(BUILD, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(nvcc, s, verbose), SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB, *objs,
           "-Xlinker", f"-rpath={cuda_lib}", "-Xlinker", "--no-undefined"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
