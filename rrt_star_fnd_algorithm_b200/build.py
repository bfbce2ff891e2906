"""Build librrtfnd.so (all CUDA sources, sm_100a) in-tree with nvcc. Used by __graft_entry__.build()."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librrtfnd.so")
SOURCES = ["planner.cu", "ops.cu", "fnd.cu", "host_api.cu"]
HEADERS = ["arith.cuh", "rng.cuh", "grid.cuh", "cull.cuh", "stream.cuh", os.path.join("..", "..", "include", "rrtfnd.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the fp64 contract: never contract a*b+c into an FMA, IEEE division and square root
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    # every file under csrc/ (sources, .cuh, .h) plus the public header and this script
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    env = dict(os.environ)
    env.pop("CC", None)   # this image exports CC=/opt/gcc/bin/gcc, a wrapper nvcc should not pick up
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="-f" in sys.argv, verbose="-v" in sys.argv))
