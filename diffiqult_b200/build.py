"""Build libdiffiqult_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m diffiqult_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "_lib", "libdiffiqult_b200.so")
SOURCES = ["dq_kernels.cu", "dq_batch.cu", "dq_api.cu"]
HEADERS = ["dq_math.cuh", "dq_layout.h", "dq_launch.h", os.path.join("..", "..", "include", "diffiqult_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    return "nvcc"


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(HERE, "_lib", src.replace(".cu", ".o"))
        extra = os.environ.get("DQ_NVCC_EXTRA", "").split()          # e.g. -DDQ_EXPERIMENT_NO_LOOPS (timing experiments)
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
