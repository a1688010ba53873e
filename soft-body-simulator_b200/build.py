"""Build libsbsb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libsbsb200.so")
SOURCES = ["csrc/sbs_b200.cu", "csrc/scene_build.cpp"]
HEADERS = ["csrc/scene_build.h", "csrc/xpbd_math.cuh", "csrc/xpbd_kernels.cuh", "csrc/xpbd_resident.cuh", "csrc/bvh.cuh", "csrc/grid_sdf.cuh",
           "../include/sbs_b200.h"]
# -ftz=true: denormals flush to zero in fp32 (they carry no information in this path); division and square root
# stay IEEE (-prec-div / -prec-sqrt at their defaults): the full -use_fast_math build failed the fp32 parity of
# an ill-conditioned scene (profiles/r02_summary.md)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ftz=true",
              "-Xcompiler", "-fPIC,-Wall", "-shared"]


def source_id():
    """Hash of the sources the library is built from: what bench.py and profiles/ stamp their records with (the
    binary itself is rebuilt wherever the sources are newer, see stale())."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        h.update(f.encode())
        h.update(open(os.path.join(HERE, f), "rb").read())
    return h.hexdigest()[:16]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
