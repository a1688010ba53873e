import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def sbs():
    """The product's host binding (ctypes over libsbsb200.so)."""
    return importlib.import_module("soft-body-simulator_b200")


@pytest.fixture(scope="session")
def scenes():
    return importlib.import_module("soft-body-simulator_b200.scenes")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


# ---- test-only host builds of the product's __host__ __device__ math (tests/host_math.cu) ----
import ctypes as C  # noqa: E402
import shutil  # noqa: E402
import subprocess  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)


def _newer(target, sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


@pytest.fixture(scope="session")
def hostmath():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libhostmath.so")
    src = [os.path.join(HERE, "host_math.cu"), os.path.join(ROOT, "soft-body-simulator_b200/csrc/xpbd_math.cuh"),
           os.path.join(ROOT, "soft-body-simulator_b200/csrc/xpbd_kernels.cuh"),
           os.path.join(ROOT, "soft-body-simulator_b200/csrc/grid_sdf.cuh")]
    if not _newer(lib, src):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                               "-shared", "-o", lib, src[0]])
    L = C.CDLL(lib)
    for f in (L.hostmath_green_project_f64, L.hostmath_green_project_f32):
        f.argtypes = [dp, dp, dp, dp] + [C.c_double] * 6 + [dp]
    for f in (L.hostmath_grid_sample_f64, L.hostmath_grid_sample_f32):
        f.argtypes = [u32p, dp, dp, dp, C.c_int64, C.c_int, dp, dp, dp]
    return L


