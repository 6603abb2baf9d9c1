"""C and C++ callers of the drop-in boundary, compiled with the system gcc / g++ against
include/*.h and libdtfft_b200.so -- the way a maintainer of a C/C++ application of the reference
would consume it (reference tests: tests/c/*.c, tests/c/*_cxx.cpp).
  * tests/c/api_host.c   plain C, no GPU: validation order, error codes, config, dry plans;
  * tests/c/api_gpu.cpp  C++ through include/dtfft_b200.hpp on a GPU: bit-exact transposes, cuFFT
                          R2C round trip, graph replay, exceptions."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c")
OUT = os.path.join(SRC, "_build")
LIBDIR = os.path.join(ROOT, "dtfft_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _env():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = LIBDIR + os.pathsep + os.path.join(CUDA, "lib64") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def _build(name):
    import dtfft_b200  # noqa: F401  (raises if the library has not been built)

    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name.split(".")[0])
    if name.endswith(".c"):
        cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
               os.path.join(SRC, name), "-o", exe, "-L" + LIBDIR, "-ldtfft_b200"]
    else:
        cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
               "-I" + os.path.join(CUDA, "include"), os.path.join(SRC, name), "-o", exe, "-L" + LIBDIR, "-ldtfft_b200",
               "-L" + os.path.join(CUDA, "lib64"), "-lcudart"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=_env(), timeout=300)
    assert out.returncode == 0, out.stderr[-4000:]
    return exe


def test_c_host_program():
    exe = _build("api_host.c")
    out = subprocess.run([exe], capture_output=True, text=True, env=_env(), timeout=120)
    assert out.returncode == 0 and "api_host OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_mpi_adapter_header():
    """include/dtfft_b200_mpi.h (MPI_Comm -> dtfftb_comm_t) builds warning-free and works, against a
    single-process stand-in for <mpi.h> (the image has no MPI)."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "api_mpi_adapter")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(SRC, "mpi_stub"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(SRC, "api_mpi_adapter.c"), "-o", exe, "-L" + LIBDIR,
           "-ldtfft_b200"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=_env(), timeout=300)
    assert out.returncode == 0, out.stderr[-4000:]
    out = subprocess.run([exe], capture_output=True, text=True, env=_env(), timeout=120)
    assert out.returncode == 0 and "api_mpi_adapter OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_fastdiv_host_program():
    """The work-item decoder's division by run-time constants is exact over its whole domain."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "fastdiv_host")
    out = subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(SRC, "fastdiv_host.cpp"), "-o", exe],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad 0" in out.stdout, out.stdout[-2000:]


def test_cxx_wrapper_compiles():
    """include/dtfft_b200.hpp + the GPU program build warning-free with -Wall -Wextra -Werror."""
    assert os.path.exists(_build("api_gpu.cpp"))


@pytest.mark.gpu
def test_cxx_gpu_program(cuda):
    exe = _build("api_gpu.cpp")
    out = subprocess.run([exe], capture_output=True, text=True, env=_env(), timeout=300)
    assert out.returncode == 0 and "api_gpu OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
