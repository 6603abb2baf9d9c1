"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        if os.path.basename(h) == "dtfft_b200_mpi.h":
            continue  # header-only adapter for MPI programs: static inline functions, nothing exported
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b(dtfftb?_[a-z0-9_]+)\s*\(", src):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import dtfft_b200

    lib = dtfft_b200.lib()
    names = declared_functions()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_version_and_no_cpu_fallback():
    import dtfft_b200
    from dtfft_b200.kernel import Kernel

    lib = dtfft_b200.lib()
    lib.dtfftb_version.restype = ctypes.c_char_p
    assert b"dtfft_b200" in lib.dtfftb_version()
    import numpy as np
    import pytest

    # host (numpy) buffers are rejected: the product never computes on the CPU
    with pytest.raises(TypeError):
        Kernel().execute(np.zeros(4), np.zeros(4))


def test_library_then_torch_share_one_nccl():
    """Loading libdtfft_b200.so BEFORE torch must not bind the older system libnccl: the loader
    preloads the NCCL that torch ships, so a later `import torch` finds its own symbols."""
    import subprocess
    import sys

    code = ("import dtfft_b200; dtfft_b200.lib(); import torch; "
            "n=[l.split()[-1] for l in open('/proc/self/maps') if 'libnccl' in l]; "
            "assert len(set(n)) == 1, set(n); print('one nccl')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0 and "one nccl" in out.stdout, out.stderr[-2000:]
