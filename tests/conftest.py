"""pytest configuration: `gpu` marker, repo root on sys.path, CPU-side builds."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: large workloads")


def _newer(target, *sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


@pytest.fixture(scope="session")
def built():
    """CPU-side artefacts: oracle, synthetic generator, emulation harness (gcc/g++, seconds)."""
    import __graft_entry__ as g
    g.build_cpu()
    return True


@pytest.fixture(scope="session")
def emul(built):
    import ctypes as C
    so = os.path.join(ROOT, "tests", "emul", "libregemul.so")
    srcs = [os.path.join(ROOT, "tests", "emul", "emul.cpp"),
            os.path.join(ROOT, "registrar_b200", "csrc", "regk_core.cuh"),
            os.path.join(ROOT, "registrar_b200", "csrc", "regk_types.hpp"),
            os.path.join(ROOT, "registrar_b200", "csrc", "regk_decode_core.cuh")]
    if not _newer(so, *srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                               "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    for f in ("emul_paths", "emul_jsons", "emul_dec", "emul_ndigits"):
        getattr(lib, f).restype = C.c_uint32
    lib.emul_build_blob.restype = C.c_int
    lib.emul_parents.restype = C.c_uint64
    lib.emul_services.restype = C.c_uint64
    lib.emul_decode.restype = None
    return lib


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
