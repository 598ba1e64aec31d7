"""pytest configuration.  `-m "not gpu"` runs here (oracle vs golden vectors, host logic, C-ABI symbol check,
kernel logic through the host-emulation test build); `-m gpu` runs the parity tests proper on a B200."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_SO = os.path.join(ROOT, "tests", "_emu", "libmm_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        pyoracle.build(ref=os.path.isdir("/root/reference/src"))
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref_harness():
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return pyoracle.RefHarness()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_small_internals.npz"))


@pytest.fixture(scope="session")
def small_workload(tmp_path_factory):
    from tests.golden.make_golden import small_workload as sw
    d = str(tmp_path_factory.mktemp("small"))
    db, fa, fq, names, reads = sw(d)
    return {"dir": d, "db": db, "fa": fa, "fq": fq, "names": names, "reads": reads}


def build_emu():
    """g++ -DMM_HOST_EMU build of the kernel sources (test infrastructure; see mm_platform.h)."""
    src = os.path.join(ROOT, "metamaps_b200", "csrc", "mm_lib.cu")
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    deps = [os.path.join(ROOT, "metamaps_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "metamaps_b200", "csrc"))]
    deps.append(os.path.join(ROOT, "include", "metamaps_b200.h"))
    if os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(d) for d in deps if os.path.isfile(d)):
        return EMU_SO
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-x", "c++", "-DMM_HOST_EMU",
                    "-I" + os.path.join(ROOT, "include"), src, "-o", EMU_SO], check=True)
    return EMU_SO


EMU_HOST = os.path.join(ROOT, "tests", "_emu", "metamaps_emu")


def build_emu_host():
    """The C++ host linked against the host-emulation library (CPU test tier only)."""
    from metamaps_b200 import build as b
    build_emu()
    return b.build_host(os.path.dirname(EMU_SO), "mm_emu", EMU_HOST)


@pytest.fixture(scope="session")
def emu_ctx():
    from metamaps_b200 import capi
    lib = capi.load(build_emu())
    assert b"HOST EMULATION" in lib.mm_version()
    return capi.Context(0, lib)


@pytest.fixture(scope="session")
def gpu_ctx():
    from metamaps_b200 import capi
    lib = capi.load()           # the nvcc build; raises if it is missing
    assert b"sm_100a" in lib.mm_version()
    return capi.Context(0, lib)
