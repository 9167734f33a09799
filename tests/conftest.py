import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _newer(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    if any(os.path.getmtime(s) > t for s in srcs):
        return True
    # a sanitizer build left behind by a debugging session cannot be loaded into python: rebuild it
    try:
        deps = subprocess.run(["ldd", target], capture_output=True, text=True).stdout
    except OSError:
        deps = ""
    return any(x in deps for x in ("libasan", "libtsan", "libubsan"))


@pytest.fixture(scope="session")
def emu_api():
    """The host emulation harness (tests/emu): same window code as the kernel, compiled for the CPU."""
    import ctypes
    from devastator_b200 import engine as E
    so = os.path.join(ROOT, "tests", "emu", "_build", "libdeva_b200_emu.so")
    csrc = os.path.join(ROOT, "devastator_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "emu", "emu_engine.cpp"), os.path.join(ROOT, "include", "deva_b200.h")] + \
        [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h", ".inc"))]
    # DEVA_EMU_LIB: a build of the same harness made elsewhere, e.g. with -fsanitize=address,undefined (then run pytest
    # under LD_PRELOAD=$(gcc -print-file-name=libasan.so)); tools/emu_sanitized.sh does both
    if os.environ.get("DEVA_EMU_LIB"):
        so = os.environ["DEVA_EMU_LIB"]
    elif _newer(so, srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-w",
                               srcs[0], "-o", so])
    api = E.Api(ctypes.CDLL(so))
    api.lib.emu_set_shuffle.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    api.lib.emu_set_shuffle.restype = None
    api.lib.emu_enable_peer_delivery.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    api.lib.emu_enable_peer_delivery.restype = None
    api.lib.emu_drain_multi.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_uint64, ctypes.c_int32,
                                        ctypes.POINTER(ctypes.c_uint64)]
    return api


@pytest.fixture(scope="session")
def cuda_api():
    """The product library.  Fails loudly when it is missing or no GPU is visible - no fallback."""
    from devastator_b200 import engine as E
    return E.load()


ENGINE_KINDS = ["tile", "lp"]


@pytest.fixture(params=ENGINE_KINDS)
def engine_kind(request, monkeypatch):
    """Both engines behind the C ABI: "tile" (default: calendar of per-tile event lists, epochs closed on
    the device) and "lp" (the first engine: one thread per LP over per-LP pools)."""
    monkeypatch.setenv("DEVA_B200_ENGINE", request.param)
    return request.param
