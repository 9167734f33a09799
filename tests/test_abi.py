"""The C-ABI library loads without a GPU and exports every symbol include/deva_b200.h declares
(no compute call is made here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "deva_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(deva_b200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "devastator_b200", "libdeva_b200.so")
    if not os.path.exists(lib):
        g.build()
    cdll = ctypes.CDLL(lib)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(cdll, s), s
    assert cdll.deva_b200_abi_version() == 1


def test_python_binding_lists_the_same_symbols():
    from devastator_b200 import engine as E
    assert sorted(E.ABI_SYMBOLS) == declared_symbols()


def test_emulation_harness_exports_the_same_abi(emu_api):
    for s in declared_symbols():
        assert hasattr(emu_api.lib, s), s


def test_no_host_symbols_of_the_oracle_in_the_product():
    """The product library must not contain the oracle or any CPU execution path."""
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "devastator_b200", "libdeva_b200.so")], text=True)
    assert "phold_oracle" not in out and "emu_" not in out


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from devastator_b200 import engine as E
    try:
        E.Engine(E.MODEL_PHOLD_TEST, 16)
    except E.EngineError as ex:
        assert ex.code == -2 and "no CPU path" in str(ex)
    else:
        raise AssertionError("engine creation must fail without a CUDA device")
