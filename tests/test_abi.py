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


def test_header_is_plain_c99_and_agrees_with_the_python_binding(tmp_path):
    """include/deva_b200.h must compile as C (the boundary a cgo / JNI / ctypes binding sees), link against the
    library, and lay its structs out exactly as devastator_b200/engine.py declares them."""
    import ctypes
    import subprocess
    from devastator_b200 import engine as E
    src = tmp_path / "abi_c.c"
    src.write_text('#include "deva_b200.h"\n#include <stddef.h>\n#include <stdio.h>\nint main(void) {\n'
                   '  printf("%d %zu %zu %zu %zu %zu %zu\\n", deva_b200_abi_version(), sizeof(deva_b200_config),\n'
                   '         sizeof(deva_b200_model_params), sizeof(deva_b200_stats), offsetof(deva_b200_config, mp),\n'
                   '         offsetof(deva_b200_config, window), offsetof(deva_b200_stats, kernel_launches));\n  return 0;\n}\n')
    exe = tmp_path / "abi_c"
    libdir = os.path.dirname(E.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", str(src),
                           f"-L{libdir}", "-ldeva_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [E.ABI_VERSION, ctypes.sizeof(E.Config), ctypes.sizeof(E.ModelParams), ctypes.sizeof(E.Stats),
            E.Config.mp.offset, E.Config.window.offset, E.Stats.kernel_launches.offset]
    assert got == want
