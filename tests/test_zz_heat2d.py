"""BASELINE config 5's wording - a 2-D heat stencil with nearest-neighbour events - does not exist in the reference
(SURVEY section 8d); tests/csrc/heat2d.cxx is that app written against the reference's public pdes API (this repo's own
test model, in the style of test/stencil.cxx: it checks itself against a serial recomputation, bit for bit).  Its handlers
are host C++, so it runs through the host executor over the pending pool; the goldens are what the SAME source prints on
the unmodified reference runtime (tests/golden/make_heat2d_goldens.py).
(The file sorts last on purpose: it is the newest GPU test of the suite.)"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "dropin", "_bin")
with open(os.path.join(ROOT, "tests", "golden", "heat2d_goldens.json")) as _f:
    GOLDEN = json.load(_f)


def check(exe, only=None):
    for tag, g in GOLDEN.items():
        if only and tag not in only:
            continue
        out = subprocess.run([exe], env=dict(os.environ, **g["env"]), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:]
        assert out.stdout.strip().splitlines() == g["lines"], (tag, out.stdout[-1000:], g["lines"])


@pytest.mark.parametrize("R", [1, 4])
def test_heat2d_host_handlers_emu(emu_api, tmp_path, R):
    exe = str(tmp_path / f"heat2d_R{R}")
    emu_dir = os.path.join(ROOT, "tests", "emu", "_build")
    subprocess.check_call([
        "g++", "-std=c++14", "-O2", "-pthread", "-w", f"-DDEVA_THREAD_N={R}", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1", "-DDEBUG=0", "-DNDEBUG=1",
        f"-I{ROOT}/include", os.path.join(ROOT, "tests", "csrc", "heat2d.cxx"), f"{ROOT}/devastator_b200/csrc/host_runtime.cpp",
        f"-L{emu_dir}", "-ldeva_b200_emu", f"-Wl,-rpath,{emu_dir}", "-o", exe])
    check(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 4])
def test_heat2d_host_handlers_on_gpu(R):
    exe = os.path.join(BIN, f"heat2d_R{R}")
    assert os.path.exists(exe), "dropin/_bin missing: run __graft_entry__.build()"
    check(exe, ("default", "ragged_thirds", "one") if R == 1 else ("thirds", "ragged", "two", "wide"))
