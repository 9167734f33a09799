"""The drop-in boundary in C++: the reference's OWN app sources, unchanged, compiled against
include/devastator/*.hxx (dropin/build_dropin.py).

 - CPU suite: when /root/reference is present (build container) test/phold.cxx is compiled against
   the header mirror and linked with the host EMULATION of the engine ABI, which checks the C++
   veneer (rank threads, collectives, state gather/scatter, root staging, rewind) without a GPU.
 - GPU suite: the prebuilt product binaries (linked with libdeva_b200.so) run on the B200."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEVA_REFERENCE_ROOT", "/root/reference")
BIN = os.path.join(ROOT, "dropin", "_bin")


def parse_iterations(text):
    its = []
    for m in re.finditer(r"rewind enabled = (\d)\s+events = (\d+)\s+commits = (\d+)\s+deterministic = (\d)\s+checksum = (\d+)", text):
        its.append(dict(rewind=int(m.group(1)), commits=int(m.group(3)), deterministic=int(m.group(4)), checksum=int(m.group(5))))
    return its


def check_phold_test_output(text):
    assert "Looks good!" in text
    its = parse_iterations(text)
    assert len(its) == 4
    for it in its:   # SURVEY.md Appendix B / tests/golden/ref_goldens.json
        assert it["checksum"] == 15341107330209340805
        assert it["commits"] == (361624 if it["rewind"] else 180812)
        assert it["deterministic"] == 1


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("R", [2, 3, 8])
def test_reference_phold_test_source_through_cpp_mirror_emu(emu_api, tmp_path, R):
    exe = str(tmp_path / f"phold_test_R{R}")
    emu_dir = os.path.join(ROOT, "tests", "emu", "_build")
    subprocess.check_call([
        "g++", "-std=c++14", "-O2", "-pthread", "-w", f"-DDEVA_THREAD_N={R}", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
        "-DDEBUG=0", f"-I{ROOT}/include", "-include", f"{ROOT}/include/devastator_b200/models/phold_test_model.hxx",
        f"{REF}/test/phold.cxx", f"{ROOT}/devastator_b200/csrc/host_runtime.cpp", f"-L{emu_dir}", "-ldeva_b200_emu",
        f"-Wl,-rpath,{emu_dir}", "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    check_phold_test_output(out.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [2, 8])
def test_reference_phold_test_source_on_gpu(R):
    exe = os.path.join(BIN, f"phold_test_R{R}")
    assert os.path.exists(exe), "dropin/_bin missing: run __graft_entry__.build() where /root/reference exists"
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout
    check_phold_test_output(out.stdout)


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("gpus", [2, 3])
def test_reference_phold_test_source_multi_gpu_cpp_mirror_emu(emu_api, tmp_path, gpus):
    """DEVA_B200_GPUS = N: the C++ mirror partitions the LPs over N engines of one process and drains them
    concurrently, one host thread each (emulation: the concurrent drains meet in a lockstep drain)."""
    exe = str(tmp_path / "phold_test_R8")
    emu_dir = os.path.join(ROOT, "tests", "emu", "_build")
    subprocess.check_call([
        "g++", "-std=c++14", "-O2", "-pthread", "-w", "-DDEVA_THREAD_N=8", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
        "-DDEBUG=0", f"-I{ROOT}/include", "-include", f"{ROOT}/include/devastator_b200/models/phold_test_model.hxx",
        f"{REF}/test/phold.cxx", f"{ROOT}/devastator_b200/csrc/host_runtime.cpp", f"-L{emu_dir}", "-ldeva_b200_emu",
        f"-Wl,-rpath,{emu_dir}", "-o", exe])
    out = subprocess.run([exe], env=dict(os.environ, DEVA_B200_GPUS=str(gpus)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    check_phold_test_output(out.stdout)


@pytest.mark.gpu
def test_reference_phold_test_source_on_several_gpus_of_one_process():
    """The reference's unchanged test/phold.cxx on 2 GPUs (and on all of the box's GPUs) behind the C++ drop-in."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(BIN, "phold_test_R8")
    for gpus in sorted({2, min(n, 8)}):
        out = subprocess.run([exe], env=dict(os.environ, DEVA_B200_GPUS=str(gpus)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert out.returncode == 0, out.stdout
        check_phold_test_output(out.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 4])
def test_reference_phold_bench_source_on_gpu(R):
    """bench/phold.cxx, unchanged: wall-clock terminated, so only plumbing is checked (config 1 of
    SURVEY.md 8d): it runs, prints a report row, deterministic = 1."""
    exe = os.path.join(BIN, f"phold_bench_R{R}")
    assert os.path.exists(exe)
    env = dict(os.environ, lp_per_rank=str(1024 // R), ray_per_lp="1", peer_stddev="2", wall_secs="1", report_file="-")
    out = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout
    m = re.search(r"row\(xs=dict\((.*?)\),\s*ys=dict\((.*?)\)\)", out.stdout, re.S)
    assert m, out.stdout
    ys = dict(kv.split("=") for kv in m.group(2).replace(" ", "").split(","))
    assert float(ys["deterministic"]) == 1
    assert float(ys["commit_per_rank_per_sec"]) > 0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("R", [2, 3, 8])
def test_reference_gvt_test_source_through_gvt_mirror(tmp_path, R):
    """test/gvt-test.cxx, unchanged, against include/devastator/gvt.hxx (host-rank GVT service; no
    engine, no GPU).  Built with the test's own asserts on (`epoch_gvt() <= t` at every send and
    arrival): it ends only if the GVT reaches ~0 and every one of rank_n*100*t_end orbits landed."""
    exe = str(tmp_path / f"gvt_test_R{R}")
    subprocess.check_call([
        "g++", "-std=c++14", "-O2", "-pthread", "-w", f"-DDEVA_THREAD_N={R}", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
        "-DDEBUG=1", f"-I{ROOT}/include", f"{REF}/test/gvt-test.cxx", f"{ROOT}/devastator_b200/csrc/host_runtime.cpp",
        f"-L{ROOT}/devastator_b200", "-ldeva_b200", f"-Wl,-rpath,{ROOT}/devastator_b200", "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "SUCCESS", out.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("app,expect", [("barrier", "quitting"), ("bcast", "done"), ("send_ring", "SUCCESS"),
                                        ("send_vlen", "send/recv n = 3000 3000")])
def test_reference_world_tests_through_world_mirror(tmp_path, app, expect):
    """The reference's own tests of the SPMD world surface (run, barrier, send, bcast_procs, progress,
    reductions), unchanged, against include/devastator/world.hxx + host_runtime.cpp at 3 ranks.
    Their asserts are built in (DEBUG=1); send_vlen checks payload bytes, epochs and counts itself."""
    exe = str(tmp_path / app)
    subprocess.check_call([
        "g++", "-std=c++14", "-O2", "-pthread", "-w", "-DDEVA_THREAD_N=3", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
        "-DDEBUG=1", f"-I{ROOT}/include", f"{REF}/test/{app}.cxx", f"{ROOT}/devastator_b200/csrc/host_runtime.cpp",
        f"-L{ROOT}/devastator_b200", "-ldeva_b200", f"-Wl,-rpath,{ROOT}/devastator_b200", "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and expect in out.stdout, out.stdout[-2000:]


# ---- host-executed handlers (Tier G): event types without a device model ------------------------------------
def _build_host_app(tmp_path, src, R, flags=()):
    exe = str(tmp_path / (os.path.basename(src).replace(".cxx", "") + f"_R{R}"))
    emu_dir = os.path.join(ROOT, "tests", "emu", "_build")
    subprocess.check_call([
        "g++", "-std=c++14", "-O2", "-pthread", "-w", f"-DDEVA_THREAD_N={R}", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
        "-DDEBUG=0", "-DNDEBUG=1", *flags, f"-I{ROOT}/include", os.path.join(REF, src), f"{ROOT}/devastator_b200/csrc/host_runtime.cpp",
        f"-L{emu_dir}", "-ldeva_b200_emu", f"-Wl,-rpath,{emu_dir}", "-o", exe])
    return exe


PHOLD_BCAST_GOLDEN = {1: (99420, 840663732868261650), 2: (196980, 10848649704370676656)}


def check_phold_bcast_output(text, R):
    commits = [int(x) for x in re.findall(r"commits = (\d+)", text)]
    sums = [int(x) for x in re.findall(r"checksum = (\d+)", text)]
    assert len(commits) == 4 and len(sums) == 4, text
    assert set(commits) == {PHOLD_BCAST_GOLDEN[R][0]} and set(sums) == {PHOLD_BCAST_GOLDEN[R][1]}, text
    assert "deterministic = 0" not in text


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("R", [1, 2])
def test_reference_phold_bcast_source_host_handlers_emu(emu_api, tmp_path, R):
    """test/phold-bcast.cxx, unchanged: bcast_procs + default (sequence-id) subtimes through the host executor; the
    pending pool is the emulation's stand-in for the device pool.  Goldens: the reference's own binaries."""
    out = subprocess.run([_build_host_app(tmp_path, "test/phold-bcast.cxx", R)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    check_phold_bcast_output(out.stdout, R)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
def test_reference_stencil_source_host_handlers_emu(emu_api, tmp_path):
    """test/stencil.cxx, unchanged (64 M events; 80 same-time events per cell; commit hooks): checks itself against a
    serial recomputation and prints `Looks good!`"""
    out = subprocess.run([_build_host_app(tmp_path, "test/stencil.cxx", 2)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "Looks good!" in out.stdout, out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 2])
def test_reference_phold_bcast_source_on_gpu(R):
    exe = os.path.join(BIN, f"phold_bcast_R{R}")
    assert os.path.exists(exe), "dropin/_bin missing: run __graft_entry__.build() where /root/reference exists"
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout
    check_phold_bcast_output(out.stdout, R)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 4])
def test_reference_stencil_source_on_gpu(R):
    exe = os.path.join(BIN, f"stencil_R{R}")
    assert os.path.exists(exe)
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0 and "Looks good!" in out.stdout, out.stdout


# ---- example/heat (QSS actors), unchanged, through the host executor ------------------
with open(os.path.join(ROOT, "tests", "golden", "heat_goldens.json")) as _f:
    HEAT_GOLDEN = json.load(_f)     # the reference's own binaries, tests/golden/make_heat_goldens.py


def check_heat_output(exe, tag):
    g = HEAT_GOLDEN[tag]
    out = subprocess.run([exe], env=dict(os.environ, **g["env"]), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("Actors : [")]
    assert len(lines) == 2, out.stdout[-2000:]
    assert lines[1] == g["final"], (tag, lines[1], g["final"])      # the printed cell values, character for character


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("kind,R", [("heat", 1), ("heat", 4), ("heat_q0", 2)])
def test_reference_heat_source_host_handlers_emu(emu_api, tmp_path, kind, R):
    """example/heat/heat.cxx, unchanged: Advance / ShareState events (qss.hxx:92-144) with double-polynomial
    neighbour tables, both same-time priorities (time.hxx:7-12), executed in the pool's order; every golden
    configuration of its QSS order, ragged partitions included."""
    exe = _build_host_app(tmp_path, "example/heat/heat.cxx", R, ["-DQSS_ORDER=0"] if kind == "heat_q0" else [])
    for tag, g in HEAT_GOLDEN.items():
        if g["kind"] == kind:
            check_heat_output(exe, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,R", [("heat", 1), ("heat", 4), ("heat_q0", 2)])
def test_reference_heat_source_on_gpu(kind, R):
    exe = os.path.join(BIN, f"{kind}_R{R}")
    assert os.path.exists(exe), "dropin/_bin missing: run __graft_entry__.build() where /root/reference exists"
    for tag, g in HEAT_GOLDEN.items():
        if g["kind"] == kind:
            check_heat_output(exe, tag)


# ---- the host executor beyond "sends move time forward": rewindable drains, time stamps undone and redone ----------
DEBUG_ON = ["-UDEBUG", "-UNDEBUG", "-DDEBUG=1"]
with open(os.path.join(ROOT, "tests", "golden", "zero_delay_goldens.json")) as _f:
    ZERO_DELAY_GOLDEN = json.load(_f)   # tests/csrc/zero_delay.cxx on the UNMODIFIED reference runtime (make_zero_delay_goldens.py)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
def test_reference_phold_test_source_host_handlers_emu(emu_api, tmp_path):
    """test/phold.cxx, unchanged, WITHOUT a device model: its handlers run on the host, its rewindable drains in thirds
    (pdes.cxx:710-739,1137-1229) go through the pool's snapshot; same goldens as the device-resident path."""
    exe = _build_host_app(tmp_path, "test/phold.cxx", 2, DEBUG_ON)
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    check_phold_test_output(out.stdout)


def check_zero_delay(exe, only=None):
    for tag, g in ZERO_DELAY_GOLDEN.items():
        if only and tag not in only:
            continue
        out = subprocess.run([exe], env=dict(os.environ, DEVA_B200_HOSTQ_VERBOSE="1", **g["env"]), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-2000:]
        lines = out.stdout.splitlines()
        assert g["line"] in lines, (tag, out.stdout[-1000:])
        m = re.search(r"executed (\d+), committed (\d+), time stamps undone and redone key by key (\d+)", out.stdout)
        assert m, out.stdout
        if tag in ("default", "thirds", "all_zero"):     # these cannot be run without undoing: the path under test was taken
            assert int(m.group(3)) > 0 and int(m.group(1)) > int(m.group(2)), (tag, m.group(0))


@pytest.mark.parametrize("R", [1, 4])
def test_zero_delay_sends_host_handlers_emu(emu_api, tmp_path, R):
    """Sends that keep the time stamp land behind events their cd has run: the open time stamp is undone (unexecute,
    DEBUG checksums compared, children erased from the pool by handle) and redone key by key.  Executions, commit hooks
    and counts equal the reference runtime's on the same source, rewindable drains included."""
    check_zero_delay(_build_host_app(tmp_path, os.path.join(ROOT, "tests", "csrc", "zero_delay.cxx"), R, DEBUG_ON))


@pytest.mark.gpu
def test_reference_phold_test_source_host_handlers_on_gpu():
    exe = os.path.join(BIN, "phold_test_host_R2")
    assert os.path.exists(exe), "dropin/_bin missing: run __graft_entry__.build() where /root/reference exists"
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:]
    check_phold_test_output(out.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("R", [1, 4])
def test_zero_delay_sends_host_handlers_on_gpu(R):
    exe = os.path.join(BIN, f"zero_delay_R{R}")
    assert os.path.exists(exe), "dropin/_bin missing: run __graft_entry__.build()"
    # (every key of an undone time stamp is a round trip to the device: the cases are split over the two rank counts)
    check_zero_delay(exe, ("default", "all_zero_thirds", "one_ray") if R == 1 else ("thirds", "rare_zero", "all_zero"))


# ---- include/devastator/diagnostic.hxx: the assertion macros and deva::say as the applications use them ----------------
@pytest.mark.parametrize("debug", [0, 1])
def test_assert_macros_and_say(emu_api, tmp_path, debug):
    """One- and two-argument DEVA_ASSERT / DEVA_ASSERT_ALWAYS, streamed messages with manipulators, the one-argument form
    in an expression, DEVA_DEBUG_ONLY; a violated check prints the reference's block (diagnostic.cxx:34-63) and aborts."""
    flags = DEBUG_ON if debug else []
    exe = _build_host_app(tmp_path, os.path.join(ROOT, "tests", "csrc", "assert_probe.cxx"), 1, flags)
    run = lambda case: subprocess.run([exe, str(case)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    ok = run(0)
    assert ok.returncode == 0 and f"[0] evaluated={104 if debug else 2} y=7 hex=ff" in ok.stdout, ok.stdout
    one = run(1)
    assert one.returncode == -6, one.stdout                       # SIGABRT
    assert re.search(r"ASSERT FAILED rank=0 at=\S*assert_probe\.cxx:\d+\n\nFailed condition: x == 0\n/{50}", one.stdout), one.stdout
    two = run(2)
    assert two.returncode == -6, two.stdout
    assert "\n\nx is 41   7\nsecond line\n" in two.stdout and "Failed condition" not in two.stdout, two.stdout
