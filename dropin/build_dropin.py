#!/usr/bin/env python3
"""Builds the reference's OWN app sources, unchanged, against this repo's header mirror
(include/devastator/*.hxx) and the B200 engine - the drop-in check of the C++ boundary.

    /root/reference/test/phold.cxx   -> dropin/_bin/phold_test_R<R>     (self-checking: "Looks good!")
    /root/reference/bench/phold.cxx  -> dropin/_bin/phold_bench_R<R>    (prints a report row)
    /root/reference/test/gvt-test.cxx -> dropin/_bin/gvt_test_R<R>      (host-rank GVT service, asserts on: "SUCCESS")
    /root/reference/test/stencil.cxx, test/phold-bcast.cxx, example/heat/heat.cxx -> stencil_R<R>, phold_bcast_R<R>,
                                        heat_R<R>, heat_q0_R<R>         (host-executed handlers over the device pool)

The sources are compiled where they lie under /root/reference (never copied).  The only additions
on the build line are -I include and the force-included model binding header.  Outputs are
git-ignored binaries that travel to the GPU box with the snapshot (like the built .so files);
when /root/reference is absent (GPU box) the prebuilt binaries are used as they are.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DEVA_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_bin")
INC = os.path.join(ROOT, "include")
RUNTIME = os.path.join(ROOT, "devastator_b200", "csrc", "host_runtime.cpp")
LIBDIR = os.path.join(ROOT, "devastator_b200")

APPS = [
    # (name, source under the reference, model header, extra reference sources, rank counts)
    ("phold_test", "test/phold.cxx", "devastator_b200/models/phold_test_model.hxx", [], [2, 8]),
    ("phold_bench", "bench/phold.cxx", "devastator_b200/models/phold_bench_model.hxx", ["bench/util/report.cxx"], [1, 4, 8]),
    # no device model: drives deva::gvt (include/devastator/gvt.hxx) directly; built with the asserts of the
    # test enabled (DEBUG=1), they are what checks that the GVT never overtakes a message
    ("gvt_test", "test/gvt-test.cxx", None, [], [2, 3, 8]),
    # no device model either: HOST-EXECUTED handlers (Tier G) - E::execute, commit hooks and bcast_procs run on the
    # rank threads, the pending set / GVT / order of execution come from the GPU pool (deva_b200_hostq_*)
    ("stencil", "test/stencil.cxx", "", [], [1, 2, 4]),
    ("phold_bcast", "test/phold-bcast.cxx", "", [], [1, 2]),
    # example/heat (1-D heat diffusion, quantized-state-system actors), unchanged, Tier G:
    # Advance / ShareState handlers (qss.hxx:92-144) run on the rank threads in the order the device pool decides
    ("heat", "example/heat/heat.cxx", "", [], [1, 4]),
    ("heat_q0", "example/heat/heat.cxx", "", [], [2], ["-DQSS_ORDER=0"]),
    # test/phold.cxx WITHOUT its device model: the same source through the host executor (rewindable drains in thirds,
    # DEBUG checksums), against the same goldens
    ("phold_test_host", "test/phold.cxx", "", [], [2], ["-UDEBUG", "-UNDEBUG", "-DDEBUG=1"]),
    # this repo's own test model whose sends may keep the time stamp (tests/csrc/zero_delay.cxx): time stamps undone and
    # redone key by key, erase by handle, rewindable drains; goldens from the same source on the reference runtime
    ("zero_delay", os.path.join(ROOT, "tests", "csrc", "zero_delay.cxx"), "", [], [1, 4], ["-UDEBUG", "-UNDEBUG", "-DDEBUG=1"]),
    # this repo's own 2-D heat stencil with nearest-neighbour events (BASELINE config 5's wording; tests/csrc/heat2d.cxx):
    # checks itself against a serial recomputation; goldens from the same source on the reference runtime
    ("heat2d", os.path.join(ROOT, "tests", "csrc", "heat2d.cxx"), "", [], [1, 4]),
]


def newer(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def main():
    force = "--force" in sys.argv
    if not os.path.isdir(REF):
        print(f"dropin: {REF} not present - keeping prebuilt dropin/_bin as is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    hdrs = [os.path.join(dp, f) for dp, _, fs in os.walk(INC) for f in fs]
    for name, src, model, extra, ranks, *more in APPS:
        flags = more[0] if more else []
        for R in ranks:
            exe = os.path.join(OUT, f"{name}_R{R}")
            srcs = [os.path.join(REF, src)] + [os.path.join(REF, e) for e in extra]
            if not (force or newer(exe, srcs + hdrs + [RUNTIME, os.path.join(LIBDIR, "libdeva_b200.so"), __file__])):
                continue
            dbg = ["-DDEBUG=1"] if model is None else ["-DDEBUG=0", "-DNDEBUG=1"]
            inc_model = [] if not model else ["-include", os.path.join(INC, model)]
            cmd = ["g++", "-std=c++14", "-O2", "-pthread", "-w", f"-DDEVA_THREAD_N={R}", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
                   *dbg, *flags, f"-I{INC}", f"-I{os.path.join(REF, 'bench')}", *inc_model,
                   *srcs, RUNTIME, f"-L{LIBDIR}", "-ldeva_b200", "-Wl,-rpath,$ORIGIN/../../devastator_b200", "-o", exe]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
                raise SystemExit("dropin build failed")
            print("built", os.path.relpath(exe, ROOT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
