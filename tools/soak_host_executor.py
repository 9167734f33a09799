#!/usr/bin/env python3
"""Randomised soak of the host executor (host-executed handlers over the pending pool) against the UNMODIFIED reference
runtime, in the build container (needs /root/reference and oracle/_ref/ref_{zero_delay,heat}_R{1,2,4}):

    python oracle/build_ref.py --only zero_delay,heat --ranks 1 2 4
    python tools/soak_host_executor.py <seed> <runs>

Every run draws a rank count and the environment knobs of tests/csrc/zero_delay.cxx (sends that keep the time stamp,
rewound drains) or of the reference's example/heat/heat.cxx, runs the reference binary and the same source built against
include/devastator/*.hxx + host_runtime.cpp + the emulation's stand-in for the device pool, and compares what they print.
"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEVA_REFERENCE_ROOT", "/root/reference")
EMU = os.path.join(ROOT, "tests", "emu", "_build")


def build(tmp, name, src, R, flags):
    exe = os.path.join(tmp, f"{name}_R{R}")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-pthread", "-w", f"-DDEVA_THREAD_N={R}", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1",
                           *flags, f"-I{ROOT}/include", src, f"{ROOT}/devastator_b200/csrc/host_runtime.cpp", f"-L{EMU}", "-ldeva_b200_emu",
                           f"-Wl,-rpath,{EMU}", "-o", exe])
    return exe


def last_line(cmd, env):
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    lines = r.stdout.strip().splitlines()
    return (lines[-1] if lines else ""), r.returncode, r.stderr[-400:]


def main():
    seed, runs = int(sys.argv[1]), int(sys.argv[2])
    random.seed(seed)
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        exes = {}
        for R in (1, 2, 4):
            exes["zero_delay", R] = build(tmp, "zero_delay", os.path.join(ROOT, "tests", "csrc", "zero_delay.cxx"), R, ["-DDEBUG=1"])
            exes["heat", R] = build(tmp, "heat", os.path.join(REF, "example", "heat", "heat.cxx"), R, ["-DDEBUG=0", "-DNDEBUG=1"])
        for _ in range(runs):
            R = random.choice([1, 2, 4])
            if random.random() < 0.5:
                app = "zero_delay"
                knobs = dict(zd_rays=random.choice([1, 3, 17, 96, 300, 1000]), zd_hops=random.choice([0, 1, 5, 50, 300]),
                             zd_zero_of=random.choice([1, 2, 3, 7, 50]), zd_thirds=random.choice([0, 1]))
            else:
                app = "heat"
                knobs = dict(cells=random.choice([1, 2, 5, 33, 100, 257]), alpha=random.choice([0.001, 0.01, 0.05, 0.2]),
                             value_thresh=random.choice([0.01, 0.1, 0.5]), local_delta_t=random.choice([1, 2, 5]),
                             sim_end_time=random.choice([1, 50, 400, 1500]), sim_conf=random.choice(["source", "diri"]),
                             max_advance_delta_t=random.choice([10, 100, 100000]))
            env = dict(os.environ, pdes_heartbeat="0", **{k: str(v) for k, v in knobs.items()})
            want, _, _ = last_line([os.path.join(ROOT, "oracle", "_ref", f"ref_{app}_R{R}")], env)
            got, rc, err = last_line([exes[app, R]], env)
            if want != got or rc != 0:
                bad += 1
                print("MISMATCH", app, R, knobs, want[:200], got[:200], err)
    print(f"seed {seed}: {runs} runs, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
