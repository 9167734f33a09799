#!/usr/bin/env python3
"""Generates tests/golden/ref_goldens.json (+ per-LP state dumps of the small cases) by running the
UNMODIFIED reference binaries in oracle/_ref/ (built by oracle/build_ref.py from /root/reference).

Run in the build container (needs /root/reference for the build step):
    python oracle/build_ref.py && python tests/golden/make_goldens.py
The outputs are committed; the GPU box only reads them.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_ref import PHOLD_T_CONFIGS  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
env0 = dict(os.environ, pdes_heartbeat="0")


def run_phold_t(tag, R, drain_end=None, end=None):
    dump = f"/tmp/_golden_dump_{os.getpid()}.bin"
    env = dict(env0, PHOLD_T_DUMP=dump)
    if drain_end is not None:
        env["PHOLD_T_DRAIN_END"] = str(drain_end)
    if end is not None:
        env["PHOLD_T_END"] = str(end)
    out = json.loads(subprocess.check_output([os.path.join(REFDIR, f"ref_phold_t_{tag}_R{R}")], env=env))
    st = np.fromfile(dump, dtype=np.uint64).reshape(-1, 3)
    os.unlink(dump)
    out.pop("drain_secs")
    out.pop("executed")      # varies run to run (rollback timing); not a parity quantity
    out["state_sha256"] = hashlib.sha256(st.tobytes()).hexdigest()
    return out, st


# PHOLD-B: bench/phold.cxx's model with an end time instead of the wall-clock cut-off (build_ref.py gen_phold_b)
PHOLD_B_CASES = [
    # (tag, ranks, lp_per_rank, ray_per_lp, peer_stddev, end_time, drain_end)
    ("b_small", 2, 64, 2, 2.0, 60000, None),
    ("b_pause", 2, 64, 2, 2.0, 60000, 25000),
    ("b_narrow", 8, 125, 2, 0.5, 40000, None),
    ("b_wide", 2, 500, 1, 8.0, 100000, None),
    ("b_big", 8, 4096, 2, 2.0, 30000, None),
]


def run_phold_b(R, lp_per_rank, ray_per_lp, stddev, end, drain_end):
    dump = f"/tmp/_golden_dump_b_{os.getpid()}.bin"
    env = dict(env0, PHOLD_B_DUMP=dump, PHOLD_B_END=str(end), lp_per_rank=str(lp_per_rank), ray_per_lp=str(ray_per_lp),
               peer_stddev=repr(stddev))
    if drain_end is not None:
        env["PHOLD_B_DRAIN_END"] = str(drain_end)
    out = json.loads(subprocess.check_output([os.path.join(REFDIR, f"ref_phold_b_R{R}")], env=env))
    st = np.fromfile(dump, dtype=np.uint64).reshape(-1, 2)
    os.unlink(dump)
    out["state_sha256"] = hashlib.sha256(st.tobytes()).hexdigest()
    return out, st


def parse_phold_test(text):
    iters = []
    cur = None
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("rewind enabled"):
            cur = {"rewind": int(line.split("=")[1])}
            iters.append(cur)
        elif "=" in line and cur is not None:
            k, v = [s.strip() for s in line.split("=")]
            if k != "events":
                cur[k] = int(v)
    return iters


def main():
    gold = {"_about": "printed by the reference's own binaries (oracle/_ref); see make_goldens.py",
            "phold_t": [], "phold_test": [], "phold_b": []}
    for tag, R, per, k, sd, end, de in PHOLD_B_CASES:
        out, st = run_phold_b(R, per, k, sd, end, de)
        out["tag"] = tag
        gold["phold_b"].append(out)
        if per * R <= 1000 and de is None:
            np.save(os.path.join(HERE, f"phold_b_{tag}_state.npy"), st)
    for tag, cfg in PHOLD_T_CONFIGS.items():
        if tag in ("cfg2", "cfg3"):   # full size: see make_big_goldens.py
            continue
        for R in (2, 8):
            out, st = run_phold_t(tag, R)
            out["tag"] = tag
            out["cfg"] = cfg
            gold["phold_t"].append(out)
            if cfg["A"] <= 1000 and R == 2:
                np.save(os.path.join(HERE, f"phold_t_{tag}_state.npy"), st)
    # drain(t_end) pause points: GVT returned and state at the pause
    for de in (50, 150):
        out, st = run_phold_t("small", 2, drain_end=de)
        out["tag"] = "small"
        out["cfg"] = PHOLD_T_CONFIGS["small"]
        gold["phold_t"].append(out)
    # BASELINE config 2 geometry (1M LPs x 16), short horizon so that the reference finishes in seconds
    out, st = run_phold_t("cfg2", 8, end=20)
    out["tag"] = "cfg2"
    out["cfg"] = dict(PHOLD_T_CONFIGS["cfg2"], END=20)
    gold["phold_t"].append(out)
    # the reference's own test/phold.cxx, untouched
    for R in (2, 8):
        txt = subprocess.check_output([os.path.join(REFDIR, f"ref_phold_test_R{R}")], env=env0, text=True)
        assert "Looks good!" in txt
        gold["phold_test"].append({"ranks": R, "iterations": parse_phold_test(txt)})
    with open(os.path.join(HERE, "ref_goldens.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "ref_goldens.json"))


if __name__ == "__main__":
    main()
