#!/usr/bin/env python3
"""Generates tests/golden/heat_goldens.json by running the UNMODIFIED reference example/heat/heat.cxx
(oracle/_ref/ref_heat_R*, ref_heat_q0_R* - built by oracle/build_ref.py from /root/reference) on a few
configurations of its own environment knobs (example/heat/heat.cxx:23-31, qss.hxx:17, run.sh).

The app prints its cells' values before and after the drain ("Actors : [...]", actor_space.hxx:79-98); the
final line is the parity quantity.  The reference's result is checked here to be the same at 1, 2 and 4
ranks (rank threads), so one vector per configuration serves every rank count.

    python oracle/build_ref.py --only heat --ranks 1 2 4 && python tests/golden/make_heat_goldens.py
"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# (tag, binary kind, environment)
CASES = [
    ("run_sh", "heat", dict(cells=101, alpha=0.01, value_thresh=0.1, local_delta_t=1, sim_end_time=1000, sim_conf="source")),
    ("defaults", "heat", dict()),
    ("diri", "heat", dict(cells=101, sim_end_time=1000, sim_conf="diri")),
    ("fine", "heat", dict(cells=400, sim_end_time=3000, alpha=0.05, value_thresh=0.02)),
    ("coarse_steps", "heat", dict(cells=64, sim_end_time=500, local_delta_t=3, max_advance_delta_t=50)),
    ("order0", "heat_q0", dict(cells=101, sim_end_time=1000)),
    ("order0_diri", "heat_q0", dict(cells=50, sim_end_time=2000, sim_conf="diri", value_thresh=0.05)),
    ("ragged", "heat", dict(cells=7, sim_end_time=300, sim_conf="diri")),     # 7 cells over 4 ranks: 2,2,2,1
]


def final_line(exe, env):
    out = subprocess.check_output([exe], env=dict(os.environ, pdes_heartbeat="0", **{k: str(v) for k, v in env.items()}), text=True)
    lines = [l for l in out.splitlines() if l.startswith("Actors : [")]
    assert len(lines) == 2, out
    return lines[1]


def main():
    gold = {}
    for tag, kind, env in CASES:
        res = {R: final_line(os.path.join(REFDIR, f"ref_{kind}_R{R}"), env) for R in (1, 2, 4)}
        assert len(set(res.values())) == 1, (tag, res)
        gold[tag] = dict(kind=kind, env={k: str(v) for k, v in env.items()}, final=res[1])
        print(tag, res[1][:100])
    with open(os.path.join(HERE, "heat_goldens.json"), "w") as f:
        json.dump(gold, f, indent=1)


if __name__ == "__main__":
    main()
