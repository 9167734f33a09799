#!/usr/bin/env python3
"""Generates tests/golden/heat2d_goldens.json: tests/csrc/heat2d.cxx (this repo's 2-D heat stencil with nearest-neighbour
events, self-checking against a serial recomputation) built against the UNMODIFIED reference runtime
(oracle/_ref/ref_heat2d_R*, oracle/build_ref.py).  The reference prints the same two lines at 1, 2 and 4 ranks (checked).

    python oracle/build_ref.py --only heat2d --ranks 1 2 4 && python tests/golden/make_heat2d_goldens.py
"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

CASES = {
    "default": dict(),                                   # 48 x 48, 40 steps
    "thirds": dict(h2_thirds=1),                         # rewindable drains, each rewound once
    "ragged": dict(h2_n=7, h2_steps=9),                  # 49 cells over 4 ranks: 13, 13, 13, 10
    "ragged_thirds": dict(h2_n=7, h2_steps=9, h2_thirds=1),
    "two": dict(h2_n=2, h2_steps=5),
    "one": dict(h2_n=1, h2_steps=5),                     # no neighbours: no events at all
    "wide": dict(h2_n=128, h2_steps=10, h2_alpha=0.05),
}


def main():
    gold = {}
    for tag, env in CASES.items():
        e = dict(os.environ, pdes_heartbeat="0", **{k: str(v) for k, v in env.items()})
        res = {R: subprocess.check_output([os.path.join(REFDIR, f"ref_heat2d_R{R}")], env=e, text=True).strip().splitlines() for R in (1, 2, 4)}
        assert all(r == res[1] for r in res.values()) and res[1][-1] == "Looks good!", (tag, res)
        gold[tag] = dict(env={k: str(v) for k, v in env.items()}, lines=res[1])
        print(tag, res[1][0])
    with open(os.path.join(HERE, "heat2d_goldens.json"), "w") as f:
        json.dump(gold, f, indent=1)


if __name__ == "__main__":
    main()
