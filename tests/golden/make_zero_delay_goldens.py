#!/usr/bin/env python3
"""Generates tests/golden/zero_delay_goldens.json: tests/csrc/zero_delay.cxx (this repo's test model whose sends may keep
the time stamp) built against the UNMODIFIED reference runtime (oracle/_ref/ref_zero_delay_R*, oracle/build_ref.py) and
run on a few configurations.  The line it prints (commits, deterministic flag, order-sensitive hashes of the executions
and of the commit hooks) is the parity quantity; the reference gives the same line at 1, 2 and 4 ranks (checked here).

    python oracle/build_ref.py --only zero_delay --ranks 1 2 4 && python tests/golden/make_zero_delay_goldens.py
"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

CASES = {
    "default": dict(),
    "thirds": dict(zd_thirds=1),                                   # rewindable drains, each rewound once
    "all_zero": dict(zd_zero_of=1, zd_rays=50, zd_hops=200),       # every send keeps the time stamp: one time stamp per root time
    "all_zero_thirds": dict(zd_zero_of=1, zd_rays=50, zd_hops=200, zd_thirds=1),
    "rare_zero": dict(zd_zero_of=20, zd_rays=500, zd_hops=100),
    "one_ray": dict(zd_rays=1, zd_hops=1000, zd_zero_of=2),
}


def main():
    gold = {}
    for tag, env in CASES.items():
        e = dict(os.environ, pdes_heartbeat="0", **{k: str(v) for k, v in env.items()})
        res = {R: subprocess.check_output([os.path.join(REFDIR, f"ref_zero_delay_R{R}")], env=e, text=True).strip() for R in (1, 2, 4)}
        assert len(set(res.values())) == 1 and res[1].startswith("zero_delay commits"), (tag, res)
        gold[tag] = dict(env={k: str(v) for k, v in env.items()}, line=res[1])
        print(tag, res[1])
    with open(os.path.join(HERE, "zero_delay_goldens.json"), "w") as f:
        json.dump(gold, f, indent=1)


if __name__ == "__main__":
    main()
