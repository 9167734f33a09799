#!/usr/bin/env python3
"""Goldens at the BENCHMARKED configurations -> tests/golden/bench_goldens.json.

bench.py asserts its gathered commit count, XOR checksum, additive state hash and final GVT against
these at every GPU count, and tests/test_full_size_gpu.py does the same through the C ABI.

Sources (recorded per vector in "source"):
  * "reference": the reference's own binary oracle/_ref/ref_phold_t_cfg2_R8 (UNMODIFIED runtime built
    from /root/reference by oracle/build_ref.py) - BASELINE config 2 exactly as bench.py runs it
    (1 048 576 LPs x 16 roots, 10 % remote, end_time 1000).
  * "oracle": oracle/phold_oracle.c (the sequential restatement, itself pinned to the reference
    binaries on 30 vectors, tests/test_oracle.py) for the multi-GPU sizes, where the reference needs
    128 B x 16 events x 8 M LPs = 16 GB and ~half an hour of all host cores per vector.  With the
    reference's `subtime() = ray` tiebreak results do not depend on the rank decomposition
    (SURVEY Appendix A.8), so one vector per (LPs, remote fraction) serves every GPU count.
    The cfg2 vector is computed BOTH ways and must agree (asserted below).

Run in the build container:   python tests/golden/make_big_goldens.py [--jobs 7]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "bench_goldens.json")

K, LAMBDA, END = 16, 100.0, 1000
GRID = [(1 << 20, 0.10), (1 << 21, 0.10), (1 << 22, 0.10), (1 << 23, 0.10),
        (1 << 23, 0.50), (1 << 22, 0.50), (1 << 21, 0.50), (1 << 20, 0.50)]


def mix64(x):
    x = x.copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30); x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


def state_hash(st, lp_first=0):
    """== deva_b200_digest out[1] summed over all GPUs: sum_lp sum_w mix64(st[lp][w] ^ mix64(4*lp + w + 1)) mod 2^64."""
    gid = np.arange(lp_first, lp_first + st.shape[0], dtype=np.uint64)
    acc = np.uint64(0)
    with np.errstate(over="ignore"):
        for w in range(st.shape[1]):
            acc += np.add.reduce(mix64(st[:, w] ^ mix64(gid * np.uint64(4) + np.uint64(w + 1))))
    return int(acc)


def summarize(out, st, A, p, source, secs):
    return {"lp_n": A, "k": K, "p_remote": p, "lambda": LAMBDA, "end_time": END, "user_subtime": 1, "rootdiv": 1,
            "commits": int(out["commits"]), "checksum": int(out["checksum"]), "gvt": int(out["gvt"]),
            "deterministic": int(out["deterministic"]), "state_hash": state_hash(st),
            "state_sha256": hashlib.sha256(st.tobytes()).hexdigest(), "source": source, "secs": round(secs, 1)}


def run_oracle(args):
    A, p = args
    import oracle_lib
    t0 = time.time()
    out, st = oracle_lib.run(A, K, p, LAMBDA, END, ranks=1, user_subtime=1, rootdiv=1)
    return summarize(out, st, A, p, "oracle", time.time() - t0)


def run_reference_cfg2():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_phold_t_cfg2_R8")
    dump = f"/tmp/_big_golden_{os.getpid()}.bin"
    env = dict(os.environ, pdes_heartbeat="0", PHOLD_T_DUMP=dump, PHOLD_T_END=str(END))
    t0 = time.time()
    out = json.loads(subprocess.check_output([exe], env=env))
    st = np.fromfile(dump, dtype=np.uint64).reshape(-1, 3)
    os.unlink(dump)
    assert out["actor_n"] == 1 << 20 and out["end_time"] == END and out["user_subtime"] == 1
    return summarize(out, st, 1 << 20, 0.10, "reference", time.time() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=7)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    vecs = []
    if not a.skip_reference:
        r = run_reference_cfg2()
        print("reference cfg2:", r, flush=True)
        vecs.append(r)
    with ProcessPoolExecutor(max_workers=a.jobs) as ex:
        for r in ex.map(run_oracle, GRID):
            print("oracle:", r, flush=True)
            vecs.append(r)
            json.dump({"_about": "see make_big_goldens.py", "vectors": vecs}, open(OUT + ".partial", "w"), indent=1)
    if not a.skip_reference:
        ref, orc = vecs[0], vecs[1]
        for k in ("commits", "checksum", "gvt", "state_hash", "state_sha256", "deterministic"):
            assert ref[k] == orc[k], f"oracle and reference disagree on cfg2 {k}: {orc[k]} vs {ref[k]}"
    json.dump({"_about": "goldens at the benchmarked configurations; see make_big_goldens.py", "vectors": vecs},
              open(OUT, "w"), indent=1)
    os.path.exists(OUT + ".partial") and os.unlink(OUT + ".partial")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
