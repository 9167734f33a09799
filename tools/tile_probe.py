#!/usr/bin/env python3
"""GPU probe: the TILE engine vs the first (thread-per-LP) engine on BASELINE config 2 geometry.
Checks commits / XOR checksum / additive state hash against tests/golden/bench_goldens.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from devastator_b200 import engine as E
if os.environ.get("QB_LIB"):   # A/B against another build of the same ABI
    E.LIB_PATH = os.path.join(ROOT, os.environ["QB_LIB"])

A = int(os.environ.get("QB_A", 1 << 20)); K = 16; P = float(os.environ.get("QB_P", 0.1)); LAM = 100.0
END = int(os.environ.get("QB_END", 1000)); REPS = int(os.environ.get("QB_REPS", 3))
gold = {(v["lp_n"], v["p_remote"]): v for v in json.load(open(os.path.join(ROOT, "tests/golden/bench_goldens.json")))["vectors"]}
g = gold.get((A, P)) if END == 1000 else None

def small_check(kind):
    import numpy as np, oracle_lib as O
    os.environ["DEVA_B200_ENGINE"] = kind
    ok = True
    for (a, k, p, lam, end, w) in [(64, 4, 0.5, 10.0, 400, 0), (1000, 16, 0.5, 100.0, 300, 0), (5000, 16, 0.5, 100.0, 300, 64), (4096, 16, 0.1, 100.0, 400, 0)]:
        o, ost = O.run(a, k, p, lam, end)
        eng = E.Engine(E.MODEL_PHOLD_TEST, a, window=w, p_remote=p, lam=lam, end_time=end, user_subtime=1)
        eng.seed_phold(k, 1, 1); gvt = eng.drain(); st = eng.get_state(); s = eng.stats()
        good = s["committed_n"] == o["commits"] and np.array_equal(st, ost) and gvt == o["gvt"]
        ok &= good
        print(kind, "small", a, p, w, "OK" if good else "FAIL", s["committed_n"], o["commits"], flush=True)
        eng.close()
    return ok

for kind in os.environ.get("QB_ENGINES", "tile,lp").split(","):
    os.environ["DEVA_B200_ENGINE"] = kind
    if not os.environ.get("QB_SKIP_SMALL"): small_check(kind)
    eng = E.Engine(E.MODEL_PHOLD_TEST, A, p_remote=P, lam=LAM, end_time=END, window=int(os.environ.get("QB_WINDOW", 0)))
    for rep in range(REPS):
        eng.seed_phold(K, 1, 1)
        s0 = eng.stats(); t0 = time.time(); gvt = eng.drain(); wall = time.time() - t0; s1 = eng.stats()
        dms = s1["drain_ms"] - s0["drain_ms"]; kms = s1["exec_kernel_ms"] - s0["exec_kernel_ms"]
        c = s1["committed_n"] - s0["committed_n"]; x = s1["executed_n"] - s0["executed_n"]; p = s1["passes"] - s0["passes"]
        d = eng.digest()
        par = None
        if g: par = "MATCH" if (c == g["commits"] and d["checksum"] == g["checksum"] and d["state_hash"] == g["state_hash"]) else "MISMATCH"
        print(json.dumps(dict(engine=kind, rep=rep, commits=c, executed=x, eff=round(c / max(x, 1), 4), launches=p, drain_ms=round(dms, 2), kernel_ms=round(kms, 2),
                              wall_ms=round(wall * 1e3, 2), Gev_s=round(c / dms / 1e6, 3), Wlast=s1["window_last"], checksum=d["checksum"], golden=par,
                              rolled=s1["rolled_back_n"] - s0["rolled_back_n"], frac_120B=round(c * 120 / (dms * 1e-3) / 6545.6e9, 4))), flush=True)
    eng.close()
