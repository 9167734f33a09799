#!/usr/bin/env python3
"""Throw-away exploration: PHOLD-T at BASELINE config-2 geometry, sweep the window width."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from devastator_b200 import engine as E
if os.environ.get("QB_LIB"):   # A/B against another build of the same ABI
    E.LIB_PATH = os.path.join(ROOT, os.environ["QB_LIB"])

A = int(os.environ.get("QB_A", 1 << 20)); K = 16; P = float(os.environ.get("QB_P", 0.1)); LAM = 100.0
END = int(os.environ.get("QB_END", 300))
windows = [int(x) for x in os.environ.get("QB_WINDOWS", "4,16,32,64,128,256,0").split(",")]
gold = json.load(open(os.path.join(ROOT, "tests/golden/ref_goldens.json")))
g = [x for x in gold["phold_t"] if x["tag"] == "cfg2"][0]
if A == 1 << 20 and P == 0.1:
    eng = E.Engine(E.MODEL_PHOLD_TEST, A, p_remote=P, lam=LAM, end_time=20, window=16)
    eng.seed_phold(K, 1, 1); gvt = eng.drain(); st = eng.get_state(); s = eng.stats(); d = eng.digest()
    ok = s["committed_n"] == g["commits"] and d["checksum"] == g["checksum"] and hashlib.sha256(st.tobytes()).hexdigest() == g["state_sha256"]
    print("cfg2 END=20 golden (reference binary, 1M LPs):", "MATCH" if ok else "MISMATCH", s["committed_n"], g["commits"], flush=True)
    eng.close()
for W in windows:
    eng = E.Engine(E.MODEL_PHOLD_TEST, A, p_remote=P, lam=LAM, end_time=END, window=W)
    best = None
    for rep in range(2):
        eng.seed_phold(K, 1, 1)
        s0 = eng.stats(); t0 = time.time(); gvt = eng.drain(); wall = time.time() - t0; s1 = eng.stats()
        dms = s1["drain_ms"] - s0["drain_ms"]; kms = s1["exec_kernel_ms"] - s0["exec_kernel_ms"]
        c = s1["committed_n"] - s0["committed_n"]; x = s1["executed_n"] - s0["executed_n"]; p = s1["passes"] - s0["passes"]
        best = dict(W=W, commits=c, executed=x, eff=round(c / x, 3), passes=p, drain_ms=round(dms, 2), kernel_ms=round(kms, 2),
                    wall_ms=round(wall * 1e3, 2), Mev_s=round(c / dms / 1e3, 1), Mev_s_kernel=round(c / kms / 1e3, 1), Wlast=s1["window_last"], checksum=eng.digest()["checksum"])
    print(json.dumps(best), flush=True)
    eng.close()
