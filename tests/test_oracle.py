"""The oracle restatement (oracle/phold_oracle.c) against the golden vectors printed by the
reference's own binaries (tests/golden/ref_goldens.json, made by tests/golden/make_goldens.py)."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O
import scenarios as S


def _cases():
    return [g for g in S.goldens()["phold_t"]]


@pytest.mark.parametrize("g", _cases(), ids=lambda g: f"{g['tag']}-R{g['ranks']}-de{g['drain_end'] % 1000}-end{g['end_time']}")
def test_oracle_matches_reference_binary(g):
    c = g["cfg"]
    if c["A"] > 200000 and os.environ.get("DEVA_TEST_FULL") != "1":
        pytest.skip("1M-LP oracle run (about a minute of CPU); set DEVA_TEST_FULL=1")
    o, st = O.run(c["A"], c["K"], c["P"], c["LAMBDA"], g["end_time"], ranks=g["ranks"], user_subtime=c["sub"],
                  rootdiv=c["rootdiv"], drain_end=g["drain_end"])
    assert o["commits"] == g["commits"]
    assert o["checksum"] == g["checksum"]
    assert o["gvt"] == g["gvt"]
    assert o["deterministic"] == g["deterministic"]
    assert hashlib.sha256(st.tobytes()).hexdigest() == g["state_sha256"]


@pytest.mark.parametrize("tag", ["tiny", "small", "smallq", "seqid", "allrem", "norem", "dense", "seq3"])
def test_oracle_per_lp_state_fixture(tag):
    g = [x for x in S.goldens()["phold_t"] if x["tag"] == tag and x["ranks"] == 2 and x["drain_end"] == 2**64 - 1][0]
    c = g["cfg"]
    want = np.load(os.path.join(S.GOLDEN, f"phold_t_{tag}_state.npy"))
    o, st = O.run(c["A"], c["K"], c["P"], c["LAMBDA"], c["END"], ranks=2, user_subtime=c["sub"], rootdiv=c["rootdiv"])
    np.testing.assert_array_equal(st, want)


def test_reference_phold_test_golden():
    """test/phold.cxx itself: 180 812 commits per plain drain, checksum 15341107330209340805
     - the restatement of its model reproduces both."""
    for run in S.goldens()["phold_test"]:
        for it in run["iterations"]:
            assert it["checksum"] == 15341107330209340805
            assert it["commits"] == (361624 if it["rewind"] else 180812)
    o, _ = O.run(1000, 2, 0.5, 100.0, 10000, ranks=2, user_subtime=0, rootdiv=0)
    assert (o["commits"], o["checksum"], o["deterministic"]) == (180812, 15341107330209340805, 1)


@pytest.mark.parametrize("g", S.goldens().get("phold_b", []), ids=lambda g: g["tag"])
def test_oracle_bench_model_matches_reference_binary(g):
    """bench/phold.cxx's model (xorshift rng, exponential delay, Box-Muller destination, sequence-id
    tiebreak) against PHOLD-B: the reference's own handler code compiled with an end time in place of
    the wall-clock cut-off (oracle/build_ref.py gen_phold_b)."""
    A = g["lp_per_rank"] * g["ranks"]
    o, st = O.run(A, g["ray_per_lp"], 0.0, 10000.0, g["end_time"], model=O.MODEL_BENCH, ranks=g["ranks"], user_subtime=0,
                  rootdiv=0, bench_lp_per_rank=g["lp_per_rank"], bench_peer_stddev=g["peer_stddev"], drain_end=g["drain_end"])
    assert (o["commits"], o["gvt"], o["deterministic"]) == (g["commits"], g["gvt"], g["deterministic"])
    assert hashlib.sha256(np.ascontiguousarray(st[:, :2]).tobytes()).hexdigest() == g["state_sha256"]
