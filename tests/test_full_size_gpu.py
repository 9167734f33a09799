"""BASELINE.json config-2 geometry (1 048 576 LPs x 16 roots, 10 % remote) on the GPU.

At this size the sequential oracle takes minutes, so parity is checked (a) against the golden the
REFERENCE BINARY itself produced at this size (tests/golden/ref_goldens.json, tag cfg2: commits,
checksum, SHA-256 of every per-LP state word) and (b) through size-independent properties of the
domain: the optimism window never changes committed results, drain(t_end) segments add up to one
drain, and drain -> rewind(true) -> drain reproduces the plain run (test/phold.cxx:180-198)."""
import hashlib

import numpy as np
import pytest

import scenarios as S
from devastator_b200 import engine as E

A, K, P, LAM = 1 << 20, 16, 0.10, 100.0
pytestmark = pytest.mark.gpu


def fresh(cuda_api, end, window):
    eng = E.Engine(E.MODEL_PHOLD_TEST, A, p_remote=P, lam=LAM, end_time=end, window=window, api=cuda_api)
    eng.seed_phold(K, 1, 1)
    return eng


def summary(eng, gvt):
    s, d = eng.stats(), eng.digest()
    return dict(commits=s["committed_n"], checksum=d["checksum"], state_hash=d["state_hash"], pending=d["pending"],
                pending_hash=d["pending_hash"], gvt=gvt, deterministic=s["deterministic"])


def test_reference_binary_golden_1m_lps(cuda_api):
    g = [x for x in S.goldens()["phold_t"] if x["tag"] == "cfg2"][0]
    eng = fresh(cuda_api, g["end_time"], 0)
    gvt = eng.drain()
    st = eng.get_state()
    s, d = eng.stats(), eng.digest()
    assert s["committed_n"] == g["commits"] == 18697550
    assert d["checksum"] == g["checksum"]
    assert gvt == g["gvt"] and s["deterministic"] == g["deterministic"]
    assert hashlib.sha256(st.tobytes()).hexdigest() == g["state_sha256"]


def test_window_width_never_changes_results_full_size(cuda_api):
    ref = None
    for window in (8, 32, 200, 0):
        eng = fresh(cuda_api, 150, window)
        out = summary(eng, eng.drain())
        eng.close()
        if ref is None:
            ref = out
        assert out == ref, (window, out, ref)
    assert ref["commits"] > 3e7 and ref["pending"] == 0


def test_segments_and_rewind_round_trip_full_size(cuda_api):
    plain = fresh(cuda_api, 150, 0)
    want = summary(plain, plain.drain())
    plain.close()
    # drain(t_end) slices add up to the single drain
    eng = fresh(cuda_api, 150, 0)
    for t_end in (40, 90):
        assert eng.drain(t_end) >= t_end
    got = summary(eng, eng.drain())
    assert got == want
    eng.close()
    # pause / rewind / resume: statistics are not rewound, everything else is
    eng = fresh(cuda_api, 150, 0)
    eng.drain(60, rewindable=True)
    c1 = eng.stats()["committed_n"]
    eng.rewind(True)
    np.testing.assert_array_equal(eng.get_state(0, 4096), E.phold_initial_state_np(E.MODEL_PHOLD_TEST, 0, 4096))
    got = summary(eng, eng.drain())
    got["commits"] -= c1
    assert got == want
