"""Parity scenarios shared by the CPU (emulation harness) and GPU (libdeva_b200.so) suites.

Every scenario compares an engine reached through the C ABI with the oracle (tests/oracle_lib.py,
pinned against the reference's own binaries) on the same seeded input: committed-event count,
per-LP state words (word for word), checksum, final GVT, deterministic flag."""
import json
import os

import numpy as np

import oracle_lib as O
from devastator_b200 import engine as E

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def goldens():
    return json.load(open(os.path.join(GOLDEN, "ref_goldens.json")))


def make(api, A, *, K=None, P=0.1, LAM=100.0, END=300, sub=1, window=0, model=E.MODEL_PHOLD_TEST, **kw):
    return E.Engine(model, A, window=window, p_remote=P, lam=LAM, end_time=END, user_subtime=sub, api=api, **kw)


def check_against_oracle(eng, o, ost, gvt, *, rewound_factor=1):
    st = eng.get_state()
    s = eng.stats()
    d = eng.digest()
    assert s["committed_n"] == rewound_factor * o["commits"], (s, o)
    np.testing.assert_array_equal(st[:, :ost.shape[1]] if st.shape[1] > ost.shape[1] else st, ost[:, :st.shape[1]])
    assert gvt == o["gvt"]
    if st.shape[1] == 3:
        assert d["checksum"] == o["checksum"]
    assert s["deterministic"] == o["deterministic"]
    assert d["pending"] == o["pending"]
    return s


def run_phold_t(api, A, K, P, LAM, END, *, window=0, sub=1, rootdiv=1, ranks=1, drain_end=E.T_INF, **kw):
    o, ost = O.run(A, K, P, LAM, END, ranks=ranks, user_subtime=sub, rootdiv=rootdiv, drain_end=drain_end)
    eng = make(api, A if sub else _lp_space(A, ranks), P=P, LAM=LAM, END=END, sub=sub, window=window,
               lp_n_model=A, **kw)
    eng.seed_phold(K, rootdiv, ranks)
    gvt = eng.drain(drain_end)
    if not sub:   # default sequence ids: LP space is ranks*ceil(A/ranks) (test/phold.cxx:46, pdes.cxx:319)
        st = eng.get_state()[:A]
        s = eng.stats()
        assert s["committed_n"] == o["commits"]
        np.testing.assert_array_equal(st, ost)
        assert gvt == o["gvt"] and s["deterministic"] == o["deterministic"]
    else:
        s = check_against_oracle(eng, o, ost, gvt)
    eng.close()
    return s, o


def _lp_space(A, ranks):
    per = (A + ranks - 1) // ranks
    return per * ranks
