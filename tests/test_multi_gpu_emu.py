"""Block-partitioned multi-GPU protocol (per-window exchange of cross-GPU records + min GVT),
checked on the CPU emulation harness: G engines in lockstep, emu_drain_multi standing in for the
NCCL grouped send/recv and min-allreduce of engine.cu.  Results must equal the 1-GPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import scenarios as S
from devastator_b200 import engine as E


def drain_multi(api, engs, t_end=E.T_INF, rewindable=False):
    arr = (C.c_void_p * len(engs))(*[e.h for e in engs])
    gvt = C.c_uint64()
    api.check(api.lib.emu_drain_multi(arr, len(engs), t_end, int(rewindable), C.byref(gvt)))
    return gvt.value


def make_group(api, G, A, peer=False, **kw):
    per = (A + G - 1) // G
    engs = [E.Engine(E.MODEL_PHOLD_TEST, A, lp_begin=g * per, lp_count=min(per, A - g * per), n_gpus=G, gpu_rank=g,
                     api=api, **kw) for g in range(G)]
    if peer:   # records stored straight into the destination engine's receive buffer (deva_b200_comm_init's IPC mapping)
        arr = (C.c_void_p * G)(*[e.h for e in engs])
        api.lib.emu_enable_peer_delivery(arr, G)
    return engs


@pytest.mark.parametrize("peer", [False, True], ids=["staged", "peer"])
@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("window", [1, 20, 0])
@pytest.mark.parametrize("P", [0.1, 0.5])
def test_partitioned_equals_oracle(emu_api, engine_kind, G, window, P, peer):
    if engine_kind == "tile" and not peer:
        pytest.skip("the tile engine has one transport: peer delivery + device-side exchange")
    A, K, LAM, END = 1000, 16, 100.0, 300
    o, ost = O.run(A, K, P, LAM, END)
    engs = make_group(emu_api, G, A, peer=peer, window=window, p_remote=P, lam=LAM, end_time=END)
    for e in engs:
        e.seed_phold(K, 1, 1)
    gvt = drain_multi(emu_api, engs)
    st = np.concatenate([e.get_state() for e in engs])
    np.testing.assert_array_equal(st, ost)
    stats = [e.stats() for e in engs]
    assert sum(s["committed_n"] for s in stats) == o["commits"]
    assert gvt == o["gvt"]
    assert sum(s["remote_sent_n"] for s in stats) > 0
    chk = 0
    for e in engs:
        chk ^= e.digest()["checksum"]
    assert chk == o["checksum"]


def test_partitioned_default_subtime_and_rewind(emu_api, engine_kind):
    """test/phold.cxx on 2 GPUs x (virtual ranks = 4): sequence-id tiebreak, drain/rewind thirds."""
    A, K, P, LAM, END, ranks, G = 1000, 2, 0.5, 100.0, 10000, 4, 2
    engs = make_group(emu_api, G, A, peer=True, window=0, p_remote=P, lam=LAM, end_time=END, user_subtime=0)
    for e in engs:
        e.seed_phold(K, 0, ranks)
    t0, dt = 0, END // 3
    while True:
        t1 = t0 + dt
        drain_multi(emu_api, engs, t1, True)
        for e in engs:
            e.rewind(True)
        gvt = drain_multi(emu_api, engs, t1, True)
        for e in engs:
            e.rewind(False)
        if gvt == E.T_INF:
            break
        t0 = t1
    chk = 0
    for e in engs:
        chk ^= e.digest()["checksum"]
    assert chk == 15341107330209340805
    assert sum(e.stats()["committed_n"] for e in engs) == 361624
