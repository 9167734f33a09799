"""Several engines in ONE process (deva_b200_comm_init_group: peer access instead of CUDA IPC + NCCL), drained
concurrently by one host thread each - what the C++ drop-in does with DEVA_B200_GPUS = N.
CPU suite: the emulation (the concurrent drains meet in its lockstep drain).  GPU suite: real GPUs, when the box
has at least two; every round of the drain is closed by the GPUs among themselves over NVLink."""
import threading

import numpy as np
import pytest

import oracle_lib as O
from devastator_b200 import engine as E


def run_group(api, G, A, K, P, LAM, END, window=0, pause=None, devices=None):
    per = (A + G - 1) // G
    engs = [E.Engine(E.MODEL_PHOLD_TEST, A, lp_begin=g * per, lp_count=min(per, A - g * per), n_gpus=G, gpu_rank=g,
                     device=devices[g] if devices else 0, window=window, p_remote=P, lam=LAM, end_time=END, api=api) for g in range(G)]
    E.comm_init_group(engs)
    for e in engs:
        e.seed_phold(K, 1, 1)

    def drain_all(t_end):
        gvts, errs = [None] * G, [None] * G

        def work(g):
            try:
                gvts[g] = engs[g].drain(t_end)
            except Exception as ex:   # noqa: BLE001
                errs[g] = ex
        ts = [threading.Thread(target=work, args=(g,)) for g in range(G)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not any(errs), errs
        assert len(set(gvts)) == 1
        return gvts[0]

    out = {}
    if pause is not None:
        out["gvt_pause"] = drain_all(pause)
        out["pending_pause"] = sum(e.digest()["pending"] for e in engs)
    out["gvt"] = drain_all(E.T_INF)
    out["state"] = np.concatenate([e.get_state() for e in engs])
    out["commits"] = sum(e.stats()["committed_n"] for e in engs)
    out["crossed"] = sum(e.stats()["remote_sent_n"] for e in engs)
    for e in engs:
        e.close()
    return out


def check(out, A, K, P, LAM, END, pause=None):
    o, ost = O.run(A, K, P, LAM, END)
    np.testing.assert_array_equal(out["state"], ost)
    assert out["commits"] == o["commits"] and out["gvt"] == o["gvt"] and out["crossed"] > 0
    if pause is not None:
        op, _ = O.run(A, K, P, LAM, END, drain_end=pause)
        assert out["gvt_pause"] == op["gvt"] and out["pending_pause"] == op["pending"]


@pytest.mark.parametrize("G", [2, 3])
def test_group_of_engines_in_one_process_emu(emu_api, monkeypatch, G):
    monkeypatch.setenv("DEVA_B200_ENGINE", "tile")
    A, K, P, LAM, END = 1000, 16, 0.5, 100.0, 300
    check(run_group(emu_api, G, A, K, P, LAM, END, pause=150), A, K, P, LAM, END, pause=150)


@pytest.mark.gpu
def test_group_of_engines_in_one_process_gpus(cuda_api, monkeypatch):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    monkeypatch.setenv("DEVA_B200_ENGINE", "tile")
    for G in sorted({2, min(n, 8)}):
        for (A, K, P, LAM, END, window, pause) in [(1000, 16, 0.5, 100.0, 300, 20, 150), (100000, 16, 0.5, 100.0, 300, 0, None)]:
            check(run_group(cuda_api, G, A, K, P, LAM, END, window=window, pause=pause, devices=list(range(G))), A, K, P, LAM, END, pause=pause)
