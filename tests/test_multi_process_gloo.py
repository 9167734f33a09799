"""World-size-2 tests of the multi-GPU protocols with REAL processes and a real transport (gloo):
each process owns one block of LPs in an emulated engine (tests/emu).

tile engine (default): the round sequence of k_tile_step + k_tile_xchg - step, ship the records stored
for the peer, append them, exchange the PeerMsg figures, run the epoch state machine on every process
with the same sums - with gloo in place of the NVLink stores and the device-side mailboxes.

lp engine: the per-window sequence of engine.cu's run_window() - execute pass, exchange the per-peer
counts, exchange the staged records, deliver them, min-allreduce the GVT candidate, sum-allreduce the
window statistics, lockstep window policy - with torch.distributed in place of NCCL.
Results must equal the oracle."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_INF = 2**64 - 1


def _bind(api):
    L = api.lib
    P = C.c_void_p
    L.emu_local_min_head.argtypes = [P]; L.emu_local_min_head.restype = C.c_uint64
    L.emu_pass_begin.argtypes = [P, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.emu_send_count.argtypes = [P, C.c_int]; L.emu_send_count.restype = C.c_uint32
    L.emu_copy_send.argtypes = [P, C.c_int, P]; L.emu_copy_send.restype = None
    L.emu_deliver.argtypes = [P, P, C.c_uint32]
    L.emu_pass_end.argtypes = [P]; L.emu_pass_end.restype = C.c_uint32
    L.emu_emit_min.argtypes = [P]; L.emu_emit_min.restype = C.c_uint64
    L.emu_wake_min.argtypes = [P]; L.emu_wake_min.restype = C.c_uint64
    L.emu_upper_bound.argtypes = [P, C.c_uint64, C.c_uint64]; L.emu_upper_bound.restype = C.c_uint64
    L.emu_policy_update.argtypes = [P, C.c_uint64, C.c_uint64]; L.emu_policy_update.restype = None


def _u64_min(x):
    # gloo has no uint64: order-preserving map to int64 by flipping the sign bit
    t = torch.tensor([np.uint64(x ^ (1 << 63)).astype(np.int64)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(np.int64(t.item()).astype(np.uint64)) ^ (1 << 63)


def _sum(*xs):
    t = torch.tensor(list(xs), dtype=torch.int64)
    dist.all_reduce(t)
    return [int(v) for v in t]


def _pass(api, eng, rank, world, gvt, ub, sweep, full):
    """One pass of an epoch on this rank + the exchange (engine.cu run_pass)."""
    wx, wr = C.c_uint64(), C.c_uint64()
    api.check(api.lib.emu_pass_begin(eng.h, gvt, ub, sweep, int(full), C.byref(wx), C.byref(wr)))
    # 1. counts all-to-all
    counts = torch.tensor([api.lib.emu_send_count(eng.h, g) for g in range(world)], dtype=torch.int64)
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)
    # 2. payload exchange (pairwise send/recv), 3. deliver (+ wake LPs hit inside the window)
    for peer in range(world):
        if peer == rank:
            continue
        ns, nr = int(counts[peer]), int(allc[peer][rank])
        sbuf = torch.zeros(max(ns, 1) * 32, dtype=torch.uint8)
        if ns:
            api.lib.emu_copy_send(eng.h, peer, sbuf.data_ptr())
        rbuf = torch.zeros(max(nr, 1) * 32, dtype=torch.uint8)
        if rank < peer:
            if ns: dist.send(sbuf[:ns * 32], peer)
            if nr: dist.recv(rbuf[:nr * 32], peer)
        else:
            if nr: dist.recv(rbuf[:nr * 32], peer)
            if ns: dist.send(sbuf[:ns * 32], peer)
        if nr:
            api.check(api.lib.emu_deliver(eng.h, rbuf.data_ptr(), nr))
    # 4. statistics and woken-LP counts summed over ranks (same decisions on every rank)
    emit_min = _u64_min(api.lib.emu_emit_min(eng.h))
    woken = api.lib.emu_pass_end(eng.h)
    wxs, wrs, wk = _sum(wx.value, wr.value, woken)
    return wxs, wrs, wk, emit_min


def _worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    os.environ["DEVA_B200_ENGINE"] = "lp"
    from devastator_b200 import engine as E
    api = E.Api(C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libdeva_b200_emu.so")))
    _bind(api)
    A = cfg["A"]
    per = (A + world - 1) // world
    cnt = min(per, A - rank * per)
    eng = E.Engine(E.MODEL_PHOLD_TEST, A, lp_begin=rank * per, lp_count=cnt, n_gpus=world, gpu_rank=rank, window=cfg["window"],
                   p_remote=cfg["P"], lam=cfg["LAM"], end_time=cfg["END"], api=api)
    eng.seed_phold(cfg["K"], 1, 1)
    t_end = T_INF
    gvt = _u64_min(api.lib.emu_local_min_head(eng.h))
    windows = 0
    while gvt < t_end:                      # epochs, as deva_b200_drain drives them
        ub = api.lib.emu_upper_bound(eng.h, gvt, t_end)
        ex, rb, woken, emit_min = _pass(api, eng, rank, world, gvt, ub, 0, True)
        while woken > api.lib.emu_wake_min(eng.h):
            x, r, woken, emit_min = _pass(api, eng, rank, world, gvt, ub, 0, False)
            ex += x; rb += r
        gvt = _u64_min(api.lib.emu_local_min_head(eng.h))
        if woken > 0 and emit_min < gvt:
            gvt = emit_min
        api.lib.emu_policy_update(eng.h, ex, rb)
        windows += 1
        assert windows < 100000
    wx, _, woken, _ = _pass(api, eng, rank, world, gvt, t_end, 1, True)     # final sweep
    assert wx == 0 and woken == 0
    gvt = _u64_min(api.lib.emu_local_min_head(eng.h))   # exact: the sweep may have annihilated pending pairs
    st = torch.from_numpy(eng.get_state().astype(np.int64))
    pad = torch.zeros((per, 3), dtype=torch.int64); pad[:cnt] = st
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    s = eng.stats()
    tot = _sum(s["committed_n"], s["remote_sent_n"])
    if rank == 0:
        full = torch.cat(gathered)[:A].numpy().astype(np.uint64)
        ret["state"] = full
        ret["commits"], ret["crossed"], ret["gvt"], ret["windows"] = tot[0], tot[1], gvt, windows
    dist.barrier()
    dist.destroy_process_group()


def _exchange_records(api, eng, rank, world, count_fn, copy_fn, deliver_fn):
    """counts all-gather, then pairwise send/recv of the 32-byte records, then deliver"""
    counts = torch.tensor([0 if g == rank else count_fn(g) for g in range(world)], dtype=torch.int64)
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)
    for peer in range(world):
        if peer == rank:
            continue
        ns, nr = int(counts[peer]), int(allc[peer][rank])
        sbuf = torch.zeros(max(ns, 1) * 32, dtype=torch.uint8)
        if ns:
            copy_fn(peer, sbuf.data_ptr())
        rbuf = torch.zeros(max(nr, 1) * 32, dtype=torch.uint8)
        if rank < peer:
            if ns: dist.send(sbuf[:ns * 32], peer)
            if nr: dist.recv(rbuf[:nr * 32], peer)
        else:
            if nr: dist.recv(rbuf[:nr * 32], peer)
            if ns: dist.send(sbuf[:ns * 32], peer)
        if nr:
            deliver_fn(rbuf.data_ptr(), nr)


def _tile_worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["DEVA_B200_ENGINE"] = "tile"
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from devastator_b200 import engine as E
    api = E.Api(C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libdeva_b200_emu.so")))
    L, P = api.lib, C.c_void_p
    L.emu_tile_msg_bytes.restype = C.c_uint32
    L.emu_tile_mp_setup.argtypes = [P]
    L.emu_tile_mp_begin.argtypes = [P, C.c_uint64]
    L.emu_tile_mp_msg.argtypes = [P, P]; L.emu_tile_mp_msg.restype = None
    L.emu_tile_mp_close.argtypes = [P, P, C.c_int, C.c_int]
    L.emu_tile_mp_step.argtypes = [P]
    L.emu_tile_mp_send.argtypes = [P, C.c_int, P]; L.emu_tile_mp_send.restype = C.c_uint32
    L.emu_tile_mp_deliver.argtypes = [P, P, C.c_uint32]
    L.emu_tile_mp_finish.argtypes = [P, C.POINTER(C.c_uint64)]
    A = cfg["A"]
    per = (A + world - 1) // world
    cnt = min(per, A - rank * per)
    eng = E.Engine(E.MODEL_PHOLD_TEST, A, lp_begin=rank * per, lp_count=cnt, n_gpus=world, gpu_rank=rank, window=cfg["window"],
                   p_remote=cfg["P"], lam=cfg["LAM"], end_time=cfg["END"], api=api)
    api.check(L.emu_tile_mp_setup(eng.h))
    eng.seed_phold(cfg["K"], 1, 1)
    nmsg = L.emu_tile_msg_bytes()

    def figures():   # every process's PeerMsg, as k_tile_xchg's mailboxes carry them
        mine = torch.zeros(nmsg, dtype=torch.uint8)
        L.emu_tile_mp_msg(eng.h, mine.data_ptr())
        allm = [torch.zeros(nmsg, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(allm, mine)
        return torch.cat(allm).contiguous()

    PH_DONE = 3
    for t_end in cfg.get("ends", [T_INF]):
        api.check(L.emu_tile_mp_begin(eng.h, t_end))
        f = figures()   # (keep the tensor alive across the call)
        phase = L.emu_tile_mp_close(eng.h, f.data_ptr(), world, 1)
        rounds = 0
        while phase != PH_DONE:
            L.emu_tile_mp_step(eng.h)
            _exchange_records(api, eng, rank, world, lambda g: L.emu_tile_mp_send(eng.h, g, None),
                              lambda g, ptr: L.emu_tile_mp_send(eng.h, g, ptr), lambda ptr, n: L.emu_tile_mp_deliver(eng.h, ptr, n))
            f = figures()
            phase = L.emu_tile_mp_close(eng.h, f.data_ptr(), world, 0)
            rounds += 1
            assert rounds < 1000000
        m = C.c_uint64()
        api.check(L.emu_tile_mp_finish(eng.h, C.byref(m)))
        gvt = _u64_min(m.value)
    st = torch.from_numpy(eng.get_state().astype(np.int64))
    pad = torch.zeros((per, 3), dtype=torch.int64); pad[:cnt] = st
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    s = eng.stats()
    tot = _sum(s["committed_n"], s["remote_sent_n"])
    if rank == 0:
        ret["state"] = torch.cat(gathered)[:A].numpy().astype(np.uint64)
        ret["commits"], ret["crossed"], ret["gvt"], ret["windows"] = tot[0], tot[1], gvt, rounds
    dist.barrier()
    dist.destroy_process_group()


CFGS = [dict(A=1000, K=16, P=0.5, LAM=100.0, END=300, window=20), dict(A=1000, K=16, P=0.1, LAM=100.0, END=300, window=0)]


@pytest.mark.parametrize("cfg", CFGS + [dict(A=777, K=8, P=1.0, LAM=20.0, END=200, window=0, ends=[60, 130, T_INF])],
                         ids=["p50-w20", "p10-adaptive", "allremote-segments"])
def test_two_processes_gloo_tile_engine_equal_oracle(emu_api, cfg):
    import oracle_lib as O
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_tile_worker, args=(2, port, cfg, ret), nprocs=2, join=True)
    o, ost = O.run(cfg["A"], cfg["K"], cfg["P"], cfg["LAM"], cfg["END"])
    np.testing.assert_array_equal(ret["state"], ost)
    assert ret["commits"] == o["commits"] and ret["gvt"] == o["gvt"]
    assert ret["crossed"] > 0


@pytest.mark.parametrize("cfg", CFGS, ids=["p50-w20", "p10-adaptive"])
def test_two_processes_gloo_equal_oracle(emu_api, cfg):
    import oracle_lib as O
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, cfg, ret), nprocs=2, join=True)
    o, ost = O.run(cfg["A"], cfg["K"], cfg["P"], cfg["LAM"], cfg["END"])
    np.testing.assert_array_equal(ret["state"], ost)
    assert ret["commits"] == o["commits"] and ret["gvt"] == o["gvt"]
    assert ret["crossed"] > 0
