#!/usr/bin/env python3
"""Run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
PHOLD-T block-partitioned over N engines (one process per GPU, NCCL exchange inside libdeva_b200.so),
per-LP state / commits / GVT compared with the oracle on rank 0.  Prints one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from devastator_b200 import engine as E  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_uid():   # an ncclUniqueId makes exactly one communicator
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(E.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return uid.cpu().numpy().tobytes()
    ok_all = True
    # `pause`: drain(t) first and compare the GVT returned and the pending count at the pause as well (cross-GPU
    # anti-messages beyond t are matched by the final sweep, so no positive / anti pair may be left to bound the GVT)
    cases = [dict(A=1000, K=16, P=0.5, LAM=100.0, END=300, window=20, pause=150), dict(A=1000, K=16, P=0.1, LAM=100.0, END=300, window=0),
             dict(A=100000, K=16, P=0.5, LAM=100.0, END=300, window=0), dict(A=100000, K=16, P=0.10, LAM=100.0, END=300, window=32)]
    for c in cases:
        A = c["A"]
        per = (A + world - 1) // world
        cnt = min(per, A - rank * per)
        eng = E.Engine(E.MODEL_PHOLD_TEST, A, lp_begin=rank * per, lp_count=cnt, device=local, n_gpus=world, gpu_rank=rank,
                       window=c["window"], p_remote=c["P"], lam=c["LAM"], end_time=c["END"])
        eng.comm_init(fresh_uid())
        eng.seed_phold(c["K"], 1, 1)
        pause_ok = None
        if c.get("pause"):
            gvt_p = eng.drain(c["pause"])
            pend = torch.tensor([eng.digest()["pending"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(pend)
            if rank == 0:
                import oracle_lib as O
                op, _ = O.run(A, c["K"], c["P"], c["LAM"], c["END"], drain_end=c["pause"])
                pause_ok = gvt_p == op["gvt"] and int(pend[0]) == op["pending"]
        gvt = eng.drain()
        st = torch.from_numpy(eng.get_state().astype(np.int64)).cuda()
        s = eng.stats()
        pad = torch.zeros((per, 3), dtype=torch.int64, device="cuda")
        pad[:cnt] = st
        gathered = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        tot = torch.tensor([s["committed_n"], s["executed_n"], s["remote_sent_n"], s["rolled_back_n"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        if rank == 0:
            import oracle_lib as O   # the checker
            o, ost = O.run(A, c["K"], c["P"], c["LAM"], c["END"])
            full = torch.cat(gathered)[:A].cpu().numpy().astype(np.uint64)
            ok = bool((full == ost).all()) and int(tot[0]) == o["commits"] and gvt == o["gvt"] and pause_ok is not False
            ok_all &= ok
            gold = None
            if A == 100000 and c["P"] == 0.10:   # SURVEY Appendix B golden from the reference binary
                gold = (int(tot[0]) == 6241953 and int(np.bitwise_xor.reduce(full[:, 2])) == 3349230152040954854)
                ok_all &= gold
            print(json.dumps(dict(case=c, n_gpus=world, ok=ok, reference_golden=gold, commits=int(tot[0]), executed=int(tot[1]),
                                  crossed=int(tot[2]), rolled_back=int(tot[3]), passes=s["passes"], drain_ms=round(s["drain_ms"], 2))), flush=True)
        eng.close()
        dist.barrier()
    flag = torch.tensor([1 if ok_all else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
