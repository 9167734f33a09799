#!/usr/bin/env python3
"""N-GPU probe under torchrun (one rank per GPU): QB_A LPs per GPU (weak) or QB_TOTAL LPs in all, remote
fraction QB_P; QB_LIB selects another build of the ABI (e.g. the -DDEVA_TILE_PROF one).  Golden-checked where
tests/golden/bench_goldens.json has the size.  Prints one JSON line per repetition on rank 0."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from devastator_b200 import engine as E
if os.environ.get("QB_LIB"):
    E.LIB_PATH = os.path.join(ROOT, os.environ["QB_LIB"])
rank, local, N = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = float(os.environ.get("QB_P", 0.1)); END = int(os.environ.get("QB_END", 1000)); REPS = int(os.environ.get("QB_REPS", 3))
total = int(os.environ.get("QB_TOTAL", 0)) or int(os.environ.get("QB_A", 1 << 20)) * N
per = (total + N - 1) // N
gold = {(v["lp_n"], v["p_remote"]): v for v in json.load(open(os.path.join(ROOT, "tests/golden/bench_goldens.json")))["vectors"]}
g = gold.get((total, P)) if END == 1000 else None
eng = E.Engine(E.MODEL_PHOLD_TEST, total, lp_begin=rank * per, lp_count=min(per, total - rank * per), device=local, n_gpus=N, gpu_rank=rank,
               p_remote=P, lam=100.0, end_time=END)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(E.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
eng.comm_init(bytes(uid.cpu().numpy().tobytes()))
for rep in range(REPS):
    eng.seed_phold(16, 1, 1)
    torch.cuda.synchronize(); dist.barrier()
    s0 = eng.stats(); eng.drain(); s1 = eng.stats(); d = eng.digest()
    t = torch.tensor([s1["committed_n"] - s0["committed_n"], np.uint64(d["checksum"]).astype(np.int64), s1["passes"] - s0["passes"]], dtype=torch.int64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(N)]; dist.all_gather(out, t)
    ms = torch.tensor([s1["drain_ms"] - s0["drain_ms"]], dtype=torch.float64, device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        c = sum(int(o[0]) for o in out); x = 0
        for o in out: x ^= int(o[1]) & ((1 << 64) - 1)
        par = None if not g else ("MATCH" if (c == g["commits"] and x == g["checksum"]) else "MISMATCH")
        print(json.dumps(dict(n_gpus=N, lps=total, p=P, rep=rep, commits=c, drain_ms=round(float(ms), 2), Gev_s=round(c / float(ms) / 1e6, 3), launches=int(out[0][2]), golden=par)), flush=True)
eng.close(); dist.barrier(); dist.destroy_process_group()
