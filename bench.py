#!/usr/bin/env python3
"""bench.py - committed events/sec of the PDES hot path on PHOLD-T (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference on host cores
    torchrun ... bench.py --gpus N ...                                   # N > 1: one rank per GPU

Workload (config.workload): PHOLD-T = the reference's test/phold.cxx model (xorshift128+ state,
check hash, exponential delay with glibc log, uniform remote destination) with its documented
`subtime() = ray` tiebreak (test/phold.cxx:74-76), 1 048 576 LPs per GPU, 16 root events per LP,
10 % remote fraction, lambda = 100, end_time = 1000 (about 1.7e8 committed events per GPU).
Weak scaling: LPs are block-partitioned, 1M per GPU; destinations are uniform over ALL LPs, so a
fraction p*(N-1)/N of the events crosses GPUs through the per-window NCCL exchange.

One "step" = one complete optimistic drain of a freshly seeded instance: every event with
time < end_time executed, rolled back where needed, committed.
  value : whole-job committed events/sec, instance seeded on the device (inputs resident in HBM),
          timed with CUDA events on the engine's stream over exactly K steps, max over ranks.
  e2e   : same metric through the host-buffer API (pinned host state words + root columns copied
          H2D, drain, state words copied D2H) with the copies inside the timed region.
  roofline : k_pass (the execute kernel): 120 algorithmic bytes per committed event
          x commits / summed k_pass CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the reference itself (oracle/_ref, built from /root/reference) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LP_PER_GPU = int(os.environ.get("DEVA_BENCH_LP_PER_GPU", 1 << 20))
K_ROOTS = 16
P_REMOTE = float(os.environ.get("DEVA_BENCH_P_REMOTE", 0.10))
LAMBDA = 100.0
END_TIME = int(os.environ.get("DEVA_BENCH_END", 1000))
WINDOW = int(os.environ.get("DEVA_BENCH_WINDOW", 0))      # 0 = engine's adaptive policy
B_ALG = 120.0   # algorithmic bytes per committed event: 2*E + 2*S + V, E = S = V = 24 B


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def ref_binary(threads=None):
    """oracle/_ref PHOLD-T binary (BASELINE config 2 geometry) for the largest power-of-two rank
    count <= host cores (world size is a compile-time constant in the reference)."""
    ncpu = os.cpu_count() or 1
    want = threads or int(os.environ.get("DEVA_BENCH_REF_RANKS", 0)) or ncpu
    for R in (64, 32, 16, 8, 4, 2):
        p = os.path.join(ROOT, "oracle", "_ref", f"ref_phold_t_cfg2_R{R}")
        if R <= want and os.path.exists(p):
            return p, R
    return None, 0


REF_LEAD_IN = 200   # untimed lead-in [0, 200): the start-up transient (16 roots/LP inside [0,16)) is not steady state


def run_reference_segments(n_seg, dt, timeout=900):
    exe, R = ref_binary()
    if not exe:
        return None
    env = dict(os.environ, pdes_heartbeat="0", PHOLD_T_SEGMENTS=f"{n_seg}:{dt}:{REF_LEAD_IN}", PHOLD_T_END=str(10**9))
    out = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    segs = []
    for line in out.stdout.splitlines():
        try:
            j = json.loads(line)
        except Exception:
            continue
        if "segment" in j:
            segs.append(j)
    return dict(ranks=R, segments=segs, rc=out.returncode)


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref, UNMODIFIED runtime built from
    /root/reference), all host threads, steps = successive drain(t_end) slices of one simulation."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    K, W = args.steps, args.warmup
    per_step_s = min(15.0, max(1.0, 120.0 / (K + W)))
    dt = max(2, int(round(per_step_s * 1.0e6 / (LP_PER_GPU * K_ROOTS / LAMBDA))))
    res = run_reference_segments(K + W, dt)
    if not res or len(res["segments"]) < K + W:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binary missing or failed (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    timed = res["segments"][W:]
    commits = sum(s["commits"] for s in timed)
    secs = sum(s["drain_secs"] for s in timed)
    val = commits / secs
    sample = (f"{K} timed drain(t_end) slices of {dt} sim-time units after an untimed lead-in [0,{REF_LEAD_IN}) and {W} warm-up slices, 1M-LP PHOLD-T "
              f"(single-GPU geometry; also used for N>1, where the larger LP count would only slow the CPU run)")
    line = {
        "impl": "reference", "metric": "committed events/sec (PHOLD-T)", "value": val, "unit": "events/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "events/s", "cores": res["ranks"], "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n):
    return {"workload": "PHOLD-T (test/phold.cxx model, subtime()=ray): %d LPs/GPU x %d GPU, %d roots/LP, p_remote=%.2f, "
                        "lambda=%g, end_time=%d" % (LP_PER_GPU, n, K_ROOTS, P_REMOTE, LAMBDA, END_TIME),
            "lps_total": LP_PER_GPU * n, "events_live": LP_PER_GPU * n * K_ROOTS, "partition": "block, %d LPs per GPU" % LP_PER_GPU,
            "cross_gpu_fraction": P_REMOTE * (n - 1) / n, "window": "adaptive" if WINDOW == 0 else WINDOW,
            "l2": "working set per window (pending 2 GiB + undo log 3 GiB per GPU) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from devastator_b200 import engine as E

    N = args.gpus
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != N:
        raise SystemExit(f"--gpus {N} but WORLD_SIZE={world}: launch N>1 with torchrun, one rank per GPU")
    torch.cuda.set_device(local)
    if N > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lp_n = LP_PER_GPU * N
    eng = E.Engine(E.MODEL_PHOLD_TEST, lp_n, lp_begin=rank * LP_PER_GPU, lp_count=LP_PER_GPU, device=local, n_gpus=N,
                   gpu_rank=rank, window=WINDOW, p_remote=P_REMOTE, lam=LAMBDA, end_time=END_TIME, user_subtime=1)
    if N > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(E.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(bytes(uid.cpu().numpy().tobytes()))
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=torch.device("cuda", local))

    def sync_all():
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()

    def allmax(x):
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def step_device():
        eng.seed_phold(K_ROOTS, 1, 1)
        return eng.drain()

    # ---- value: inputs resident in HBM ----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    s0 = eng.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        gvt = step_device()
    ev1.record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    s1 = eng.stats()
    clocks = sampler.stop() if rank == 0 else None
    ms = allmax(ev0.elapsed_time(ev1))
    commits_local = s1["committed_n"] - s0["committed_n"]
    commits = allsum(commits_local)
    value = commits / (ms * 1e-3)
    kernel_ms = s1["exec_kernel_ms"] - s0["exec_kernel_ms"]
    passes = s1["passes"] - s0["passes"]
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    executed = s1["executed_n"] - s0["executed_n"]
    digest = eng.digest()

    # ---- e2e: host buffers, copies inside the timed region --------------------------------------
    st_host = torch.from_numpy(E.phold_initial_state_np(E.MODEL_PHOLD_TEST, rank * LP_PER_GPU, LP_PER_GPU)).pin_memory()
    cols = [torch.from_numpy(np.ascontiguousarray(c)).pin_memory()
            for c in E.phold_roots(E.MODEL_PHOLD_TEST, lp_n, K_ROOTS, rank * LP_PER_GPU, LP_PER_GPU)]
    st_out = torch.empty_like(st_host).pin_memory()
    h2d = st_host.numel() * 8 + sum(c.numel() * c.element_size() for c in cols)
    d2h = st_out.numel() * 8

    st_np, st_out_np, cols_np = st_host.numpy(), st_out.numpy(), [c.numpy() for c in cols]
    dbg = os.environ.get("DEVA_BENCH_DEBUG")

    def step_host():
        t = [time.perf_counter()]
        eng.reset(); t.append(time.perf_counter())
        eng.set_state(st_np); t.append(time.perf_counter())
        eng.root_events(*cols_np); t.append(time.perf_counter())
        g = eng.drain(); t.append(time.perf_counter())
        eng.get_state(out=st_out_np); t.append(time.perf_counter())
        if dbg and rank == 0:
            sys.stderr.write("e2e step ms reset/set_state/roots/drain/get_state: %s\n" % [round((b - a) * 1e3, 2) for a, b in zip(t, t[1:])])
        return g, eng.stats()

    e2e_steps = max(1, min(args.steps, 3))
    step_host()
    sync_all()
    se0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    sync_all()
    e2e_s = allmax(time.perf_counter() - t0)
    se1 = eng.stats()
    e2e_commits = allsum(se1["committed_n"] - se0["committed_n"])
    e2e_val = e2e_commits / e2e_s
    # the host path and the device-seeded path must end in the same state
    chk_host = int(np.bitwise_xor.reduce(st_out.numpy()[:, 2]))
    assert chk_host == digest["checksum"], "host-buffer path and device-seeded path disagree"

    peak, peak_kind = hbm_peak()
    achieved = B_ALG * commits_local / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    line = {
        "metric": "committed events/sec (PHOLD-T)", "value": value, "unit": "events/s", "n_gpus": N,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
        "config": workload_config(N),
        "e2e": {"value": e2e_val, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": measured_traffic(),
                     "kernel": "k_pass<PholdTest>", "bytes_per_event": B_ALG, "launches": int(passes),
                     "avg_launch_ms": kernel_ms / passes if passes else None, "kernel_share_of_step": kernel_ms / ms,
                     "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback"},
        "clocks": clocks,
        "parity": {"commits_per_step": commits // args.steps, "executed_per_step": executed // args.steps,
                   "efficiency": commits_local / executed if executed else None, "final_gvt": gvt,
                   "checksum_rank0": digest["checksum"], "wall_ms_per_step": 1e3 * t_wall / args.steps,
                   "windows_per_step": passes / args.steps, "window_last": s1["window_last"]},
    }
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def measured_traffic():
    """dram bytes per k_pass launch from the committed ncu capture (profiles/*.json), or None."""
    p = os.path.join(ROOT, "profiles", "k_pass_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"]
    except Exception:
        return None


def cpu_baseline():
    """Bounded sample (about 10-30 s) of the same workload on the reference itself, host cores."""
    dt = 40
    res = run_reference_segments(3, dt, timeout=600)
    if not res or len(res["segments"]) < 3:
        return {"value": None, "unit": "events/s", "cores": 0, "kind": "reference", "sample": "unavailable: oracle/_ref binary missing"}
    timed = res["segments"][1:]
    commits = sum(s["commits"] for s in timed)
    secs = sum(s["drain_secs"] for s in timed)
    return {"value": commits / secs, "unit": "events/s", "cores": res["ranks"], "kind": "reference",
            "sample": f"reference binary (oracle/_ref, {res['ranks']} rank threads of {os.cpu_count()} host CPUs): 2 timed drain(t_end) "
                      f"slices of {dt} sim-time units ({commits} commits in {secs:.1f} s) after an untimed lead-in [0,{REF_LEAD_IN}) and "
                      f"1 warm-up slice, 1M-LP PHOLD-T"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
