#!/usr/bin/env python3
"""bench.py - committed events/sec of the PDES hot path on PHOLD-T (BASELINE.json configs[1] and configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W]                    # our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference on host cores
    torchrun ... bench.py --gpus N ...                                     # N > 1: one rank per GPU

Workloads (config.workload names the one a line is quoted on).  PHOLD-T = the reference's
test/phold.cxx model (xorshift128+ state, check hash, exponential delay through glibc's log, uniform
remote destination) with its documented `subtime() = ray` tiebreak (test/phold.cxx:74-76), 16 root
events per LP, lambda = 100, end_time = 1000:
  N = 1 : BASELINE config 2 - 1 048 576 LPs, 10 % remote                      (1.82e8 commits per step)
  N > 1 : BASELINE config 3 - 8 388 608 LPs block-partitioned over the N GPUs, 50 % remote
          (1.46e9 commits per step; a fraction 0.5 (N-1)/N of all events crosses GPUs over NVLink)
Both are also run at the other GPU counts and reported inside the same JSON line under "also":
config 3 on one GPU (the strong-scaling base of the N > 1 lines) and config 2's geometry weak-scaled
(1 M LPs per GPU, 10 % remote: the series round 1 reported).

One "step" = one complete optimistic drain of a freshly seeded instance: every event with
time < end_time executed, rolled back where needed, committed.  Seeding (building the root events,
the counterpart of the reference's root_event() calls before drain()) is outside the timed region of
`value` - the reference arm times its drain() calls the same way - and inside the one of `e2e`.
  value : whole-job committed events/sec, inputs resident in HBM (instance seeded on the device), CUDA
          events on the engine's stream around each of the K drains, summed, max over ranks.
  e2e   : same metric through the host-buffer API (pinned host state words + root columns copied H2D and
          inserted, drain, state words copied D2H), everything inside the timed region, all K steps.
  roofline : the step kernel k_tile_step (all its launches): B_alg algorithmic bytes per committed event
          (120 + 48 f, f = cross-GPU fraction) x commits / summed CUDA-event time of the
          launches, against MEASURED_PEAKS.json hbm_gbs; `traffic` from the committed ncu capture.
  parity : commits, XOR checksum, additive state hash and final GVT gathered over the ranks and ASSERTED
          against tests/golden/bench_goldens.json (reference binary / pinned oracle) - a mismatch exits 1.
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
# stdout carries ONE JSON line: NCCL's banner ("NCCL version ...", printed at NCCL_DEBUG=VERSION) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

K_ROOTS = 16
LAMBDA = 100.0
END_TIME = int(os.environ.get("DEVA_BENCH_END", 1000))
WINDOW = int(os.environ.get("DEVA_BENCH_WINDOW", 0))      # 0 = engine's adaptive policy
CFG3_LPS = int(os.environ.get("DEVA_BENCH_CFG3_LPS", 1 << 23))
CFG2_LPS = int(os.environ.get("DEVA_BENCH_CFG2_LPS", 1 << 20))


def b_alg(p_remote, n):
    """algorithmic bytes per committed event: 2*E + 2*S + V = 120 B (E = S = V = 24 B), plus 48 B for
    an event that crosses GPUs (written into the peer's receive region, read back, appended): SURVEY 8d"""
    return 120.0 + 48.0 * p_remote * (n - 1) / n


def workloads(n):
    """(tag, total LPs, remote fraction, scaling) - the first one is the line's headline"""
    cfg2 = ("cfg2", CFG2_LPS * n, 0.10, "weak")
    cfg3 = ("cfg3", CFG3_LPS, 0.50, "strong")
    forced = os.environ.get("DEVA_BENCH_WORKLOAD")
    if forced:
        return [w for w in (cfg2, cfg3) if w[0] == forced]
    return [cfg2, cfg3] if n == 1 else [cfg3, cfg2]


def workload_config(tag, lp_total, p, n):
    name = {"cfg2": "BASELINE config 2 geometry", "cfg3": "BASELINE config 3"}[tag]
    return {"workload": "%s: PHOLD-T (test/phold.cxx model, subtime()=ray), %d LPs block-partitioned over %d GPU(s) (%d per GPU), "
                        "%d roots/LP, p_remote=%.2f, lambda=%g, end_time=%d" % (name, lp_total, n, lp_total // n, K_ROOTS, p, LAMBDA, END_TIME),
            "tag": tag, "lps_total": lp_total, "events_live": lp_total * K_ROOTS, "partition": "block, %d LPs per GPU" % (lp_total // n),
            "p_remote": p, "cross_gpu_fraction": p * (n - 1) / n, "window": "adaptive" if WINDOW == 0 else WINDOW,
            "l2": "inputs larger than L2: the calendar lists and LP states touched per epoch (>= 250 MB per GPU) exceed the 126 MB L2; no flush"}


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def goldens():
    try:
        vs = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_goldens.json")))["vectors"]
    except Exception:
        return {}
    out = {}
    for v in vs:
        if v["end_time"] == END_TIME and v["k"] == K_ROOTS:
            out.setdefault((v["lp_n"], v["p_remote"]), v)   # (the reference-binary vector comes first)
    return out


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
            self.lines.append((time.perf_counter(), line))

    def stop(self, t0=None, t1=None):
        """median SM clock / reasons over the samples that arrived inside [t0, t1] (the timed region; the
        sampler is started before the warm-up so that nvidia-smi is already streaming by then)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(lines):
            sm, mx, reasons = [], [], set()
            for _, line in lines:
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
            return sm, mx, reasons
        inside = [x for x in self.lines if t0 is None or (t0 <= x[0] <= t1 + 0.12)]
        window = "timed region"
        if not digest(inside)[0]:       # a timed region shorter than the sampling period: the samples around it (warm-up = same load)
            inside, window = self.lines, "warm-up + timed region (the timed region was shorter than the 100 ms sampling period)"
        sm, mx, reasons = digest(inside)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


# ---- the reference on the host cores -------------------------------------------------------------
def ref_binary(tag):
    """oracle/_ref PHOLD-T binary of the workload for the largest power-of-two rank count <= host cores
    (world size is a compile-time constant in the reference)."""
    ncpu = os.cpu_count() or 1
    want = int(os.environ.get("DEVA_BENCH_REF_RANKS", 0)) or ncpu
    for R in (64, 32, 16, 8, 4, 2):
        p = os.path.join(ROOT, "oracle", "_ref", f"ref_phold_t_{tag}_R{R}")
        if R <= want and os.path.exists(p):
            return p, R
    return None, 0


# untimed lead-in [0, lead_in): the start-up transient (16 roots per LP inside [0, 16)) is not steady state
REF_LEAD_IN = {"cfg2": 200, "cfg3": 40}    # (config 3 holds 8x the events: a shorter lead-in keeps the run within minutes)
REF_LPS = {"cfg2": 1 << 20, "cfg3": 1 << 23}      # what the binaries were compiled for (oracle/build_ref.py)
REF_P = {"cfg2": 0.10, "cfg3": 0.50}


def run_reference_segments(tag, n_seg, dt, timeout=1500):
    exe, R = ref_binary(tag)
    if not exe:
        return None
    env = dict(os.environ, pdes_heartbeat="0", PHOLD_T_SEGMENTS=f"{n_seg}:{dt}:{REF_LEAD_IN[tag]}", PHOLD_T_END=str(10**9))
    try:
        out = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None
    segs = []
    for line in out.stdout.splitlines():
        try:
            j = json.loads(line)
        except Exception:
            continue
        if "segment" in j:
            segs.append(j)
    return dict(ranks=R, segments=segs, rc=out.returncode, tag=tag)


def reference_sample(tag, n_timed, n_warm, events_per_seg):
    """n_warm + n_timed drain(t_end) slices of the reference binary of workload `tag`; falls back to the
    config-2 binary when the config-3 one cannot run on this host (memory), and says so."""
    for t in ([tag] if tag == "cfg2" else [tag, "cfg2"]):
        dt = max(2, int(round(events_per_seg / (REF_LPS[t] * K_ROOTS / LAMBDA))))
        res = run_reference_segments(t, n_warm + n_timed, dt)
        if res and len(res["segments"]) >= n_warm + n_timed:
            timed = res["segments"][n_warm:]
            commits = sum(s["commits"] for s in timed)
            secs = sum(s["drain_secs"] for s in timed)
            res.update(dt=dt, commits=commits, secs=secs, value=commits / secs)
            return res
    return None


def sample_text(res, n_timed, n_warm):
    t = res["tag"]
    return (f"reference binary oracle/_ref/ref_phold_t_{t}_R{res['ranks']} ({res['ranks']} rank threads of {os.cpu_count()} host CPUs; "
            f"{REF_LPS[t]} LPs, p_remote={REF_P[t]}): {n_timed} timed drain(t_end) slices of {res['dt']} sim-time units "
            f"({res['commits']} commits in {res['secs']:.1f} s) after an untimed lead-in [0,{REF_LEAD_IN[t]}) and {n_warm} warm-up slice(s)")


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref, UNMODIFIED runtime built from
    /root/reference), all host threads, steps = successive drain(t_end) slices of one simulation of the
    workload our arm's headline is quoted on (config 2 at N = 1, config 3 at N > 1)."""
    if int(os.environ.get("RANK", 0)) != 0:
        return 0
    K, W = args.steps, args.warmup
    tag, lp_total, p, scaling = workloads(args.gpus)[0]
    # about 1e7 (config 2) / 3e6 (config 3) committed events per slice, 1-3 s of host time each: the whole run ends within minutes
    res = reference_sample(tag, K, W, float(os.environ.get("DEVA_BENCH_REF_EVENTS_PER_STEP", 1.0e7 if tag == "cfg2" else 3.0e6)))
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binary missing or failed (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    t = res["tag"]
    cfg = workload_config(t, REF_LPS[t], REF_P[t], args.gpus)
    cfg["workload"] = (f"{'BASELINE config 3' if t == 'cfg3' else 'BASELINE config 2 geometry'}: PHOLD-T, {REF_LPS[t]} LPs on the host cores "
                       f"({res['ranks']} rank threads), {K_ROOTS} roots/LP, p_remote={REF_P[t]:.2f}, lambda={LAMBDA:g}, drain(t_end) slices")
    cfg["partition"] = "block over %d host rank threads" % res["ranks"]
    cfg["cross_gpu_fraction"] = 0.0
    if t != tag:
        cfg["note"] = f"the {tag} binary could not run on this host (memory / time-out): the config-2 binary was timed instead"
    line = {
        "impl": "reference", "metric": "committed events/sec (PHOLD-T)", "value": res["value"], "unit": "events/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * res["secs"] / K, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": res["value"], "unit": "events/s", "cores": res["ranks"], "kind": "reference", "sample": sample_text(res, K, W)},
        "e2e": {"value": res["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(tag):
    """Bounded sample (about 10-30 s of drain time) of the same workload on the reference itself, host cores."""
    res = reference_sample(tag, 2, 1, float(os.environ.get("DEVA_BENCH_CPU_EVENTS_PER_STEP", 3.0e7 if tag == "cfg2" else 1.0e7)))
    if not res:
        return {"value": None, "unit": "events/s", "cores": 0, "kind": "reference", "sample": "unavailable: oracle/_ref binary missing"}
    return {"value": res["value"], "unit": "events/s", "cores": res["ranks"], "kind": "reference", "sample": sample_text(res, 2, 1)}


def measured_traffic():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture of this very
    configuration (profiles/k_tile_step_traffic.json, written by tools/ncu_traffic.py), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "k_tile_step_traffic.json")))
    except Exception:
        return None


# ---- our arm -------------------------------------------------------------------------------------
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
    dev = torch.device("cuda", local)
    if N > 1:
        stale = os.path.join("/tmp", "deva_bench_cpu_done_%s" % os.environ.get("MASTER_PORT", "0"))
        if rank == 0 and os.path.exists(stale):
            os.remove(stale)
        dist.init_process_group("nccl", device_id=dev)
    gold = goldens()
    MASK = (1 << 64) - 1

    def sync_all():
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()

    def gather_u64(xs):
        """every rank's list of u64 values -> [rank][i] on every rank"""
        if N == 1:
            return [list(xs)]
        t = torch.tensor([np.uint64(x).astype(np.int64) for x in xs], dtype=torch.int64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(N)]
        dist.all_gather(out, t)
        return [[int(v) & MASK for v in o.tolist()] for o in out]

    def allmax(x):
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run(tag, lp_total, p, scaling, headline):
        per = (lp_total + N - 1) // N
        cnt = min(per, lp_total - rank * per)
        eng = E.Engine(E.MODEL_PHOLD_TEST, lp_total, lp_begin=rank * per, lp_count=cnt, device=local, n_gpus=N,
                       gpu_rank=rank, window=WINDOW, p_remote=p, lam=LAMBDA, end_time=END_TIME, user_subtime=1)
        if N > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(E.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            eng.comm_init(bytes(uid.cpu().numpy().tobytes()))
        stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
        K, W = args.steps, args.warmup

        # ---- value: inputs resident in HBM, drains timed with CUDA events on the engine's stream
        sampler = ClockSampler(local) if (headline and rank == 0) else None
        if sampler:
            sampler.start()
        for _ in range(W):
            eng.seed_phold(K_ROOTS, 1, 1)
            eng.drain()
        sync_all()
        s0 = eng.stats()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        seed_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t_wall0 = time.perf_counter()
        gvt = None
        for k in range(K):
            seed_evs[k][0].record(stream)
            eng.seed_phold(K_ROOTS, 1, 1)
            seed_evs[k][1].record(stream)
            if N > 1:
                sync_all()          # every rank's step starts together (the drain itself synchronises the ranks on the device)
            evs[k][0].record(stream)
            gvt = eng.drain()
            evs[k][1].record(stream)
        sync_all()
        t_wall = time.perf_counter() - t_wall0
        s1 = eng.stats()
        clocks = sampler.stop(t_wall0, t_wall0 + t_wall) if sampler else None
        ms = allmax(sum(a.elapsed_time(b) for a, b in evs))
        seed_ms = allmax(sum(a.elapsed_time(b) for a, b in seed_evs))
        d = eng.digest()
        commits_local = s1["committed_n"] - s0["committed_n"]
        executed_local = s1["executed_n"] - s0["executed_n"]
        kernel_ms = s1["exec_kernel_ms"] - s0["exec_kernel_ms"]
        passes = s1["passes"] - s0["passes"]
        launches = s1["kernel_launches"] - s0["kernel_launches"]
        g = gather_u64([commits_local, executed_local, d["checksum"], d["state_hash"], gvt, d["pending"], s1["rolled_back_n"] - s0["rolled_back_n"],
                        s1["remote_sent_n"] - s0["remote_sent_n"]])
        commits = sum(r[0] for r in g)
        executed = sum(r[1] for r in g)
        checksum = 0
        for r in g:
            checksum ^= r[2]
        state_hash = sum(r[3] for r in g) & MASK
        value = commits / (ms * 1e-3)
        # ---- parity against the goldens of the benchmarked configuration
        gv = gold.get((lp_total, p))
        if gv is None:
            verdict = "none (no golden for this size)"
        else:
            ok = (commits == K * gv["commits"] and checksum == gv["checksum"] and state_hash == gv["state_hash"] and
                  all(r[4] == gv["gvt"] for r in g) and all(r[5] == 0 for r in g) and s1["deterministic"] == gv["deterministic"])
            verdict = "match" if ok else "MISMATCH"
        parity = {"golden": verdict, "golden_source": gv["source"] if gv else None, "commits_per_step": commits // K,
                  "checksum": checksum, "state_hash": state_hash, "final_gvt": gvt, "executed_per_step": executed // K,
                  "efficiency": commits / executed if executed else None, "rolled_back_per_step": sum(r[6] for r in g) // K,
                  "crossed_gpus_per_step": sum(r[7] for r in g) // K}

        out = {"tag": tag, "value": value, "ms_per_step": ms / K, "seed_ms_per_step": seed_ms / K, "scaling": scaling,
               "config": workload_config(tag, lp_total, p, N), "parity": parity, "gpu_launches": int(launches),
               "wall_ms_per_step": 1e3 * t_wall / K, "launches_per_step": passes / K, "window_last": s1["window_last"]}
        peak, peak_kind = hbm_peak()
        B = b_alg(p, N)
        achieved = B * commits_local / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                           "traffic": None, "kernel": "k_tile_step<PholdTest> (an epoch's full launch + its follow-up launches)",
                           "bytes_per_event": B, "launches": int(passes), "avg_launch_ms": kernel_ms / passes if passes else None,
                           "kernel_share_of_step": kernel_ms / ms if ms else None, "peak_kind": peak_kind}
        if clocks is not None:
            out["clocks"] = clocks

        # ---- e2e: host buffers, copies inside the timed region, all K steps
        if headline or args.e2e_all:
            first = rank * per
            st_host = torch.from_numpy(E.phold_initial_state_np(E.MODEL_PHOLD_TEST, first, cnt)).pin_memory()
            cols = [torch.from_numpy(np.ascontiguousarray(c)).pin_memory() for c in E.phold_roots(E.MODEL_PHOLD_TEST, lp_total, K_ROOTS, first, cnt)]
            st_out = torch.empty_like(st_host).pin_memory()
            h2d = st_host.numel() * 8 + sum(c.numel() * c.element_size() for c in cols)
            d2h = st_out.numel() * 8
            st_np, st_out_np, cols_np = st_host.numpy(), st_out.numpy(), [c.numpy() for c in cols]

            parts = {"reset": 0.0, "h2d_state": 0.0, "h2d_roots_and_insert": 0.0, "drain": 0.0, "d2h_state": 0.0}

            def step_host():      # (every call but reset returns synchronised: reset's device time shows under h2d_state)
                t = [time.perf_counter()]
                eng.reset(); t.append(time.perf_counter())
                eng.set_state(st_np); t.append(time.perf_counter())
                eng.root_events(*cols_np); t.append(time.perf_counter())
                eng.drain(); t.append(time.perf_counter())
                eng.get_state(out=st_out_np); t.append(time.perf_counter())
                for k, a, b in zip(parts, t, t[1:]):
                    parts[k] += b - a

            step_host()
            sync_all()
            se0 = eng.stats()
            for k in parts:
                parts[k] = 0.0
            t0 = time.perf_counter()
            for _ in range(K):
                step_host()
            sync_all()
            e2e_s = allmax(time.perf_counter() - t0)
            se1 = eng.stats()
            ge = gather_u64([se1["committed_n"] - se0["committed_n"], int(np.bitwise_xor.reduce(st_out_np[:, 2]))])
            e2e_commits = sum(r[0] for r in ge)
            chk = 0
            for r in ge:
                chk ^= r[1]
            # the host path and the device-seeded path must end in the same state (and in the golden's)
            if chk != checksum or e2e_commits != commits:
                parity["golden"] = "MISMATCH"
                parity["e2e_path"] = "host-buffer path and device-seeded path disagree"
            out["e2e"] = {"value": e2e_commits / e2e_s, "unit": "events/s", "h2d_bytes_per_step": h2d * N if N > 1 else h2d,
                          "d2h_bytes_per_step": d2h * N if N > 1 else d2h, "steps": K, "bytes_are": "summed over the ranks",
                          "ms_per_step_by_call": {k: round(1e3 * v / K, 3) for k, v in parts.items()}, "by_call_is": "rank 0's wall clock"}
        eng.close()
        del eng
        torch.cuda.empty_cache()
        if N > 1:
            dist.barrier()
        return out

    results = []
    wl = workloads(N)
    for i, (tag, lp_total, p, scaling) in enumerate(wl):
        results.append(run(tag, lp_total, p, scaling, headline=(i == 0)))
    head = results[0]
    line = {
        "metric": "committed events/sec (PHOLD-T)", "value": head["value"], "unit": "events/s", "n_gpus": N,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": head["scaling"], "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
        "config": head["config"], "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"],
        "roofline": head["roofline"], "clocks": head.get("clocks"), "parity": head["parity"],
        "seed_ms_per_step": head["seed_ms_per_step"], "wall_ms_per_step": head["wall_ms_per_step"],
        "launches_per_step": head["launches_per_step"], "window_last": head["window_last"],
        "also": {r["tag"]: {k: r[k] for k in ("value", "ms_per_step", "seed_ms_per_step", "scaling", "config", "parity", "roofline", "gpu_launches") if k in r}
                 | ({"e2e": r["e2e"]} if "e2e" in r else {}) for r in results[1:]},
    }
    tr = measured_traffic()
    if tr and N == 1:
        line["roofline"]["traffic"] = tr.get("dram_bytes_per_full_launch")
        line["roofline"]["traffic_detail"] = tr
    # the reference on the host cores, rank 0; the other ranks sleep meanwhile (a rank spinning in an NCCL
    # barrier would take a core away from the reference's pinned rank threads)
    flag = os.path.join("/tmp", "deva_bench_cpu_done_%s" % os.environ.get("MASTER_PORT", "0"))
    if not args.no_cpu_baseline:
        if rank == 0:
            line["cpu_baseline"] = cpu_baseline(head["tag"])
            if N > 1:
                open(flag, "w").close()
        else:
            t_wait = time.time()
            while not os.path.exists(flag) and time.time() - t_wait < 1800:
                time.sleep(0.25)
    bad = [r["tag"] for r in results if r["parity"]["golden"] == "MISMATCH"]
    if rank == 0:
        emit_report_rows(results, N)
        print(json.dumps(line), flush=True)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0 and os.path.exists(flag):
            os.remove(flag)
    if bad:
        sys.stderr.write("PARITY MISMATCH against tests/golden/bench_goldens.json: %s\n" % bad)
        return 1
    return 0


def emit_report_rows(results, n_gpus):
    """With `report_file` set (the reference's own variable, bench/util/report.cxx:25), every workload of the
    run is appended as one `row(xs=dict(...), ys=dict(...))` in the format the reference's bench/util/show.py
    and table.py read, with the y names of bench/phold.cxx:176-179 (a "rank" is a GPU here)."""
    path = os.environ.get("report_file")
    if not path:
        return
    out = sys.stderr if path == "-" else open(path, "a")
    if out is not sys.stderr:
        out.write("# Opened for append at %s\n" % time.strftime("%Y-%m-%d %X"))
    for r in results:
        c, p = r["config"], r["parity"]
        xs = dict(app="phold", engine="deva_b200", gpus=n_gpus, ranks=n_gpus, workload=c["tag"], lp_per_rank=c["lps_total"] // n_gpus,
                  ray_per_lp=K_ROOTS, p_remote=c["p_remote"], end_time=END_TIME)
        secs = r["ms_per_step"] * 1e-3
        ys = dict(execute_per_rank_per_sec=p["executed_per_step"] / secs / n_gpus, commit_per_rank_per_sec=p["commits_per_step"] / secs / n_gpus,
                  deterministic=1 if p["golden"] != "MISMATCH" else 0)
        kw = lambda d: ", ".join("%s=%r" % kv for kv in d.items())
        out.write("row(xs=dict(%s),\n    ys=dict(%s))\n" % (kw(xs), kw(ys)))
    if out is not sys.stderr:
        out.close()
        sys.stderr.write("Report written to '%s'.\n" % path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-all", action="store_true", help="measure e2e for the secondary workloads, too")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
