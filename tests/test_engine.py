"""Engine parity through the C ABI (include/deva_b200.h).

backend "emu"  : tests/emu host emulation of the SAME window code (CPU suite, logic only)
backend "cuda" : libdeva_b200.so on cuda:0 (the product; -m gpu)

Mirrors the reference's own test strategy: test/phold.cxx compares checksums
across plain drains and drain/rewind thirds; here every run is additionally compared word for word
with the oracle, for several optimism-window widths (the window must never change committed
results)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import scenarios as S
from devastator_b200 import engine as E

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def api(request, engine_kind):
    return request.getfixturevalue("emu_api" if request.param == "emu" else "cuda_api")


# ---- window width never changes results ----------------------------------------------------------
@pytest.mark.parametrize("window", [1, 3, 20, 100, 0])
@pytest.mark.parametrize("cfg", [
    dict(A=64, K=4, P=0.5, LAM=10.0, END=400),        # tiny, heavy cross-LP traffic
    dict(A=1000, K=16, P=0.10, LAM=100.0, END=300),   # "small"
    dict(A=1000, K=16, P=0.50, LAM=100.0, END=300),   # "smallq"
], ids=["tiny", "small", "smallq"])
def test_phold_t_matches_oracle(api, cfg, window):
    s, o = S.run_phold_t(api, cfg["A"], cfg["K"], cfg["P"], cfg["LAM"], cfg["END"], window=window)
    if window == 1:   # dt >= 1: a 1-tick window is conservative, nothing may roll back
        assert s["rolled_back_n"] == 0 and s["executed_n"] == o["commits"]
    if window == 100:  # ... and a window of a mean delay (or ten) speculates: the rollback / anti-message path was taken
        assert s["rolled_back_n"] > 0 and s["anti_sent_n"] > 0 and s["executed_n"] > o["commits"]


# ---- how early an epoch's follow-up passes stop never changes results either ----------------------
@pytest.mark.parametrize("wake_min", ["0", "3", "50", "100000"])
def test_epoch_cut_threshold_never_changes_results(api, monkeypatch, wake_min):
    monkeypatch.setenv("DEVA_B200_WAKE_MIN", wake_min)
    S.run_phold_t(api, 1000, 16, 0.5, 100.0, 300, window=25)
    S.run_phold_t(api, 64, 4, 0.5, 10.0, 400, window=40)


# ---- default tiebreak = per-LP sequence ids, root subtime = local cd index (test/phold.cxx) ------
@pytest.mark.parametrize("ranks", [2, 3, 8])
@pytest.mark.parametrize("window", [1, 50, 0])
def test_reference_phold_test_config(api, ranks, window):
    """Config 1': unchanged test/phold.cxx model. 180 812 commits, checksum 15341107330209340805
    (tests/golden/ref_goldens.json)."""
    s, o = S.run_phold_t(api, 1000, 2, 0.5, 100.0, 10000, window=window, sub=0, rootdiv=0, ranks=ranks)
    assert o["commits"] == 180812 and o["checksum"] == 15341107330209340805
    gold = [g for g in S.goldens()["phold_t"] if g["tag"] == "seqid"][0]
    assert s["committed_n"] == gold["commits"]


# ---- golden fixtures straight from the reference binaries ---------------------------------------
@pytest.mark.parametrize("tag", ["tiny", "small", "smallq", "allrem", "norem", "dense"])
def test_against_reference_state_fixture(api, tag):
    g = [x for x in S.goldens()["phold_t"] if x["tag"] == tag and x["drain_end"] == 2**64 - 1][0]
    c = g["cfg"]
    want = np.load(os.path.join(S.GOLDEN, f"phold_t_{tag}_state.npy"))
    eng = S.make(api, c["A"], P=c["P"], LAM=c["LAMBDA"], END=c["END"])
    eng.seed_phold(c["K"], c["rootdiv"], 1)
    gvt = eng.drain()
    np.testing.assert_array_equal(eng.get_state(), want)
    assert eng.stats()["committed_n"] == g["commits"] and eng.digest()["checksum"] == g["checksum"] and gvt == g["gvt"]


# ---- drain(t_end): pause points, returned GVT, resume -------------------------------------------
def test_drain_t_end_segments(api):
    A, K, P, LAM, END = 1000, 16, 0.1, 100.0, 300
    eng = S.make(api, A, P=P, LAM=LAM, END=END, window=20)
    eng.seed_phold(K, 1, 1)
    for t_end in (50, 150, E.T_INF):
        gvt = eng.drain(t_end)
        o, ost = O.run(A, K, P, LAM, END, drain_end=t_end)
        S.check_against_oracle(eng, o, ost, gvt)
    # the two pause points are also pinned by the reference itself (goldens with drain_end 50/150)
    for g in S.goldens()["phold_t"]:
        if g["tag"] == "small" and g["drain_end"] in (50, 150):
            o, _ = O.run(A, K, P, LAM, END, drain_end=g["drain_end"])
            assert (o["commits"], o["gvt"]) == (g["commits"], g["gvt"])


# ---- pause / rewind / resume, as test/phold.cxx:180-198 drives it -------------------------------
@pytest.mark.parametrize("window", [7, 0])
def test_rewind_thirds_like_reference_test(api, window):
    A, K, P, LAM, END = 1000, 2, 0.5, 100.0, 10000
    ranks = 2
    eng = S.make(api, A, P=P, LAM=LAM, END=END, sub=0, window=window, lp_n_model=A)
    eng.seed_phold(K, 0, ranks)
    t0, dt = 0, END // 3
    while True:
        t1 = t0 + dt
        eng.drain(t1, rewindable=True)
        eng.rewind(True)
        gvt = eng.drain(t1, rewindable=True)
        eng.rewind(False)
        if gvt == E.T_INF:
            break
        t0 = t1
    s, d = eng.stats(), eng.digest()
    assert s["committed_n"] == 361624          # stats are not rewound: 2 x 180 812 (Appendix A.7)
    assert d["checksum"] == 15341107330209340805
    o, ost = O.run(A, K, P, LAM, END, ranks=ranks, user_subtime=0, rootdiv=0)
    np.testing.assert_array_equal(eng.get_state(), ost)


def test_rewind_requires_rewindable_drain(api):
    eng = S.make(api, 64, P=0.5, LAM=10.0, END=100)
    eng.seed_phold(2, 1, 1)
    eng.drain(50)
    with pytest.raises(E.EngineError):
        eng.rewind(True)


# ---- host-buffer path: set_lp_state + root_events (what pdes::register_state/root_event feed) ---
def test_host_roots_and_state_equal_device_seeding(api):
    A, K, P, LAM, END = 1000, 16, 0.1, 100.0, 300
    o, ost = O.run(A, K, P, LAM, END)
    eng = S.make(api, A, P=P, LAM=LAM, END=END, window=30)
    eng.set_state(E.phold_initial_state_np(E.MODEL_PHOLD_TEST, 0, A))
    eng.root_events(*E.phold_roots(E.MODEL_PHOLD_TEST, A, K, 0, A))
    gvt = eng.drain()
    S.check_against_oracle(eng, o, ost, gvt)
    np.testing.assert_array_equal(E.phold_initial_state_np(E.MODEL_PHOLD_TEST, 5, 40), E.phold_initial_state(E.MODEL_PHOLD_TEST, 5, 40))


def test_roots_into_an_empty_calendar_choose_the_bucket_width(api, engine_kind):
    """Tile engine, adaptive window: the first roots of an empty calendar pick the bucket width from their density
    (as device-side seeding does), a reset makes the calendar empty again, later roots keep the width; sparse roots keep
    the default; the results are the oracle's either way."""
    if engine_kind != "tile":
        pytest.skip("the LP engine has no calendar")
    A, K, P, LAM, END = 2048, 16, 0.1, 100.0, 200
    o, ost = O.run(A, K, P, LAM, END)
    eng = S.make(api, A, P=P, LAM=LAM, END=END, window=0)
    w0 = eng.stats()["window_last"]
    cols = E.phold_roots(E.MODEL_PHOLD_TEST, A, K, 0, A)
    for rep in range(2):
        eng.reset()
        eng.set_state(E.phold_initial_state_np(E.MODEL_PHOLD_TEST, 0, A))
        half = len(cols[0]) // 2
        eng.root_events(*[c[:half] for c in cols])            # K roots per LP at times 0 .. K-1: one per LP and tick
        w1 = eng.stats()["window_last"]
        assert w1 < w0 and w1 <= 8, (w0, w1)                  # (half the roots: one per LP and two ticks -> 8; all of them -> 4)
        eng.root_events(*[c[half:] for c in cols])
        assert eng.stats()["window_last"] == w1
        gvt = eng.drain()
        S.check_against_oracle(eng, o, ost, gvt, rewound_factor=rep + 1)      # (the counters are cumulative since create)
    eng.reset()
    eng.root_events(np.array([5, 900], np.uint64), np.array([0, 1], np.uint64), np.array([0, 1], np.uint32), np.array([0, 1], np.uint32))
    assert eng.stats()["window_last"] >= w1                   # two roots over 2048 LPs: nothing to narrow for


def test_next_simulation_does_not_inherit_the_last_window(api, engine_kind):
    """A finished PHOLD run leaves the adaptive window 30 times wider than it started (the events thin out towards
    end_time).  The next simulation on the same engine (reset = finalize + init) must not open its dense start with
    that window - a hundred generations of events in one window used to end in a capacity refusal (found by the
    randomised host-roots test: A=193, K=5, p=0.5, lambda=5, roots at time = ray)."""
    if engine_kind != "tile":
        pytest.skip("tile engine's calendar")
    A, K, P, LAM, END = 193, 5, 0.5, 5.0, 264
    o, ost = O.run(A, K, P, LAM, END, rootdiv=0)
    cols = E.phold_roots(E.MODEL_PHOLD_TEST, A, K, 0, A, rootdiv=0)
    eng = S.make(api, A, P=P, LAM=LAM, END=END, window=0)
    for rep in range(3):
        eng.reset()
        eng.set_state(E.phold_initial_state_np(E.MODEL_PHOLD_TEST, 0, A))
        eng.root_events(*cols)
        assert eng.stats()["window_last"] <= 16
        S.check_against_oracle(eng, o, ost, eng.drain(), rewound_factor=rep + 1)
    for rep in range(2):                       # the same through device-side seeding
        eng.seed_phold(K, 0, 1)
        S.check_against_oracle(eng, o, ost, eng.drain(), rewound_factor=4 + rep)


# ---- edge cases --------------------------------------------------------------------------------
def test_empty_simulation(api):
    eng = S.make(api, 16, P=0.5, LAM=10.0, END=100)
    assert eng.drain() == E.T_INF                      # no events at all: GVT = ~0 (pdes.cxx:878-886)
    assert eng.stats()["committed_n"] == 0


def test_single_lp_and_ragged_sizes(api):
    for A in (1, 33, 129):
        S.run_phold_t(api, A, 3, 0.5, 10.0, 200, window=25)


def test_events_beyond_t_end_stay_pending(api):
    A, K, P, LAM, END = 64, 4, 0.5, 10.0, 400
    eng = S.make(api, A, P=P, LAM=LAM, END=END, window=10)
    eng.seed_phold(K, 1, 1)
    gvt = eng.drain(0)          # nothing below t_end = 0
    assert gvt == 0 and eng.stats()["committed_n"] == 0 and eng.digest()["pending"] == A * K


def test_capacity_overflow_fails_loudly(api, engine_kind, monkeypatch):
    if engine_kind == "tile":   # its capacities are per tile: a far list too small for the roots it must park
        monkeypatch.setenv("DEVA_B200_TILE_C", "16")
        monkeypatch.setenv("DEVA_B200_TILE_CF", "64")
    eng = S.make(api, 64, P=0.5, LAM=10.0, END=400, pq_cap=4, inbox_cap=2, log_cap=8)
    with pytest.raises(E.EngineError) as ei:
        eng.seed_phold(4, 1, 1)
        eng.drain()
    assert ei.value.code == -3   # DEVA_B200_ERR_CAPACITY: never a silent drop


def test_small_capacities_still_exact(api):
    """Inbox spill, working-set eviction and a full undo log are slow paths, not wrong paths."""
    S.run_phold_t(api, 64, 4, 0.5, 10.0, 400, window=40, pq_cap=64, inbox_cap=1, log_cap=6)


@pytest.mark.parametrize("knobs", [
    dict(DEVA_B200_TILE_C="32", DEVA_B200_TILE_CF="65536"),                                   # bucket lists overflow into the far list (rebin + slow path)
    dict(DEVA_B200_TILE_CL="4"),                                   # late lists overflow into the job-wide spill list
    dict(DEVA_B200_TILE_RCAP="160"),                               # a tile's window does not fit the workspace: LP sub-range sweeps
    dict(DEVA_B200_TILE_C="24", DEVA_B200_TILE_CL="3", DEVA_B200_TILE_RCAP="160", DEVA_B200_TILE_CF="65536"),
], ids=["near", "late", "rcap", "all"])
@pytest.mark.parametrize("window", [1, 16, 40, 0])
def test_tile_engine_capacity_fallbacks_are_exact(api, engine_kind, monkeypatch, knobs, window):
    """The tile engine's fall-backs - far list for a full bucket list, spill list for a full late list, LP
    sub-ranges for a window larger than shared memory - are slow paths, not wrong paths."""
    if engine_kind != "tile":
        pytest.skip("tile engine knobs")
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    S.run_phold_t(api, 300, 8, 0.5, 20.0, 300, window=window)
    S.run_phold_t(api, 64, 4, 0.5, 10.0, 400, window=window, drain_end=150)


def test_tile_engine_runs_dense_models_with_default_settings(api, engine_kind):
    """lambda ~ 1 with every child leaving its LP (the reference runs these; the first engine refuses them
    with its default start-up window): no knob needed."""
    if engine_kind != "tile":
        pytest.skip("the first engine needs DEVA_B200_WINDOW_START=1 here (next tests)")
    S.run_phold_t(api, 47, 7, 1.0, 1.0, 15, window=0)
    S.run_phold_t(api, 181, 8, 1.0, 5.0, 40, window=0)
    S.run_phold_t(api, 600, 16, 1.0, 1.0, 30, window=0)


# ---- bench/phold.cxx model (rng + Box-Muller destination), made reproducible by an end_time -----
@pytest.mark.parametrize("window", [1, 2000, 0])
def test_bench_model_matches_oracle(api, window):
    lp_per_rank, R, K, END = 128, 4, 2, 60000
    A = lp_per_rank * R
    o, ost = O.run(A, K, 0.0, 10000.0, END, model=O.MODEL_BENCH, ranks=R, user_subtime=0, rootdiv=0,
                   bench_lp_per_rank=lp_per_rank, bench_peer_stddev=2.0)
    eng = E.Engine(E.MODEL_PHOLD_BENCH, A, window=window, lam=10000.0, end_time=END, user_subtime=0,
                   bench_lp_per_rank=lp_per_rank, bench_peer_stddev=2.0, api=api)
    eng.seed_phold(K, 0, R)
    gvt = eng.drain()
    st, s = eng.get_state(), eng.stats()
    assert s["committed_n"] == o["commits"] and gvt == o["gvt"]
    np.testing.assert_array_equal(st, ost[:, :2])


def test_bench_model_stop_flag(api):
    """bench/phold.cxx:95-104: after the wall-clock cut-off events stop re-sending and the drain ends."""
    lp_per_rank, R, K = 64, 2, 2
    A = lp_per_rank * R
    eng = E.Engine(E.MODEL_PHOLD_BENCH, A, window=5000, lam=10000.0, user_subtime=0,
                   bench_lp_per_rank=lp_per_rank, bench_peer_stddev=2.0, api=api)
    eng.seed_phold(K, 0, R)
    eng.drain(200000)
    before = eng.stats()["committed_n"]
    assert before > A * K
    eng.set_stop(True)
    assert eng.drain() == E.T_INF
    s = eng.stats()
    assert s["committed_n"] - before == eng_pending_before(A, K)  # every pending event runs once, sends nothing


def eng_pending_before(A, K):
    return A * K   # PHOLD keeps the live-event population constant until the cut-off


# ---- bench model against fixtures from the reference's own handler code (PHOLD-B, build_ref.py) ----
@pytest.mark.parametrize("tag", ["b_small", "b_narrow", "b_wide"])
def test_bench_model_against_reference_state_fixture(api, tag):
    g = [x for x in S.goldens()["phold_b"] if x["tag"] == tag][0]
    want = np.load(os.path.join(S.GOLDEN, f"phold_b_{tag}_state.npy"))
    A = g["lp_per_rank"] * g["ranks"]
    eng = E.Engine(E.MODEL_PHOLD_BENCH, A, window=0, lam=10000.0, end_time=g["end_time"], user_subtime=0,
                   bench_lp_per_rank=g["lp_per_rank"], bench_peer_stddev=g["peer_stddev"], api=api)
    eng.seed_phold(g["ray_per_lp"], 0, g["ranks"])
    gvt = eng.drain()
    np.testing.assert_array_equal(eng.get_state(), want)
    s = eng.stats()
    assert (s["committed_n"], gvt, s["deterministic"]) == (g["commits"], g["gvt"], 1)


# ---- DEVA_B200_WINDOW_START: reference-like start-up (window 1, doubling) for dense models -----------
def test_window_start_knob_keeps_dense_models_within_capacity(api, monkeypatch):
    """lambda = 1 with every child leaving its LP: the default 16-tick start-up window spans ~16 generations of
    events and overflows the pools (a loud refusal, see the next test); started at 1 the run completes and
    equals the oracle."""
    monkeypatch.setenv("DEVA_B200_WAKE_MIN", "0")
    monkeypatch.setenv("DEVA_B200_WINDOW_START", "1")
    S.run_phold_t(api, 47, 7, 1.0, 1.0, 15, window=0)
    S.run_phold_t(api, 181, 8, 1.0, 5.0, 40, window=0)
    S.run_phold_t(api, 1000, 16, 0.5, 100.0, 300, window=0)   # and an ordinary one


def test_default_start_window_refuses_loudly_when_far_too_optimistic(emu_api, monkeypatch):
    monkeypatch.setenv("DEVA_B200_ENGINE", "lp")   # (the tile engine runs this model with default settings, see above)
    monkeypatch.setenv("DEVA_B200_WAKE_MIN", "0")
    with pytest.raises(E.EngineError) as ei:
        S.run_phold_t(emu_api, 47, 7, 1.0, 1.0, 15, window=0)
    assert ei.value.code == -3   # DEVA_B200_ERR_CAPACITY: never a silent drop


def test_drain_timer_csv_rows_add_up(api, engine_kind, tmp_path, monkeypatch):
    """DEVA_B200_DRAIN_TIMER=<prefix> (the counterpart of the reference's DrainTimer, pdes.hxx:139-308): one CSV
    row per batch of launches / per epoch; the rows' executed / committed columns add up to the statistics."""
    if engine_kind != "tile":
        pytest.skip("the first engine's per-epoch CSV has other columns (profiles/r1_k_drain_timer.0.csv)")
    prefix = str(tmp_path / "timer")
    monkeypatch.setenv("DEVA_B200_DRAIN_TIMER", prefix)
    monkeypatch.setenv("DEVA_B200_TILE_BATCH", "3")
    eng = S.make(api, 1000, P=0.5, LAM=100.0, END=300, sub=1)
    eng.seed_phold(16, 1, 1)
    eng.drain()
    st = eng.stats()
    eng.close()
    import csv
    with open(prefix + ".0.csv") as f:
        rows = list(csv.DictReader(f))
    assert len(rows) >= 2
    assert sum(int(r["executed"]) for r in rows) == st["executed_n"]
    if "committed" in rows[0]:
        assert sum(int(r["committed"]) for r in rows) == st["committed_n"]
    times = [int(r["sim_time"]) for r in rows]
    assert times == sorted(times)


def test_heartbeat_callback(api, engine_kind):
    """deva_b200_set_heartbeat: the counterpart of pdes::drain()'s status block (pdes.cxx:282-301)"""
    if engine_kind != "tile":
        pytest.skip("the first engine has no heartbeat")
    beats = []
    eng = S.make(api, 4096, P=0.5, LAM=100.0, END=400, sub=1)
    eng.set_heartbeat(1e-9, lambda gvt, window, cps, eff: beats.append((gvt, window, cps, eff)))   # (every batch)
    eng.seed_phold(16, 1, 1)
    eng.drain()
    assert len(beats) >= 2
    assert all(0 < b[3] <= 1.0 for b in beats) and all(b[1] >= 1 for b in beats)
    assert [b[0] for b in beats] == sorted(b[0] for b in beats)
    n = len(beats)
    eng.set_heartbeat(0, None)
    eng.seed_phold(16, 1, 1)
    eng.drain()
    assert len(beats) == n
    eng.close()
