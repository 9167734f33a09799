"""The pending-event pool behind host-executed handlers (deva_b200_hostq_*): whatever is pushed comes back exactly once,
time stamp by time stamp (the commit horizon = smallest pending time), ordered by (lp, subtime, handle) within one.
CPU suite: the emulation's host stand-in (same contract).  GPU suite: the device pool - min-reduction, stream compaction,
merge sort (devastator_b200/csrc/hostq.cuh)."""
import numpy as np
import pytest

from devastator_b200 import engine as E

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def qapi(request):
    return request.getfixturevalue("emu_api" if request.param == "emu" else "cuda_api")


# (the device pool pops small pools - up to 4096 records, up to 1024 of them at the horizon - with one block: the cases
#  cover that path, the general one, the hand-over between them and a horizon too large for the block)
@pytest.mark.parametrize("n,times", [(1, 1), (1000, 7), (3000, 2), (4096, 4), (4097, 5), (50000, 300), (200000, 3)])
def test_everything_comes_back_in_order(qapi, n, times):
    rng = np.random.default_rng(n)
    t = rng.integers(5, 5 + times, n).astype(np.uint64); s = rng.integers(0, 50, n).astype(np.uint64)
    lp = rng.integers(0, 1000, n).astype(np.uint32); h = np.arange(n, dtype=np.uint64) * 8 + 4096
    q = E.HostQueue(capacity=1 << 20, api=qapi)
    half = n // 2
    q.push(t[:half], s[:half], lp[:half], h[:half])
    q.push(t[half:], s[half:], lp[half:], h[half:])
    seen, last_time = [], -1
    while True:
        b = q.pop_min(room=max(n, 1))
        if len(b["time"]) == 0:
            assert b["gvt"] == E.T_INF and b["pending"] == 0
            break
        tm = int(b["time"][0])
        assert tm > last_time and (b["time"] == tm).all() and b["gvt"] == tm     # one time stamp per pop, strictly increasing
        last_time = tm
        key = np.stack([b["lp"].astype(np.uint64), b["subtime"], b["handle"]], axis=1)
        assert (np.lexsort((key[:, 2], key[:, 1], key[:, 0])) == np.arange(len(key))).all()   # sorted by (lp, subtime, handle)
        seen.append(b["handle"])
    seen = np.concatenate(seen)
    assert len(seen) == n and np.array_equal(np.sort(seen), np.sort(h))
    q.close()


def test_horizon_respects_t_end_and_capacity_is_loud(qapi):
    q = E.HostQueue(capacity=8, api=qapi)
    q.push([10, 20, 20], [0, 0, 1], [1, 2, 3], [8, 16, 24])
    b = q.pop_min(t_end=10)
    assert len(b["time"]) == 0 and b["gvt"] == 10 and b["pending"] == 3     # nothing below t_end: GVT is returned, nothing removed
    b = q.pop_min(t_end=11)
    assert list(b["handle"]) == [8] and b["pending"] == 2
    with pytest.raises(E.EngineError) as ei:
        q.push(np.arange(10), np.arange(10), np.arange(10), np.arange(10))
    assert ei.value.code == -3
    with pytest.raises(E.EngineError):     # more events at one time stamp than the caller has room for: nothing is lost
        q.pop_min(room=1)
    b = q.pop_min(room=4)
    assert sorted(b["handle"]) == [16, 24]
    q.close()


def _drain_all(q, by_key=False, room=1 << 18):
    out = []
    while True:
        b = q.pop_min(room=room, by_key=by_key)
        if len(b["time"]) == 0:
            return out
        out.append(b)


@pytest.mark.parametrize("n", [1, 777, 60000])
def test_key_by_key_pops_follow_the_stamped_order(qapi, n):
    """deva_b200_hostq_pop_min_key: one (time, subtime) pair per pop, pairs strictly increasing (pdes.hxx:936-941)"""
    rng = np.random.default_rng(n + 1)
    t = rng.integers(3, 9, n).astype(np.uint64); s = rng.integers(0, 40, n).astype(np.uint64)
    s[rng.integers(0, n)] = np.uint64(2**64 - 1)                       # (the largest subtime is a legal one)
    lp = rng.integers(0, 500, n).astype(np.uint32); h = np.arange(n, dtype=np.uint64) * 16 + 64
    q = E.HostQueue(capacity=1 << 17, api=qapi)
    q.push(t, s, lp, h)
    last, seen = (-1, -1), []
    for b in _drain_all(q, by_key=True):
        key = (int(b["time"][0]), int(b["subtime"][0]))
        assert key > last and (b["time"] == key[0]).all() and (b["subtime"] == key[1]).all()
        assert (np.lexsort((b["handle"], b["lp"])) == np.arange(len(b["lp"]))).all()
        last = key
        seen.append(b["handle"])
    assert np.array_equal(np.sort(np.concatenate(seen)), h)
    q.close()


@pytest.mark.parametrize("n,m", [(10, 3), (5000, 1234), (100000, 99999)])
def test_erase_by_handle(qapi, n, m):
    """deva_b200_hostq_erase: exactly the named events leave the pool (unknown handles are ignored), the rest still
    comes back in order"""
    rng = np.random.default_rng(n)
    t = rng.integers(0, 50, n).astype(np.uint64); s = rng.integers(0, 9, n).astype(np.uint64)
    lp = rng.integers(0, 64, n).astype(np.uint32); h = rng.permutation(n).astype(np.uint64) * 8 + 800
    q = E.HostQueue(capacity=1 << 17, api=qapi)
    q.push(t, s, lp, h)
    gone = rng.choice(h, m, replace=False)
    assert q.erase(np.concatenate([gone, np.array([1, 2, 3], dtype=np.uint64)])) == m
    assert q.erase(gone) == 0
    left = np.concatenate([b["handle"] for b in _drain_all(q)]) if m < n else np.zeros(0, np.uint64)
    assert np.array_equal(np.sort(left), np.sort(np.setdiff1d(h, gone)))
    q.close()


def test_snapshot_and_rewind(qapi):
    """rewindable drains: the pool as it was at the snapshot comes back on rewind(true); rewind(false) keeps the present"""
    rng = np.random.default_rng(5)
    n = 30000
    t = rng.integers(0, 100, n).astype(np.uint64); s = rng.integers(0, 9, n).astype(np.uint64)
    lp = rng.integers(0, 64, n).astype(np.uint32); h = np.arange(n, dtype=np.uint64) + 1
    q = E.HostQueue(capacity=1 << 16, api=qapi)
    q.push(t, s, lp, h)
    q.snapshot()
    with pytest.raises(E.EngineError):
        q.snapshot()                                  # lingering snapshot
    first = [q.pop_min(t_end=40, room=n) for _ in range(60)]
    q.push([1000, 1001], [0, 0], [1, 2], [10**9, 10**9 + 1])
    q.rewind(True)
    again = [q.pop_min(t_end=40, room=n) for _ in range(60)]
    for a, b in zip(first, again):
        assert np.array_equal(a["handle"], b["handle"]) and a["gvt"] == b["gvt"] and a["pending"] == b["pending"]
    with pytest.raises(E.EngineError):
        q.rewind(False)                               # nothing to rewind
    q.snapshot()
    q.pop_min(room=n)
    q.rewind(False)
    rest = np.concatenate([b["handle"] for b in _drain_all(q)])
    assert len(rest) == int((t > 40).sum()) and 10**9 not in rest
    q.close()
