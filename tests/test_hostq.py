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


@pytest.mark.parametrize("n,times", [(1, 1), (1000, 7), (50000, 300), (200000, 3)])
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
