"""ctypes binding of the C ABI in include/deva_b200.h (libdeva_b200.so).

This is plumbing: every call goes straight to the CUDA engine.  There is no CPU fallback - if the
shared library is missing or no B200 is visible, `load()` / `Engine()` raise.

Names follow the reference's pdes surface (src/devastator/pdes.hxx:82-128):
    Engine(...)              ~ pdes::init
    Engine.root_events       ~ pdes::root_event
    Engine.set_state/get_state ~ pdes::register_state (gather / scatter of the registered words)
    Engine.drain             ~ pdes::drain
    Engine.rewind            ~ pdes::rewind
    Engine.stats             ~ pdes::local_stats
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libdeva_b200.so")

MODEL_PHOLD_TEST = 1
MODEL_PHOLD_BENCH = 2
ABI_VERSION = 1
T_INF = 2**64 - 1


class ModelParams(C.Structure):
    _fields_ = [("lp_n_model", C.c_uint64), ("p_remote_thresh", C.c_uint64), ("lam", C.c_double),
                ("end_time", C.c_uint64), ("user_subtime", C.c_int32), ("bench_lp_per_rank", C.c_int32),
                ("bench_peer_stddev", C.c_double)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("model", C.c_int32), ("device", C.c_int32), ("n_gpus", C.c_int32),
                ("gpu_rank", C.c_int32), ("pq_cap", C.c_int32), ("inbox_cap", C.c_int32), ("log_cap", C.c_int32),
                ("lp_n", C.c_uint64), ("lp_begin", C.c_uint64), ("lp_count", C.c_uint64), ("window", C.c_uint64),
                ("mp", ModelParams)]


class Stats(C.Structure):
    _fields_ = [("executed_n", C.c_uint64), ("committed_n", C.c_uint64), ("deterministic", C.c_int32),
                ("_pad", C.c_int32), ("rolled_back_n", C.c_uint64), ("anti_sent_n", C.c_uint64),
                ("remote_sent_n", C.c_uint64), ("passes", C.c_uint64), ("window_last", C.c_uint64),
                ("drain_ms", C.c_double), ("exec_kernel_ms", C.c_double), ("kernel_launches", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "_pad"}


# every symbol include/deva_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "deva_b200_create", "deva_b200_destroy", "deva_b200_last_error", "deva_b200_abi_version",
    "deva_b200_state_words", "deva_b200_set_lp_state", "deva_b200_get_lp_state", "deva_b200_root_events",
    "deva_b200_seed_phold", "deva_b200_drain", "deva_b200_rewind", "deva_b200_get_stats", "deva_b200_set_stop",
    "deva_b200_digest", "deva_b200_comm_unique_id", "deva_b200_comm_init", "deva_b200_stream",
    "deva_b200_reset", "deva_b200_comm_init_group", "deva_b200_set_heartbeat",
    "deva_b200_hostq_create", "deva_b200_hostq_destroy", "deva_b200_hostq_push", "deva_b200_hostq_pop_min",
    "deva_b200_hostq_pop_min_key", "deva_b200_hostq_erase", "deva_b200_hostq_snapshot", "deva_b200_hostq_rewind",
]


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"deva_b200 error {code}: {msg}")
        self.code = code


HEARTBEAT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_double)


class Api:
    """Typed view of a loaded library exporting the deva_b200 C ABI."""

    def __init__(self, cdll):
        self.lib = cdll
        L = cdll
        P = C.c_void_p
        L.deva_b200_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
        L.deva_b200_destroy.argtypes = [P]
        L.deva_b200_destroy.restype = None
        L.deva_b200_last_error.restype = C.c_char_p
        L.deva_b200_state_words.argtypes = [P]
        L.deva_b200_set_lp_state.argtypes = [P, C.c_uint64, C.c_uint64, P]
        L.deva_b200_get_lp_state.argtypes = [P, C.c_uint64, C.c_uint64, P]
        L.deva_b200_root_events.argtypes = [P, C.c_uint64, P, P, P, P]
        L.deva_b200_seed_phold.argtypes = [P, C.c_uint32, C.c_int32, C.c_int32]
        L.deva_b200_drain.argtypes = [P, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64)]
        L.deva_b200_rewind.argtypes = [P, C.c_int32]
        L.deva_b200_get_stats.argtypes = [P, C.POINTER(Stats)]
        L.deva_b200_set_stop.argtypes = [P, C.c_int32]
        L.deva_b200_digest.argtypes = [P, P]
        L.deva_b200_comm_unique_id.argtypes = [P]
        L.deva_b200_comm_init.argtypes = [P, P]
        L.deva_b200_stream.argtypes = [P]
        L.deva_b200_reset.argtypes = [P]
        L.deva_b200_stream.restype = P
        L.deva_b200_comm_init_group.argtypes = [C.POINTER(P), C.c_int32]
        L.deva_b200_set_heartbeat.argtypes = [P, C.c_double, HEARTBEAT_FN, P]
        L.deva_b200_hostq_create.argtypes = [C.c_int32, C.c_uint64, C.POINTER(P)]
        L.deva_b200_hostq_destroy.argtypes = [P]
        L.deva_b200_hostq_destroy.restype = None
        L.deva_b200_hostq_push.argtypes = [P, C.c_uint64, P, P, P, P]
        L.deva_b200_hostq_pop_min.argtypes = [P, C.c_uint64, C.c_uint64, P, P, P, P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.deva_b200_hostq_pop_min_key.argtypes = L.deva_b200_hostq_pop_min.argtypes
        L.deva_b200_hostq_erase.argtypes = [P, C.c_uint64, P, C.POINTER(C.c_uint64)]
        L.deva_b200_hostq_snapshot.argtypes = [P]
        L.deva_b200_hostq_rewind.argtypes = [P, C.c_int32]

    def check(self, rc):
        if rc != 0:
            raise EngineError(rc, self.lib.deva_b200_last_error().decode())


_api = None


def load():
    """Loads libdeva_b200.so (built in-tree by __graft_entry__.build()).  No fallback."""
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU path.")
        _api = Api(C.CDLL(LIB_PATH))
        if _api.lib.deva_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("libdeva_b200.so ABI version mismatch")
    return _api


def remote_thresh(p_remote):
    """uint64_t(percent_remote*double(-1ull)) as g++ evaluates it in the reference (test/phold.cxx:119).
    percent_remote is a compile-time constant there, so the conversion is constant-folded: for
    p = 1.0 the product is 2^64, out of range, and GCC folds it to the saturated 2^64-1 (the
    run-time cvttsd2si sequence would give 0; checked with g++ 13.3 and pinned by the reference
    binary's golden `allrem`, tests/golden/ref_goldens.json)."""
    x = float(p_remote) * 18446744073709551616.0
    if x >= 18446744073709551616.0:
        return 2**64 - 1
    return int(x)


class Engine:
    """One engine = one GPU's block of LPs."""

    def __init__(self, model, lp_n, *, lp_begin=0, lp_count=None, device=0, n_gpus=1, gpu_rank=0, window=0,
                 pq_cap=0, inbox_cap=0, log_cap=0, lp_n_model=None, p_remote=0.0, lam=100.0, end_time=T_INF,
                 user_subtime=1, bench_lp_per_rank=0, bench_peer_stddev=2.0, api=None):
        self.api = api or load()
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.model = model
        cfg.device = device
        cfg.n_gpus = n_gpus
        cfg.gpu_rank = gpu_rank
        cfg.pq_cap, cfg.inbox_cap, cfg.log_cap = pq_cap, inbox_cap, log_cap
        cfg.lp_n = lp_n
        cfg.lp_begin = lp_begin
        cfg.lp_count = lp_n - lp_begin if lp_count is None else lp_count
        cfg.window = window
        cfg.mp.lp_n_model = lp_n if lp_n_model is None else lp_n_model
        cfg.mp.p_remote_thresh = remote_thresh(p_remote)
        cfg.mp.lam = lam
        cfg.mp.end_time = end_time
        cfg.mp.user_subtime = user_subtime
        cfg.mp.bench_lp_per_rank = bench_lp_per_rank
        cfg.mp.bench_peer_stddev = bench_peer_stddev
        self.cfg = cfg
        self.h = C.c_void_p()
        self.api.check(self.api.lib.deva_b200_create(C.byref(cfg), C.byref(self.h)))
        self.state_words = self.api.lib.deva_b200_state_words(self.h)

    def close(self):
        if self.h:
            self.api.lib.deva_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pdes::register_state: gather / scatter of the registered state words -----------------
    def set_state(self, words, lp_first=None):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        assert w.ndim == 2 and w.shape[1] == self.state_words
        first = self.cfg.lp_begin if lp_first is None else lp_first
        self.api.check(self.api.lib.deva_b200_set_lp_state(self.h, first, w.shape[0], w.ctypes.data))

    def get_state(self, lp_first=None, count=None, out=None):
        first = self.cfg.lp_begin if lp_first is None else lp_first
        n = self.cfg.lp_count if count is None else count
        w = out if out is not None else np.empty((n, self.state_words), dtype=np.uint64)
        self.api.check(self.api.lib.deva_b200_get_lp_state(self.h, first, n, w.ctypes.data))
        return w

    # ---- pdes::root_event ---------------------------------------------------------------------
    def root_events(self, time, subtime, lp, payload):
        t = np.ascontiguousarray(time, dtype=np.uint64)
        s = np.ascontiguousarray(subtime, dtype=np.uint64)
        l = np.ascontiguousarray(lp, dtype=np.uint32)
        p = np.ascontiguousarray(payload, dtype=np.uint32)
        assert t.shape == s.shape == l.shape == p.shape
        self.api.check(self.api.lib.deva_b200_root_events(self.h, t.size, t.ctypes.data, s.ctypes.data,
                                                          l.ctypes.data, p.ctypes.data))

    def reset(self):
        self.api.check(self.api.lib.deva_b200_reset(self.h))

    def seed_phold(self, k_per_lp, rootdiv=1, ranks=1):
        self.api.check(self.api.lib.deva_b200_seed_phold(self.h, k_per_lp, rootdiv, ranks))

    # ---- pdes::drain / rewind / local_stats ---------------------------------------------------
    def drain(self, t_end=T_INF, rewindable=False):
        gvt = C.c_uint64()
        self.api.check(self.api.lib.deva_b200_drain(self.h, t_end, int(rewindable), C.byref(gvt)))
        return gvt.value

    def rewind(self, do_rewind):
        self.api.check(self.api.lib.deva_b200_rewind(self.h, int(do_rewind)))

    def stats(self):
        s = Stats()
        self.api.check(self.api.lib.deva_b200_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def set_heartbeat(self, secs, fn):
        """fn(gvt, window, commits_per_sec, efficiency) about every `secs` seconds of a drain (pdes::chitter_secs)"""
        self._hb = HEARTBEAT_FN(lambda user, gvt, window, cps, eff: fn(gvt, window, cps, eff)) if fn else C.cast(None, HEARTBEAT_FN)
        self.api.check(self.api.lib.deva_b200_set_heartbeat(self.h, float(secs), self._hb, None))

    def set_stop(self, stop):
        self.api.check(self.api.lib.deva_b200_set_stop(self.h, int(stop)))

    def digest(self):
        out = (C.c_uint64 * 4)()
        self.api.check(self.api.lib.deva_b200_digest(self.h, out))
        return dict(checksum=out[0], state_hash=out[1], pending=out[2], pending_hash=out[3])

    def stream_ptr(self):
        """cudaStream_t the engine launches on (wrap with torch.cuda.ExternalStream to time it)."""
        return self.api.lib.deva_b200_stream(self.h)

    def comm_init(self, uid_bytes):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self.api.check(self.api.lib.deva_b200_comm_init(self.h, buf))


class HostQueue:
    """The device-resident pending-event pool behind host-executed handlers (deva_b200_hostq_*): push events, pop all
    events at the commit horizon (smallest pending time) ordered by (lp, subtime, handle)."""

    def __init__(self, capacity=0, device=0, api=None):
        self.api = api or load()
        self.h = C.c_void_p()
        self.api.check(self.api.lib.deva_b200_hostq_create(device, capacity, C.byref(self.h)))

    def push(self, time, subtime, lp, handle):
        t = np.ascontiguousarray(time, dtype=np.uint64); s = np.ascontiguousarray(subtime, dtype=np.uint64)
        l = np.ascontiguousarray(lp, dtype=np.uint32); h = np.ascontiguousarray(handle, dtype=np.uint64)
        self.api.check(self.api.lib.deva_b200_hostq_push(self.h, t.size, t.ctypes.data, s.ctypes.data, l.ctypes.data, h.ctypes.data))

    def pop_min(self, t_end=T_INF, room=1 << 16, by_key=False):
        """all events at the smallest time stamp; by_key: only those at the smallest (time, subtime) pair"""
        t = np.empty(room, np.uint64); s = np.empty(room, np.uint64); l = np.empty(room, np.uint32); h = np.empty(room, np.uint64)
        n, gvt, pending = C.c_uint64(), C.c_uint64(), C.c_uint64()
        fn = self.api.lib.deva_b200_hostq_pop_min_key if by_key else self.api.lib.deva_b200_hostq_pop_min
        self.api.check(fn(self.h, t_end, room, t.ctypes.data, s.ctypes.data, l.ctypes.data, h.ctypes.data, C.byref(n), C.byref(gvt), C.byref(pending)))
        k = n.value
        return dict(time=t[:k], subtime=s[:k], lp=l[:k], handle=h[:k], gvt=gvt.value, pending=pending.value)

    def erase(self, handles):
        """removes the pending events with these handles; returns how many were found"""
        h = np.ascontiguousarray(handles, dtype=np.uint64)
        n = C.c_uint64()
        self.api.check(self.api.lib.deva_b200_hostq_erase(self.h, h.size, h.ctypes.data, C.byref(n)))
        return n.value

    def snapshot(self):
        self.api.check(self.api.lib.deva_b200_hostq_snapshot(self.h))

    def rewind(self, do_rewind):
        self.api.check(self.api.lib.deva_b200_hostq_rewind(self.h, int(do_rewind)))

    def close(self):
        if self.h:
            self.api.lib.deva_b200_hostq_destroy(self.h)
            self.h = C.c_void_p()


def comm_init_group(engines):
    """engines that share this process (one per GPU, gpu_rank = index): peer access instead of IPC + NCCL"""
    api = engines[0].api
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    api.check(api.lib.deva_b200_comm_init_group(arr, len(engines)))


def comm_unique_id(api=None):
    api = api or load()
    buf = C.create_string_buffer(128)
    api.check(api.lib.deva_b200_comm_unique_id(buf))
    return buf.raw


# ---- host-side construction of the PHOLD instances exactly as the reference apps do -------------
def phold_initial_state(model, lp_first, count):
    """state_cur[cd] = rng_state{lp}; check[cd] = lp  (test/phold.cxx:19-28,161-162; bench/phold.cxx:150)."""
    M = (1 << 64) - 1
    out = np.empty((count, 3 if model == MODEL_PHOLD_TEST else 2), dtype=np.uint64)
    for i in range(count):
        seed = lp_first + i
        a = (0x1234567812345678 * (1 + seed)) & M
        b = (0xdeadbeefdeadbeef * (10 + seed)) & M
        for _ in range(2):
            x, y = a, b
            a = y
            x ^= (x << 23) & M
            b = x ^ y ^ (x >> 17) ^ (y >> 26)
        out[i, 0], out[i, 1] = a, b
        if model == MODEL_PHOLD_TEST:
            out[i, 2] = seed
    return out


def phold_initial_state_np(model, lp_first, count):
    """Vectorised version of phold_initial_state (numpy u64 arithmetic wraps mod 2^64)."""
    seed = np.arange(lp_first, lp_first + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        a = np.uint64(0x1234567812345678) * (np.uint64(1) + seed)
        b = np.uint64(0xdeadbeefdeadbeef) * (np.uint64(10) + seed)
        for _ in range(2):
            x, y = a, b
            a = y
            x = x ^ (x << np.uint64(23))
            b = x ^ y ^ (x >> np.uint64(17)) ^ (y >> np.uint64(26))
    cols = [a, b] + ([seed] if model == MODEL_PHOLD_TEST else [])
    return np.ascontiguousarray(np.stack(cols, axis=1))


def phold_roots(model, lp_n_model, k, lp_first, lp_count, *, rootdiv=1, user_subtime=1, ranks=1):
    """Root events of the LPs [lp_first, lp_first+lp_count) as the reference apps create them
    (test/phold.cxx:172-178 / PHOLD-T root time; bench/phold.cxx:152-158).  Returns SoA columns."""
    lp = np.repeat(np.arange(lp_first, lp_first + lp_count, dtype=np.uint64), k)
    j = np.tile(np.arange(k, dtype=np.uint64), lp_count)
    if model == MODEL_PHOLD_BENCH:
        ray = lp * np.uint64(k) + j
        time = ray.copy()
    else:
        ray = lp + j * np.uint64(lp_n_model)
        time = j.copy() if rootdiv else ray.copy()
    per = (lp_n_model + ranks - 1) // ranks
    sub = ray.copy() if user_subtime else lp % np.uint64(per)
    return time, sub, lp.astype(np.uint32), ray.astype(np.uint32)
