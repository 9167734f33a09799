"""ctypes binding of oracle/_build/libphold_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(ROOT, "oracle", "_build", "libphold_oracle.so")

MODEL_TEST, MODEL_BENCH = 0, 1


class Cfg(C.Structure):
    _fields_ = [("model", C.c_int32), ("ranks", C.c_int32), ("lp_n", C.c_uint64), ("k", C.c_uint32),
                ("user_subtime", C.c_int32), ("rootdiv", C.c_int32), ("_pad", C.c_int32),
                ("p_remote", C.c_double), ("lam", C.c_double), ("end_time", C.c_uint64),
                ("drain_end", C.c_uint64), ("bench_lp_per_rank", C.c_int32), ("_pad2", C.c_int32),
                ("bench_peer_stddev", C.c_double)]


class Out(C.Structure):
    _fields_ = [("commits", C.c_uint64), ("checksum", C.c_uint64), ("gvt", C.c_uint64),
                ("pending", C.c_uint64), ("deterministic", C.c_int32), ("_pad", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "phold_oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB)
        _lib.phold_oracle_run.argtypes = [C.POINTER(Cfg), C.POINTER(Out), C.c_void_p]
        _lib.phold_oracle_run.restype = C.c_int
        _lib.phold_oracle_dt.argtypes = [C.c_uint64, C.c_double]
        _lib.phold_oracle_dt.restype = C.c_uint64
    return _lib


def run(lp_n, k, p_remote, lam, end_time, *, model=MODEL_TEST, ranks=1, user_subtime=1, rootdiv=1,
        drain_end=2**64 - 1, bench_lp_per_rank=0, bench_peer_stddev=2.0, want_state=True):
    """Returns (dict(commits, checksum, gvt, pending, deterministic), state[lp_n,3] u64 or None)."""
    cfg = Cfg(model, ranks, lp_n, k, user_subtime, rootdiv, 0, p_remote, lam, end_time, drain_end,
              bench_lp_per_rank, 0, bench_peer_stddev)
    out = Out()
    state = np.zeros((lp_n, 3), dtype=np.uint64) if want_state else None
    rc = lib().phold_oracle_run(C.byref(cfg), C.byref(out), state.ctypes.data if want_state else None)
    assert rc == 0
    return dict(commits=out.commits, checksum=out.checksum, gvt=out.gvt, pending=out.pending,
                deterministic=int(out.deterministic)), state


def dt(r, lam):
    return lib().phold_oracle_dt(r, lam)
