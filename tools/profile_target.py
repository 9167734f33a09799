#!/usr/bin/env python3
"""Small fixed target for ncu: one PHOLD-T drain at BASELINE config-2 geometry, fixed window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devastator_b200 import engine as E
A = int(os.environ.get("PT_A", 1 << 20)); W = int(os.environ.get("PT_W", 32)); END = int(os.environ.get("PT_END", 200))
eng = E.Engine(E.MODEL_PHOLD_TEST, A, p_remote=0.10, lam=100.0, end_time=END, window=W)
eng.seed_phold(16, 1, 1)
eng.drain()
s = eng.stats()
print({k: s[k] for k in ("committed_n", "executed_n", "passes", "exec_kernel_ms", "drain_ms")})
