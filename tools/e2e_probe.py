import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from devastator_b200 import engine as E
A=1<<20; K=16
eng = E.Engine(E.MODEL_PHOLD_TEST, A, p_remote=0.1, lam=100.0, end_time=1000, window=0)
st = torch.from_numpy(E.phold_initial_state_np(E.MODEL_PHOLD_TEST,0,A)).pin_memory()
cols=[torch.from_numpy(np.ascontiguousarray(c)).pin_memory() for c in E.phold_roots(E.MODEL_PHOLD_TEST,A,K,0,A)]
out=torch.empty_like(st).pin_memory()
for rep in range(4):
    t=[time.perf_counter()]
    eng.reset(); torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.set_state(st.numpy()); t.append(time.perf_counter())
    eng.root_events(*[c.numpy() for c in cols]); t.append(time.perf_counter())
    s0=eng.stats(); g=eng.drain(); s1=eng.stats(); t.append(time.perf_counter())
    eng.get_state(out=out.numpy()); t.append(time.perf_counter())
    d=[round((b-a)*1e3,2) for a,b in zip(t,t[1:])]
    print("reset,set_state,roots,drain,get_state ms:", d, "passes", s1['passes']-s0['passes'], "exec", s1['executed_n']-s0['executed_n'], "commit", s1['committed_n']-s0['committed_n'], "Wlast", s1['window_last'], flush=True)
for rep in range(2):
    eng.seed_phold(K,1,1); s0=eng.stats(); t0=time.perf_counter(); eng.drain(); t1=time.perf_counter(); s1=eng.stats()
    print("seeded drain ms", round((t1-t0)*1e3,2), "passes", s1['passes']-s0['passes'], "exec", s1['executed_n']-s0['executed_n'])
