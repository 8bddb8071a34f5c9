import sys, json, time
sys.path.insert(0,'.')
import torch
from madipm_jl_b200.problems import config_c2
from madipm_jl_b200.solver import MPCSolver
qp = config_c2(seed=2)
base=None
for tol in (1e-11, 1e-10, 1e-9, 1e-8):
    s = MPCSolver(qp, kkt_system="Normal", refine_tol=tol)
    s.solve()  # warm
    t=time.perf_counter(); r = s.solve(); torch.cuda.synchronize(); dt=time.perf_counter()-t
    tr = [(x['objective'], x['inf_pr'], x['inf_du'], x['inf_compl']) for x in r.trace]
    if base is None: base = tr
    dev = max(abs(a-b)/max(1,abs(a),abs(b)) for ta,tb in zip(tr,base) for a,b in zip(ta,tb))
    print(tol, r.status, r.iter, "solve_s %.3f"%dt, "refinements", r.counters.get("refinements",0), "max trace dev vs 1e-11: %.2e"%dev, flush=True)
