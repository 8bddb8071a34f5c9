import sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import torch
from madipm_jl_b200 import problems
from madipm_jl_b200.solver import MPCSolver
qp = problems.config_c5(0)
s = MPCSolver(qp, kkt_system="Normal"); s.solve()
for i in range(1, 3):
    qp = problems.config_c5(i)
    t = time.time(); s = MPCSolver(qp, kkt_system="Normal"); t1 = time.time() - t
    t = time.time(); s.solve(); torch.cuda.synchronize(); t2 = time.time() - t
    print("ctor %.3f solve %.3f" % (t1, t2))
qp = problems.config_c5(5)
pr = cProfile.Profile(); pr.enable(); s = MPCSolver(qp, kkt_system="Normal"); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
