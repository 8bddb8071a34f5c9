import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madipm_jl_b200.problems import config_c2
from madipm_jl_b200.solver import MPCSolver
qp = config_c2(seed=2)
s = MPCSolver(qp, kkt_system="Normal")
s.solve()
s.k = 0; s.start_time = time.time()
pr = cProfile.Profile(); pr.enable()
s.initialize(); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
