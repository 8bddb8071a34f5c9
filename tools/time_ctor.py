"""Stage times of MPCSolver(qp) on C2 in a fresh process (first constructor of the process, like bench.py's `cold`).
Usage: python tools/time_ctor.py   (MIPM_ANALYZE_LOG=1 adds the stages of the host analysis on stderr)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madipm_jl_b200 import problems
from madipm_jl_b200.solver import MPCSolver
torch.cuda.set_device(0)
qp = problems.config_c2()
t = time.time()
s = MPCSolver(qp, kkt_system="Normal")
torch.cuda.synchronize()
print("constructor %.3f s" % (time.time() - t))
prev = 0.0
for name, tt in s.setup_log:
    print("  %-42s %.3f (+%.3f)" % (name, tt, tt - prev))
    prev = tt
