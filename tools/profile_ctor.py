"""Where MPCSolver(qp) construction time goes (host analysis + uploads). Usage: python tools/profile_ctor.py [c2|c5x128]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from madipm_jl_b200 import problems  # noqa: E402
from madipm_jl_b200.solver import MPCSolver  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
torch.zeros(1, device="cuda")
if which == "c2":
    qp = problems.config_c2()
else:
    from madipm_jl_b200.batch import stack_models
    qp = stack_models([problems.config_c5(i) for i in range(128)])[0]
MPCSolver(problems.config_c5(0), kkt_system="Normal")      # first-use costs (library load, context)
os.environ["MIPM_ANALYZE_LOG"] = "1"
t = time.time()
pr = cProfile.Profile()
pr.enable()
s = MPCSolver(qp, kkt_system="Normal")
torch.cuda.synchronize()
pr.disable()
print("constructor %.3f s (host cores: %d)" % (time.time() - t, os.cpu_count()))
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
