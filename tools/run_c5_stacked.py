"""BASELINE config C5 through the stacked batch path: N units on one GPU, timing of construction and solve."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.batch import BatchedMPCSolver  # noqa: E402
from madipm_jl_b200.problems import config_c5  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kkt = sys.argv[2] if len(sys.argv) > 2 else "Normal"
models = [config_c5(i) for i in range(n)]
t0 = time.perf_counter()
b = BatchedMPCSolver(models, kkt_system=kkt)
torch.cuda.synchronize()
t1 = time.perf_counter()
res = b.solve()
t2 = time.perf_counter()
res = b.solve()
t3 = time.perf_counter()
ok = sum(r.status == "SOLVE_SUCCEEDED" for r in res)
print("units %d kkt %s: construct %.3f s, first solve %.3f s, warm solve %.3f s -> %.0f LPs/s (warm), %d succeeded, batch iterations %d, stats %s"
      % (n, kkt, t1 - t0, t2 - t1, t3 - t2, n / (t3 - t2), ok, res[0].counters["iterations_of_the_batch"],
         {k_: res[0].counters["ls_stats"][k_] for k_ in ("n", "nnz_l", "flops", "n_supernodes", "n_levels", "max_front_cols")}))
