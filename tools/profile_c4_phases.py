"""Where a C4 iteration goes on one rank: CUDA-event pairs around the phases of the fused iteration with the distributed
(external) linear solver. Usage: python tools/profile_c4_phases.py [scale]"""
import os
import sys
import time
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.problems import config_c4  # noqa: E402
from madipm_jl_b200.solver import MPCSolver  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
qp = config_c4(scale=scale)
s = MPCSolver(qp, kkt_system="Normal", linear_solver="distributed", n_border=qp.meta["n_border"])
s.solve()
s.k = 0
s.trace = []
pairs = defaultdict(list)


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        pairs[label if not callable(label) else label(*a)].append((e0, e1))
        return out
    setattr(obj, name, w)


wrap(s.h, "mpc_ext_begin", "ext_begin (termination + diagonals + assembly)")
wrap(s.h, "mpc_ext_fetch", "ext_fetch")
wrap(s.h, "mpc_ext_phase", lambda ph, *a: "ext_phase %d" % ph)
wrap(s.linear_solver, "factorize", "ls.factorize")
wrap(s.linear_solver, "solve", "ls.solve")
wrap(s.linear_solver, "is_factorized", "ls.is_factorized")
torch.cuda.synchronize()
t0 = time.perf_counter()
r = s.solve()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("solve %.3f s, %d iterations, %.2f ms/iteration" % (dt, r.iter, 1e3 * dt / r.iter))
tot = 0.0
for k, v in sorted(pairs.items()):
    ms = sum(a.elapsed_time(b) for a, b in v)
    tot += ms
    print("  %-50s calls %3d  total %8.2f ms  per call %7.3f ms" % (k, len(v), ms, ms / len(v)))
print("  sum of phases %.1f ms of %.1f" % (tot, 1e3 * dt))
print("  initialize %.1f ms" % (1e3 * r.counters.get("initialize_time", 0.0)))
# ---- per-class profile of the local factorization (stage -1 on the distributed solver's own handle)
ls = s.linear_solver
ls.h.gather(ls.nnz_loc, ls.nzval, ls.d_nzmap, ls.nz_loc)
os.environ["MIPM_PHASE_LOG"] = os.environ.get("C4_LEVELS", "gpurun_out/c4_levels.csv")
os.environ["MIPM_TASK_TRACE"] = os.environ.get("C4_TRACE", "gpurun_out/c4_trace.csv")
pf = ls.h.ls_factorize_profile(ls.nz_loc)
print({k_: round(v["ms"], 3) for k_, v in pf.items()}, "update busy-rate TF/s %.1f" % (pf["update"]["work"] / pf["update"]["ms"] / 1e9))
