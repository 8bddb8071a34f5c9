import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madipm_jl_b200.problems import config_c4
from madipm_jl_b200.solver import MPCSolver
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
mode = sys.argv[2] if len(sys.argv) > 2 else "distributed"
qp = config_c4(scale=scale)
kw = dict(linear_solver="distributed", n_border=qp.meta["n_border"]) if mode == "distributed" else {}
s = MPCSolver(qp, kkt_system="Normal", rethrow_error=False, **kw)
try:
    s.opt.rethrow_error = True
    r = s.solve()
    print("status", r.status)
except Exception as e:
    print("EXC", type(e).__name__, e)
for t in s.trace[-8:]:
    print({k: (round(v, 12) if isinstance(v, float) else v) for k, v in t.items() if k in ("k", "objective", "inf_pr", "inf_du", "inf_compl", "mu", "alpha_p", "alpha_d", "del_w")})
print("residual_ratio", getattr(s, "residual_ratio", None), "factorizations", s.cnt["factorizations"], "refinements", s.cnt.get("refinements"))
