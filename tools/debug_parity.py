import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from madipm_jl_b200.problems import bound_constrained_qp, mixed_bounds_lp
from madipm_jl_b200.solver import MPCSolver
from oracle.mpc_oracle import MPCOracle
def cmp(g, o, names):
    for nm in names:
        a = getattr(g, nm).cpu().numpy(); b = getattr(o, nm)
        fin = np.isfinite(b)
        d = np.abs(a[fin] - b[fin]).max() if fin.any() else 0.0
        bad = not np.array_equal(a[~fin], b[~fin])
        print("   %-8s maxdiff %.3e %s" % (nm, d, "INF-MISMATCH" if bad else ""))
for name, qp, kkt in (("boxqp", bound_constrained_qp(20, 4), "K2"), ("mixed30", mixed_bounds_lp(30, 90, 4, 1), "Normal")):
    g = MPCSolver(qp, kkt_system=kkt, fused=False); o = MPCOracle(qp, kkt_system=kkt)
    import time; g.start_time = o.start_time = time.time()
    o.initialize(); g.initialize()
    print(name, "after initialize: obj", g.obj_val, o.obj_val, "norm_b", g.norm_b, o.norm_b, "norm_c", g.norm_c, o.norm_c)
    cmp(g, o, ("x", "xl", "xu", "y", "zl", "zu", "f", "c", "jacl", "rhs"))
    for it in range(2):
        o.mpc_iteration(); g.mpc_iteration()
        print(name, "after iteration", it + 1, "obj", g.obj_val, o.obj_val, "alpha", g.alpha_p, o.alpha_p, g.alpha_d, o.alpha_d, "mu", g.mu, o.mu)
        cmp(g, o, ("x", "y", "zl", "zu", "d", "p", "pr_diag", "xl", "xu"))
