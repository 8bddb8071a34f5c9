import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.problems import degenerate_lp
from madipm_jl_b200.solver import MPCSolver
for args in [(300, 1200, 5, 12, 0, 1e6), (300, 1200, 5, 15, 0, 1e8), (2000, 8000, 5, 16, 0, 1e7)]:
    qp = degenerate_lp(*args)
    for kw in (dict(fused=True), dict(fused=False), dict(fused=True, cudss_algorithm="LDL"), dict(fused=False, cudss_algorithm="LDL"), dict(fused=False, cudss_algorithm="LDL", max_refine=0),
               dict(fused=True, cudss_algorithm="LDL", ir_steps=1), dict(fused=True, cudss_algorithm="LDL", ir_steps=2)):
        s = MPCSolver(qp, kkt_system="Normal", max_iter=100, **kw)
        r = s.solve()
        t = r.trace[-1]
        print(qp.name, kw, r.status, r.iter, "%.9e" % r.objective, "fact", r.counters.get("factorizations"), "refine", r.counters.get("refinements"), "rejected", r.counters.get("refinements_rejected"),
              "pr %.1e du %.1e cp %.1e" % (t["inf_pr"], t["inf_du"], t["inf_compl"]), flush=True)
