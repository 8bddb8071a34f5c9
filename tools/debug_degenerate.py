import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.problems import degenerate_lp
from madipm_jl_b200.solver import MPCSolver
from oracle.mpc_oracle import madipm
qp = degenerate_lp(300, 1200, 5, 12, 0, 1e6)
o = madipm(qp, kkt_system="Normal")
for fused in (True, False):
    s = MPCSolver(qp, kkt_system="Normal", fused=fused, max_iter=24)
    orig = s.linear_solver.is_factorized
    log = []
    def isf(orig=orig, log=log, s=s):
        v = orig(); log.append((s.k, v, s.del_w)); return v
    s.linear_solver.is_factorized = isf
    r = s.solve()
    print("fused" if fused else "fine", r.status, r.iter, r.counters)
    print(" is_factorized log", log[:60])
    for t, to in zip(r.trace[8:], o.trace[8:]):
        print("  k %2d inf_pr %.3e inf_du %.3e compl %.3e mu %.3e | oracle %.3e %.3e %.3e %.3e" % (t["k"], t["inf_pr"], t["inf_du"], t["inf_compl"], t["mu"], to["inf_pr"], to["inf_du"], to["inf_compl"], to["mu"]))
    for t in r.trace[len(o.trace):]:
        print("  k %2d inf_pr %.3e inf_du %.3e compl %.3e mu %.3e" % (t["k"], t["inf_pr"], t["inf_du"], t["inf_compl"], t["mu"]))
