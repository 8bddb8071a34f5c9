"""Trace of the regularization-retry LP (tests/test_gpu_parity.py::test_regularization_retry_end_to_end)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.problems import degenerate_lp
from madipm_jl_b200.solver import MPCSolver
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cond = float(sys.argv[2]) if len(sys.argv) > 2 else 1e7
qp = degenerate_lp(2000, 8000, 5, seed, cond=cond)
for opts in (dict(), dict(cudss_algorithm="LDL"), dict(fused=False)):
    s = MPCSolver(qp, kkt_system="Normal", max_iter=100, **opts)
    orig = s.linear_solver.is_factorized
    log = []
    def isf(orig=orig, log=log, s=s):
        v = orig(); log.append((s.k, v, s.del_w)); return v
    s.linear_solver.is_factorized = isf
    r = s.solve()
    print(opts, r.status, r.iter, r.objective, r.counters)
    print(" failed factorizations", [x for x in log if not x[1]][:40])
    for t in r.trace[-45:]:
        print("  k %2d obj %.10e inf_pr %.3e inf_du %.3e compl %.3e mu %.3e" % (t["k"], t["objective"], t["inf_pr"], t["inf_du"], t["inf_compl"], t["mu"]))
