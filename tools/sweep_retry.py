"""Which ill-conditioned LPs drive the regularization retry and still converge (choice of the instance for
tests/test_gpu_parity.py::test_regularization_retry_end_to_end)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.problems import degenerate_lp
from madipm_jl_b200.solver import madipm
for cond in (1e6, 3e6, 1e7, 3e7):
    for seed in range(14, 22):
        qp = degenerate_lp(2000, 8000, 5, seed, cond=cond)
        out = []
        for opts in (dict(), dict(cudss_algorithm="LDL")):
            r = madipm(qp, kkt_system="Normal", max_iter=100, **opts)
            out.append("%s it %d fact %d obj %.8e" % (r.status, r.iter, r.counters["factorizations"], r.objective))
        print("cond %.0e seed %d | chol: %s | ldl: %s" % (cond, seed, out[0], out[1]), flush=True)
