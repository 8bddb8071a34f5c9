"""Exploration: degenerate / ill-conditioned / rank-deficient LPs through the CUDA path and the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madipm_jl_b200.problems import degenerate_lp
from madipm_jl_b200.solver import MPCSolver
from oracle.mpc_oracle import madipm
for (m, n, k, seed, dup, cond) in [(300, 1200, 5, 11, 0, 1.0), (300, 1200, 5, 12, 0, 1e6), (300, 1200, 5, 13, 8, 1.0), (2000, 8000, 5, 14, 0, 1e4)]:
    qp = degenerate_lp(m, n, k, seed, dup, cond)
    for kkt in ("Normal", "K2"):
        o = madipm(qp, kkt_system=kkt) if m <= 300 else None
        for fused in (True, False):
            try:
                s = MPCSolver(qp, kkt_system=kkt, fused=fused)
                r = s.solve()
                print(qp.name, kkt, "fused" if fused else "fine", r.status, r.iter, "%.10e" % r.objective, "target %.10e" % qp.meta["objective"],
                      "fact", r.counters.get("factorizations"), "refine", r.counters.get("refinements"),
                      ("oracle %s %d %.10e" % (o.status, o.iter, o.objective)) if o else "", flush=True)
            except Exception as e:
                print(qp.name, kkt, fused, "EXC", repr(e)[:200], ("oracle %s %d" % (o.status, o.iter)) if o else "", flush=True)
