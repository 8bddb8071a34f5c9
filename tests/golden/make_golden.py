"""Regenerates the committed golden fixtures.

Two kinds of golden data live here:
  * reference-pinned: `simple_lp` (klamike/MadIPM.jl test/runtests.jl:29-60) must reach objective 1.0
    with SOLVE_SUCCEEDED, and NormalKKTSystem must agree with K2 at 1e-6 (:182-197). Those numbers
    come from the reference's own tests and are hard-coded in tests/test_oracle.py.
  * oracle-generated: per-iterate traces of the CPU oracle on small seeded instances, written to
    traces.json by this script (run from the repo root: python tests/golden/make_golden.py).
    The Julia reference cannot run in this image (no Julia), so these pin oracle == CUDA path
    and guard the oracle against regressions; they are NOT outputs of the Julia reference.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from madipm_jl_b200.problems import (badly_scaled_lp, bound_constrained_qp, mixed_bounds_lp, random_sparse_lp,  # noqa: E402
                                     random_sparse_qp, simple_lp)
from oracle.mpc_oracle import madipm  # noqa: E402

CASES = {
    "simple_lp": lambda: simple_lp(),
    "lp_m40_ub": lambda: random_sparse_lp(40, 160, 5, 7, structure="uniform", ub_fraction=0.5),
    "lp_m300_window": lambda: random_sparse_lp(300, 1500, 5, 7, structure="window", window=20),
    "qp_m60_window": lambda: random_sparse_qp(60, 200, 4, 9, structure="window", window=10),
    # every bound / constraint kind: range, one-sided and equality rows (slacks), free / boxed / upper-only variables
    "mixed_lp_m30": lambda: mixed_bounds_lp(30, 90, 4, 1),
    "mixed_lp_m120": lambda: mixed_bounds_lp(120, 400, 5, 2),
    # |A_ij| > 100 on a third of the rows, inequality rows included: MadNLP's con_scale path (scaled slacks, y0, rhs)
    "badscale_lp_m30": lambda: badly_scaled_lp(30, 90, 4, 1),
    # bounds only, m = 0 (DenseDummyQP(x0; m=0) in test/runtests.jl:64)
    "boxqp_n20": lambda: bound_constrained_qp(20, 4),
}


def main():
    out = {}
    for name, make in CASES.items():
        qp = make()
        for kkt in ("K2", "Normal"):
            if kkt == "Normal" and qp.nnzh > 0:
                continue
            st = madipm(qp, kkt_system=kkt)
            out[f"{name}/{kkt}"] = dict(
                status=st.status, iter=st.iter, objective=st.objective, dual_objective=st.dual_objective,
                trace=[{k: float(v) for k, v in t.items()} for t in st.trace])
            print(name, kkt, st.status, st.iter, st.objective)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traces.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
