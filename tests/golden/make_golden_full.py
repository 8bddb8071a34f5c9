"""Full-size golden traces of the configs bench.py measures (BASELINE.json configs C2-C5).

Run from the repo root (CPU only, ~10 minutes):  python tests/golden/make_golden_full.py [case ...]

The CPU oracle (oracle/mpc_oracle.py, the restatement of src/solver.jl:6-360 + src/kernels.jl) is run to
termination on the SAME seeded instances the GPU tests and bench.py build (madipm_jl_b200/problems.py) and
its per-iterate scalars (objective, dual objective, inf_pr, inf_du, inf_compl, mu, step lengths) are written
to traces_full.json. The -m gpu tests compare the CUDA path per iterate at 1e-8 (north_star: "per-iterate primal
and dual objectives, residuals ... within 1e-8 relative, iteration count within +-2") for both host
sequencings. These are oracle outputs, NOT outputs of the Julia reference (no Julia in this image).

  c2_full   BASELINE configs[1]: LP m=200 000, n=1 000 000, 8 nnz/col (the bench workload), NormalKKTSystem
  c3_full   BASELINE configs[2]: QP n=500 000, m=150 000, K2 augmented system
  c4_s15    BASELINE configs[3] at scale 0.15 (10 commodities, 68 x 68 grid, 307 linking rows): the size the
            oracle's sequential LDL' (RCM ordering, dense border) finishes in minutes; the distributed solver is
            compared on it
  c4_s25    the same at scale 0.25 (16 commodities, 88 x 88 grid, 512 linking rows, m = 124 400): 40 minutes of oracle time
  c5_u0..7  BASELINE configs[4]: the first 8 units of the batch (m=500, n=2 000), NormalKKTSystem
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from madipm_jl_b200.problems import config_c2, config_c3, config_c4, config_c5  # noqa: E402
from oracle.mpc_oracle import madipm  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traces_full.json")

CASES = {
    "c2_full/Normal": (lambda: config_c2(), dict(kkt_system="Normal", linear_solver="ldl", fast_symbolic=True)),
    "c3_full/K2": (lambda: config_c3(), dict(kkt_system="K2")),
    "c4_s15/Normal": (lambda: config_c4(scale=0.15), dict(kkt_system="Normal", linear_solver="ldl", fast_symbolic=True)),
    # 40 minutes of oracle time (sequential LDL' with a dense border): only regenerated when asked for by name
    "c4_s25/Normal": (lambda: config_c4(scale=0.25), dict(kkt_system="Normal", linear_solver="ldl", fast_symbolic=True)),
}
for _i in range(8):
    CASES["c5_u%d/Normal" % _i] = ((lambda i=_i: config_c5(i)), dict(kkt_system="Normal"))

FIELDS = ("k", "objective", "dual_objective", "inf_pr", "inf_du", "inf_compl", "mu", "alpha_p", "alpha_d")


def main():
    want = sys.argv[1:]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for key, (make, opts) in CASES.items():
        if want and not any(key.startswith(w) for w in want):
            continue
        if not want and key.startswith("c4_s25") and key in out:
            continue
        t0 = time.time()
        qp = make()
        st = madipm(qp, **opts)
        out[key] = dict(
            status=st.status, iter=st.iter, objective=st.objective, dual_objective=st.dual_objective,
            problem=qp.name, oracle_options={k: v for k, v in opts.items()}, oracle_seconds=round(time.time() - t0, 1),
            trace=[{f: float(t[f]) for f in FIELDS} for t in st.trace])
        print(key, st.status, st.iter, repr(st.objective), "%.1fs" % (time.time() - t0), flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
