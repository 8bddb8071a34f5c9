"""Small driver for ncu: C2 (or a scaled C2) normal matrix -> analyze, 2 factorizations, 2 solves.
Usage: python tools/profile_factor.py [scale]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import _lib  # noqa: E402
from madipm_jl_b200.problems import config_c2, config_c2_mesh  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    qp = config_c2_mesh(scale=scale) if os.environ.get("PROFILE_CONFIG") == "mesh" else config_c2(seed=2, scale=scale)
    m, n = qp.ncon, qp.nvar
    Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
    h = _lib.Handle(device=0, stream=torch.cuda.current_stream().cuda_stream)
    Cp, Cj = h.normal_symbolic(m, n, Bp, Bj)
    ATx = torch.from_numpy(qp.Avals[Bm]).cuda()
    pr = torch.from_numpy(np.random.default_rng(0).uniform(1e-2, 1e2, n)).cuda()
    Cx = torch.zeros(len(Cj), dtype=torch.float64, device="cuda")
    h.normal_set_jacobian(ATx)
    h.ls_analyze(m, Cp, Cj)
    print(h.ls_stats())
    for _ in range(2):
        h.normal_assemble(pr, Cx)
        assert h.ls_factorize(Cx)
    b = torch.randn(m, dtype=torch.float64, device="cuda")
    for _ in range(2):
        x = b.clone()
        h.ls_solve(x, 0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    h.ls_factorize(Cx)
    torch.cuda.synchronize()
    print("factor ms", 1e3 * (time.perf_counter() - t))
    for _ in range(3):
        pr_ = h.ls_factorize_profile(Cx)
        print({k_: round(v["ms"], 4) for k_, v in pr_.items()}, "update TF/s (busy-time rate x grid): %.2f" % (
            pr_["update"]["work"] / max(pr_["update"]["ms"], 1e-9) / 1e9), "factor TF/s: %.2f" % (h.ls_stats()["flops"] / pr_["kernel"]["ms"] / 1e9))
    out = os.environ.get("PROFILE_OUT")
    if out:
        os.environ["MIPM_TASK_TRACE"] = out + "_trace.csv"
        os.environ["MIPM_PHASE_LOG"] = out + "_levels.csv"
        print(h.ls_factorize_profile(Cx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        h.ls_factorize_async(Cx)
    e1.record()
    torch.cuda.synchronize()
    print("factorize (events, avg of 5) ms", e0.elapsed_time(e1) / 5)
    e0.record()
    for _ in range(5):
        x = b.clone()
        h.ls_solve(x, 0)
    e1.record()
    torch.cuda.synchronize()
    print("solve (events, avg of 5, incl. clone) ms", e0.elapsed_time(e1) / 5)
    if out:
        os.environ["MIPM_SOLVE_TRACE"] = out + "_solve_trace.csv"
        x = b.clone()
        h.ls_solve(x, 0)
        torch.cuda.synchronize()
        del os.environ["MIPM_SOLVE_TRACE"]


if __name__ == "__main__":
    main()
