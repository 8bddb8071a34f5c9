"""Where the end-to-end solve() time of C2 goes: initialize (H2D + start point), mpc loop, result D2H + post-processing.
Usage: python tools/profile_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madipm_jl_b200 import problems
from madipm_jl_b200.solver import MPCSolver

qp = problems.config_c2()
s = MPCSolver(qp, kkt_system="Normal")
s.solve()
for rep in range(3):
    s.k = 0; s.trace = []
    marks = {}
    orig_init, orig_mpc = s.initialize, s.mpc
    def init_():
        t = time.perf_counter(); orig_init(); torch.cuda.synchronize(); marks["initialize"] = time.perf_counter() - t
    def mpc_():
        t = time.perf_counter(); orig_mpc(); torch.cuda.synchronize(); marks["mpc"] = time.perf_counter() - t
    s.initialize, s.mpc = init_, mpc_
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = s.solve()
    torch.cuda.synchronize(); total = time.perf_counter() - t0
    s.initialize, s.mpc = orig_init, orig_mpc
    print("total %.2f ms | initialize %.2f | mpc %.2f (%d iterations, %.3f ms each) | rest (D2H + stats) %.2f" % (
        1e3 * total, 1e3 * marks["initialize"], 1e3 * marks["mpc"], r.iter, 1e3 * marks["mpc"] / r.iter,
        1e3 * (total - marks["initialize"] - marks["mpc"])))
