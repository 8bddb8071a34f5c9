"""ncu / phase-log driver for the K2 (LDL^T) system of config C3 (or C1 with `c1` as second argument).
Usage: python tools/profile_factor_k2.py [scale] [c1]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import _lib
from madipm_jl_b200.problems import config_c1, config_c3
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
qp = config_c1() if (len(sys.argv) > 2 and sys.argv[2] == 'c1') else config_c3(scale=scale)
n, m = qp.nvar, qp.ncon
I = np.concatenate([np.arange(n), qp.Hrows, n + qp.Arows, n + np.arange(m)]).astype(np.int32)
J = np.concatenate([np.arange(n), qp.Hcols, qp.Acols, n + np.arange(m)]).astype(np.int32)
h = _lib.Handle(device=0, stream=torch.cuda.current_stream().cuda_stream)
colptr, rowval, kmap = h.k2_symbolic(n + m, I, J)
rng = np.random.default_rng(1)
V = torch.from_numpy(np.concatenate([10.0 ** rng.uniform(-3, 3, n), qp.Hvals, qp.Avals, np.full(m, 1e-10)])).cuda()
nz = torch.zeros(len(rowval), dtype=torch.float64, device="cuda")
h.k2_transfer(V, nz)
t = time.time(); h.ls_analyze(n + m, colptr, rowval, kind=_lib.MIPM_LDL); print("analyze s", time.time() - t)
print(h.ls_stats())
for _ in range(2):
    assert h.ls_factorize(nz)
b = torch.randn(n + m, dtype=torch.float64, device="cuda")
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("factor ms", ev(lambda: h.ls_factorize_async(nz)))
x = b.clone(); print("solve ms", ev(lambda: h.ls_solve(x, 0)))
print(h.ls_factorize_profile(nz))
