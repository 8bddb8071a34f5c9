"""Driver for ncu / event timing of the streaming SpMV kernels on C2: A x, A' y, the pair launch and the assembly.
Usage: python tools/profile_spmv.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import _lib
from madipm_jl_b200.problems import config_c2

qp = config_c2()
m, n = qp.ncon, qp.nvar
Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
h = _lib.Handle(device=0, stream=torch.cuda.current_stream().cuda_stream)
h.spmv_setup(m, n, Bp, Bj)
Cp, Cj = h.normal_symbolic(m, n, Bp, Bj)
ATx = torch.from_numpy(qp.Avals[Bm]).cuda()
h.spmv_cache_values(ATx)
h.normal_set_jacobian(ATx)
rng = np.random.default_rng(0)
x, y = torch.from_numpy(rng.standard_normal(n)).cuda(), torch.from_numpy(rng.standard_normal(m)).cuda()
ox, oy = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(m, dtype=torch.float64, device="cuda")
pr = torch.from_numpy(rng.uniform(1e-2, 1e2, n)).cuda()
Cx = torch.zeros(len(Cj), dtype=torch.float64, device="cuda")
flush = torch.zeros(64 << 20, dtype=torch.float64, device="cuda")       # 512 MB > L2


def timed(name, fn, bytes_):
    ts = []
    for _ in range(8):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[2:]))
    print("%-10s %.4f ms  %.0f GB/s" % (name, t, bytes_ / t / 1e6))


nnz, T, nnzc = len(Bj), None, len(Cj)
timed("A x", lambda: h.spmv(0, 1.0, ATx, x, 0.0, oy), 12 * nnz + 8 * (m + n))
timed("A' y", lambda: h.spmv(1, 1.0, ATx, y, 0.0, ox), 12 * nnz + 8 * (m + n))
timed("pair", lambda: h.spmv_pair(ATx, 1.0, x, 0.0, oy, 1.0, y, 0.0, ox), 2 * (12 * nnz + 8 * (m + n)))
timed("assemble", lambda: h.normal_assemble(pr, Cx), 602503640)
