"""Sweep nested-dissection leaf size / amalgamation on the C2 normal matrix: symbolic stats + factor / solve time."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import _lib
from madipm_jl_b200.problems import config_c2
qp = config_c2(seed=2)
m, n = qp.ncon, qp.nvar
Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
h0 = _lib.Handle(device=0, stream=torch.cuda.current_stream().cuda_stream)
Cp, Cj = h0.normal_symbolic(m, n, Bp, Bj)
ATx = torch.from_numpy(qp.Avals[Bm]).cuda()
pr = torch.from_numpy(np.random.default_rng(0).uniform(1e-2, 1e2, n)).cuda()
Cx = torch.zeros(len(Cj), dtype=torch.float64, device="cuda")
h0.normal_set_jacobian(ATx); h0.normal_assemble(pr, Cx)
b = torch.randn(m, dtype=torch.float64, device="cuda")
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for leaf in (48, 96, 160, 256, 400):
    for relax in ("8,32,0.5,96,0.2,0.05", "16,64,0.5,128,0.3,0.1", "4,16,0.3,48,0.1,0.05"):
        os.environ["MIPM_ND_LEAF"] = str(leaf); os.environ["MIPM_RELAX"] = relax
        h = _lib.Handle(device=0, stream=torch.cuda.current_stream().cuda_stream)
        t = time.time(); h.ls_analyze(m, Cp, Cj); ta = time.time() - t
        st = h.ls_stats()
        assert h.ls_factorize(Cx)
        tf = ev(lambda: h.ls_factorize_async(Cx))
        x = b.clone(); ts = ev(lambda: h.ls_solve(x, 0))
        print(json.dumps(dict(leaf=leaf, relax=relax, analyze_s=round(ta, 2), factor_ms=round(tf, 3), solve_ms=round(ts, 3),
                              nnz_l=st["nnz_l"], gflop=round(st["flops"] / 1e9, 2), sn=st["n_supernodes"], levels=st["n_levels"],
                              maxk=st["max_front_cols"], upd_MB=round(st["update_doubles"] * 8e-6))), flush=True)
        del h
