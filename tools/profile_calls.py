"""Per-entry-point wall-time profile of one solve (every C-ABI call synchronised and timed; perturbs the run).
Usage: python tools/profile_calls.py <c2|c4> [scale] [plain|distributed] [fused|unfused]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import problems
from madipm_jl_b200.solver import MPCSolver

def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
    fused = (sys.argv[4] if len(sys.argv) > 4 else "unfused") == "fused"
    qp = problems.config_c4(scale=scale) if cfg == "c4" else problems.config_c2()
    kw = dict(kkt_system="Normal", fused=fused)
    if mode == "distributed":
        kw.update(linear_solver="distributed", n_border=qp.meta["n_border"])
    s = MPCSolver(qp, **kw)
    s.solve()
    s.k = 0; s.trace = []
    acc, cnt = {}, {}
    def wrap(obj, name):
        fn = getattr(obj, name)
        def w(*a, **k):
            torch.cuda.synchronize(); t = time.perf_counter()
            out = fn(*a, **k)
            torch.cuda.synchronize()
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; cnt[name] = cnt.get(name, 0) + 1
            return out
        setattr(obj, name, w)
    for name in dir(s.h):
        if not name.startswith("_") and callable(getattr(s.h, name)):
            wrap(s.h, name)
    for name in ("init_starting_point", "initialize", "mpc"):
        wrap(s, name)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = s.solve()
    torch.cuda.synchronize(); total = time.perf_counter() - t0
    rows = sorted(acc.items(), key=lambda kv: -kv[1])
    print(json.dumps({"config": cfg, "scale": scale, "mode": mode, "fused": fused, "status": r.status, "iterations": r.iter,
                      "total_s": round(total, 4), "calls": {k: [round(v, 5), cnt[k]] for k, v in rows[:25]}}))

if __name__ == "__main__":
    main()
