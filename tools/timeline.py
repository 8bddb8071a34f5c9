"""Kernel timeline of fused C2 iterations through CUPTI (torch.profiler): per-kernel device time and the idle gaps
between kernels. Diagnostic only (numbers under a profiler are never bench values).
Usage: python tools/timeline.py [n_iters]"""
import json, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import problems
from madipm_jl_b200.solver import MPCSolver
from torch.profiler import profile, ProfilerActivity

def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    qp = problems.config_c2()
    s = MPCSolver(qp, kkt_system="Normal", max_iter=6)
    s.solve()                       # warm: leaves the solver mid-run at iteration 6
    s.opt.max_iter = 10 ** 6
    from madipm_jl_b200 import solver as _sv
    s.status = _sv.REGULAR
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(iters):
            s.mpc_iteration()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    tot = collections.OrderedDict()
    gaps = 0.0; last_end = None; big = []
    for e in ev:
        d = e.time_range.end - e.time_range.start
        k = e.name[:70]
        a = tot.setdefault(k, [0.0, 0]); a[0] += d; a[1] += 1
        if last_end is not None and e.time_range.start > last_end:
            g = e.time_range.start - last_end
            gaps += g
            if g > 20: big.append((round(g, 1), prevname[:40], k[:40]))
        last_end = max(last_end or 0, e.time_range.end); prevname = k
    span = ev[-1].time_range.end - ev[0].time_range.start
    print(json.dumps({"iters": iters, "span_us_per_iter": span / iters, "gap_us_per_iter": gaps / iters,
                      "kernels_us_per_iter": {k: [round(v[0] / iters, 1), v[1] / iters] for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])},
                      "big_gaps": big[:40]}, indent=1))

if __name__ == "__main__":
    main()
