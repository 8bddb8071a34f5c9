"""Model preprocessing around the hot path.

    standard_form_qp(qp)     src/utils.jl:346-505   inequality rows get slack columns, range bounds move into rows
    scale_qp(qp)             scripts/common.jl:47-100   Ruiz equilibration; the sweeps run on the device
                             (mipm_ruiz_equilibrate / mipm_scale_coo instead of HSL.mc77 + _scale_coo!)
    to_device(qp)            ext/MadIPMCUDAExt/MadIPMCUDAExt.jl:122-137   host model -> device-resident solver state;
                             here this is MPCSolver(qp): the numeric data is staged in ONE pinned slab and uploaded at
                             the start of every solve() (solver._stage_problem / _madnlp_initialize)

Index work (building the reformulated pattern) stays on the host like in the reference, where it runs once per model.
"""
import numpy as np

from .problems import QuadraticModel


def standard_form_qp(qp):
    """src/utils.jl:373-505: min c'x + x'Hx/2  s.t.  A x - s = 0 (inequality rows), x + w = xu (range-bounded x or s),
    equality rows kept; variables [x; s; w] with w >= 0, s in [lcon, ucon] (range upper bounds moved to rows)."""
    n, m = qp.nvar, qp.ncon
    lvar, uvar, lcon, ucon = qp.lvar, qp.uvar, qp.lcon, qp.ucon
    ind_ineq = np.flatnonzero(lcon < ucon)
    ns = len(ind_ineq)
    fixed = lvar == uvar
    rng_x = np.flatnonzero(~fixed & np.isfinite(lvar) & np.isfinite(uvar) & (lvar < uvar))
    rng_s = np.flatnonzero(np.isfinite(lcon[ind_ineq]) & np.isfinite(ucon[ind_ineq]))        # (lcon < ucon holds on ind_ineq)
    ind_rng = np.concatenate([rng_x, n + rng_s]).astype(np.int64)
    xu = np.concatenate([uvar[rng_x], ucon[ind_ineq][rng_s]])
    nw = len(ind_rng)
    nvar, ncon = n + ns + nw, m + nw
    # slack contribution A x - s = 0, then the range rows x + w = xu
    Bi = np.concatenate([ind_ineq, np.repeat(m + np.arange(nw), 2)])
    Bj = np.concatenate([n + np.arange(ns), np.stack([ind_rng, n + ns + np.arange(nw)], axis=1).reshape(-1)])
    Bx = np.concatenate([-np.ones(ns), np.ones(2 * nw)])
    lcon_ = np.concatenate([np.where(lcon < ucon, 0.0, lcon), xu])
    ucon_ = np.concatenate([np.where(lcon < ucon, 0.0, ucon), xu])
    lvar_ = np.concatenate([lvar, lcon[ind_ineq], np.zeros(nw)])
    uvar_ = np.concatenate([uvar, ucon[ind_ineq], np.full(nw, np.inf)])
    uvar_[ind_rng] = np.inf                      # the upper bounds of range variables now live in their own rows
    uvar_[np.flatnonzero(fixed)] = uvar[fixed]   # fixed variables stay in the formulation
    z = np.zeros(ns + nw)
    return QuadraticModel(
        c=np.concatenate([qp.c, z]), Hrows=qp.Hrows, Hcols=qp.Hcols, Hvals=qp.Hvals,
        Arows=np.concatenate([qp.Arows, Bi]), Acols=np.concatenate([qp.Acols, Bj]), Avals=np.concatenate([qp.Avals, Bx]),
        lcon=lcon_, ucon=ucon_, lvar=lvar_, uvar=uvar_, c0=qp.c0, x0=np.concatenate([qp.x0, z]),
        y0=np.concatenate([qp.y0, np.zeros(nw)]), name=qp.name + "_std", minimize=qp.minimize)


def scale_qp(qp, max_iter=10, tol=0.0, device=0, return_scaling=False):
    """scripts/common.jl:57-100 with the Ruiz sweeps on the device: As = A ./ (Dr_i Dc_j), Hs = H ./ (Dc_i Dc_j),
    c ./ Dc, bounds .* Dc, row bounds ./ Dr, x0 .* Dc, y0 ./ Dr. The reference skips scaling when HSL is missing; here the
    library's own sweep runs (MC77's defaults: 10 sweeps, no convergence test)."""
    import torch
    from . import _lib
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    h = _lib.Handle(device=device, stream=torch.cuda.current_stream(dev).cuda_stream)
    m, n = qp.ncon, qp.nvar
    to = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    ai, aj, av = to(qp.Arows, torch.int32), to(qp.Acols, torch.int32), to(qp.Avals, torch.float64)
    Dr = torch.ones(max(m, 1), dtype=torch.float64, device=dev)
    Dc = torch.ones(max(n, 1), dtype=torch.float64, device=dev)
    h.ruiz_equilibrate(m, n, ai, aj, av, Dr, Dc, max_iter=max_iter, tol=tol)
    As = torch.empty_like(av)
    h.scale_coo(ai, aj, av, Dr, Dc, As)
    Hs = qp.Hvals
    if qp.nnzh > 0:
        hi, hj, hv = to(qp.Hrows, torch.int32), to(qp.Hcols, torch.int32), to(qp.Hvals, torch.float64)
        Ht = torch.empty_like(hv)
        h.scale_coo(hi, hj, hv, Dc, Dc, Ht)
        Hs = Ht.cpu().numpy()
    dr, dc = Dr.cpu().numpy()[:m], Dc.cpu().numpy()[:n]
    out = QuadraticModel(
        c=qp.c / dc, Hrows=qp.Hrows, Hcols=qp.Hcols, Hvals=Hs, Arows=qp.Arows, Acols=qp.Acols, Avals=As.cpu().numpy(),
        lcon=qp.lcon / dr, ucon=qp.ucon / dr, lvar=qp.lvar * dc, uvar=qp.uvar * dc, c0=qp.c0, x0=qp.x0 * dc, y0=qp.y0 / dr,
        name=qp.name + "_ruiz", minimize=qp.minimize)
    return (out, dr, dc) if return_scaling else out


def to_device(qp, **kwargs):
    """convert(QuadraticModel{T, CuVector{T}}, qp) + MPCSolver: the device-resident solver of a host model."""
    from .solver import MPCSolver
    return MPCSolver(qp, **kwargs)
