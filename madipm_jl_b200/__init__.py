"""madipm_jl_b200 -- B200-native (sm_100a) replacement for the per-iteration linear-algebra
hot path of klamike/MadIPM.jl, behind a C ABI (include/madipm_b200.h).

Layout
  csrc/          CUDA kernels, host symbolic analysis (C++) and the C-ABI entry points
  _lib.py        ctypes binding of the C ABI (what Julia's ccall would bind)
  solver.py      host-side mirror of MPCSolver / madipm / solve! driving the C ABI
  problems.py    QuadraticModel container + seeded synthetic generators (numpy only)

Importing this package never touches the GPU; `solver` needs torch + a CUDA device and fails
loudly without them (no CPU fallback).
"""
from .problems import QuadraticModel, simple_lp, random_sparse_lp, random_sparse_qp  # noqa: F401


def __getattr__(name):
    if name in ("MPCSolver", "madipm", "solve", "IPMOptions", "FixedRegularization", "AdaptiveRegularization",
                "NoRegularization", "AdaptiveStep", "ConservativeStep", "MehrotraAdaptiveStep", "B200Solver"):
        from . import solver
        return getattr(solver, name)
    raise AttributeError(name)
