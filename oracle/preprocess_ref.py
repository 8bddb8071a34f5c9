"""oracle/preprocess_ref.py -- TEST INFRASTRUCTURE ONLY (CPU, numpy).

Sequential restatement of the Ruiz equilibration that scale_qp (klamike/MadIPM.jl scripts/common.jl:57-100) obtains from
HSL.mc77(A, 0). HSL is closed source and not in the image, so the published algorithm is restated (D. Ruiz, "A scaling
algorithm to equilibrate both rows and columns norms in matrices", RAL-TR-2001-034; infinity-norm variant): PARITY
UNPINNED against MC77 itself; the CUDA path (mipm_ruiz_equilibrate) is pinned to this restatement bit for bit.
Only tests/ may import this module."""
import numpy as np


def ruiz_equilibrate(m, n, rows, cols, vals, max_iter=10, tol=0.0):
    dr, dc = np.ones(m), np.ones(n)
    it = 0
    while it < max_iter:
        v = (np.abs(vals) / dr[rows]) / dc[cols]
        r, c = np.zeros(m), np.zeros(n)
        np.maximum.at(r, rows, v)
        np.maximum.at(c, cols, v)
        dev = 0.0
        pr, pc = r > 0, c > 0
        dr[pr] = dr[pr] * np.sqrt(r[pr])
        dc[pc] = dc[pc] * np.sqrt(c[pc])
        if pr.any():
            dev = max(dev, np.abs(1.0 - r[pr]).max())
        if pc.any():
            dev = max(dev, np.abs(1.0 - c[pc]).max())
        it += 1
        if tol > 0.0 and dev <= tol:
            break
    return dr, dc, it
