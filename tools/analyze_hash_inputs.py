"""Writes the patterns tools/analyze_hash.bin is run on: C2 (full size), C1, a K2 system (LDL^T), a block-angular local
matrix with a border. Usage: python tools/analyze_hash_inputs.py <outdir>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madipm_jl_b200 import problems, _lib

out = sys.argv[1]
def dump(name, n, cp, ri, kind=0, ordering=0, nb=0):
    with open(os.path.join(out, name + ".bin"), "wb") as f:
        f.write(np.array([n, len(ri)], dtype=np.int64).tobytes())
        f.write(np.array([kind, ordering], dtype=np.int32).tobytes())
        f.write(np.array([nb], dtype=np.int64).tobytes())
        f.write(np.asarray(cp, dtype=np.int32).tobytes())
        f.write(np.asarray(ri, dtype=np.int32).tobytes())
def normal(qp):
    m, n = qp.ncon, qp.nvar
    Ap, Aj, _ = _lib.coo_to_csr(m, n, qp.Arows.astype(np.int32), qp.Acols.astype(np.int32))
    return _lib.Handle(device=-1).normal_symbolic(m, n, Ap, Aj)
qp = problems.config_c2(); Cp, Cj = normal(qp); dump("c2", qp.ncon, Cp, Cj)
qp = problems.config_c1(); Cp, Cj = normal(qp); dump("c1", qp.ncon, Cp, Cj)
qp = problems.random_sparse_qp(3000, 9000, 4, 2, structure="window", window=10)
n, m = qp.nvar, qp.ncon
I = np.concatenate([np.arange(n), qp.Hrows, n + qp.Arows, n + np.arange(m)]).astype(np.int32)
J = np.concatenate([np.arange(n), qp.Hcols, qp.Acols, n + np.arange(m)]).astype(np.int32)
colptr, rowval, _ = _lib.Handle(device=-1).k2_symbolic(n + m, I, J)
dump("k2", n + m, colptr, rowval, kind=1)
qp = problems.block_angular_lp(8, 24, 24, 96, 3); Cp, Cj = normal(qp); dump("c4s", qp.ncon, Cp, Cj, nb=96)
qp = problems.config_c2_mesh() if hasattr(problems, "config_c2_mesh") else None
if qp is not None:
    Cp, Cj = normal(qp); dump("mesh", qp.ncon, Cp, Cj)
