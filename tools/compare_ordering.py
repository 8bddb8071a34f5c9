"""Fill of the library's own nested dissection against cuSOLVER's METIS / SYMAMD / SYMRCM host orderings
(cusolverSpXcsrmetisndHost, cusolverSpXcsrsymamdHost, cusolverSpXcsrsymrcmHost: comparators only, never on the product
path). Each foreign permutation is fed back through mipm_ls_analyze(ordering = USER), so nnz(L) and flops come from the
same symbolic code. Usage: python tools/compare_ordering.py [c2|mesh|c1] [scale]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import _lib  # noqa: E402
from madipm_jl_b200 import problems  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
qp = {"c2": lambda: problems.config_c2(scale=scale), "mesh": lambda: problems.config_c2_mesh(scale=scale),
      "c1": lambda: problems.config_c1()}[which]()
m, n = qp.ncon, qp.nvar
Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
h = _lib.Handle(device=-1)
Cp, Cj = h.normal_symbolic(m, n, Bp, Bj)


def stats(perm=None, ordering=_lib.MIPM_ORDER_ND):
    hh = _lib.Handle(device=-1)
    t = time.time()
    hh.ls_analyze(m, Cp, Cj, ordering=ordering, user_perm=perm)
    st = hh.ls_stats()
    return {"nnz_l_exact": st["nnz_l_exact"], "nnz_l_stored": st["nnz_l"], "gflop": st["flops"] / 1e9, "supernodes": st["n_supernodes"],
            "levels": st["n_levels"], "max_front_cols": st["max_front_cols"], "analyze_s": round(time.time() - t, 2)}


out = {"workload": qp.name, "m": m, "nnz_tril": len(Cj), "own_nested_dissection": stats()}
# full symmetric pattern in CSR for cuSOLVER
low = sp.csc_matrix((np.ones(len(Cj)), Cj, Cp), shape=(m, m))
full = (low + sp.tril(low, -1).T).tocsr()
full.sort_indices()
ia, ja = full.indptr.astype(np.int32), full.indices.astype(np.int32)
try:
    lib = C.CDLL("libcusolver.so.11")
    hs = C.c_void_p()
    assert lib.cusolverSpCreate(C.byref(hs)) == 0
    lib2 = C.CDLL("libcusparse.so.12")
    descr = C.c_void_p()
    assert lib2.cusparseCreateMatDescr(C.byref(descr)) == 0
    for name, fn, extra in (("metis_nd", "cusolverSpXcsrmetisndHost", True), ("symamd", "cusolverSpXcsrsymamdHost", False),
                            ("symrcm", "cusolverSpXcsrsymrcmHost", False)):
        p = np.zeros(m, dtype=np.int32)
        t = time.time()
        f = getattr(lib, fn)
        args = [hs, C.c_int(m), C.c_int(len(ja)), descr, ia.ctypes.data_as(C.c_void_p), ja.ctypes.data_as(C.c_void_p)]
        if extra:
            args.append(C.c_void_p(0))          # default METIS options
        args.append(p.ctypes.data_as(C.c_void_p))
        rc = f(*args)
        dt = time.time() - t
        if rc != 0:
            out[name] = {"error": rc}
            continue
        st = stats(perm=p, ordering=_lib.MIPM_ORDER_USER)
        st["ordering_s"] = round(dt, 2)
        out[name] = st
except Exception as exc:      # noqa: BLE001
    out["cusolver"] = repr(exc)
print(json.dumps(out))
