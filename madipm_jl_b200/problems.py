"""Problem container and seeded synthetic generators.

`QuadraticModel` mirrors the fields of QuadraticModels.QuadraticModel that MadIPM reads
(reference call sites: /root/reference/test/runtests.jl:29-60, src/structure.jl:79-178):

    min  c0 + c'x + 1/2 x'Hx   s.t.  lcon <= A x <= ucon,  lvar <= x <= uvar

H is given by its LOWER triangle in COO form, A in COO form, both 0-based here (the Julia
glue converts from 1-based). Pure numpy: shared byte-for-byte by the CUDA path, the oracle
and the benchmarks, and importable without a GPU.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class QuadraticModel:
    c: np.ndarray
    Hrows: np.ndarray
    Hcols: np.ndarray
    Hvals: np.ndarray
    Arows: np.ndarray
    Acols: np.ndarray
    Avals: np.ndarray
    lcon: np.ndarray
    ucon: np.ndarray
    lvar: np.ndarray
    uvar: np.ndarray
    c0: float = 0.0
    x0: np.ndarray = None
    y0: np.ndarray = None
    name: str = "qp"
    minimize: bool = True
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.c = f64(self.c)
        self.Hrows, self.Hcols, self.Hvals = i32(self.Hrows), i32(self.Hcols), f64(self.Hvals)
        self.Arows, self.Acols, self.Avals = i32(self.Arows), i32(self.Acols), f64(self.Avals)
        self.lcon, self.ucon = f64(self.lcon), f64(self.ucon)
        self.lvar, self.uvar = f64(self.lvar), f64(self.uvar)
        self.x0 = np.zeros(self.nvar) if self.x0 is None else f64(self.x0)
        self.y0 = np.zeros(self.ncon) if self.y0 is None else f64(self.y0)
        assert len(self.Hrows) == len(self.Hcols) == len(self.Hvals)
        assert len(self.Arows) == len(self.Acols) == len(self.Avals)
        assert np.all(self.Hrows >= self.Hcols), "H must be given by its lower triangle"

    @property
    def nvar(self):
        return len(self.c)

    @property
    def ncon(self):
        return len(self.lcon)

    @property
    def nnzj(self):
        return len(self.Avals)

    @property
    def nnzh(self):
        return len(self.Hvals)


def simple_lp():
    """The reference's only self-contained fixture: test/runtests.jl:29-60.
    min x1 + x2  s.t.  x1 + x2 = 1, x >= 0;  optimal objective 1.0."""
    return QuadraticModel(
        c=np.ones(2), Hrows=[], Hcols=[], Hvals=[],
        Arows=[0, 0], Acols=[0, 1], Avals=[1.0, 1.0],
        lcon=[1.0], ucon=[1.0], lvar=[0.0, 0.0], uvar=[np.inf, np.inf],
        c0=0.0, x0=np.ones(2), name="simpleLP",
    )


def _sparsity(rng, m, n, k, structure, window):
    """k structural nonzeros per column. 'uniform': rows uniform in [0,m).
    'window': rows drawn within +-window of the column's home row j*m/n."""
    k = min(k, m)
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    if structure == "uniform":
        # k distinct rows per column: sample with a random offset + distinct strides trick is
        # biased; use argpartition on random keys for small m, rejection for large m.
        rows = rng.integers(0, m, size=(n, k))
        for _ in range(64):
            srt = np.sort(rows, axis=1)
            dup = np.zeros_like(rows, dtype=bool)
            dup[:, 1:] = srt[:, 1:] == srt[:, :-1]
            if not dup.any():
                rows = srt
                break
            srt[dup] = rng.integers(0, m, size=int(dup.sum()))
            rows = srt
        else:
            raise RuntimeError("could not draw distinct rows")
    elif structure == "window":
        w = min(window, (m - 1) // 2)
        assert 2 * w + 1 >= k
        home = (np.arange(n, dtype=np.int64) * m) // n
        lo = np.clip(home - w, 0, m - (2 * w + 1))
        # k distinct offsets in [0, 2w]: take the k smallest of 2w+1 random keys
        keys = rng.random((n, 2 * w + 1))
        off = np.argpartition(keys, k - 1, axis=1)[:, :k]
        rows = np.sort(lo[:, None] + off, axis=1)
    else:
        raise ValueError(structure)
    return rows.reshape(-1).astype(np.int32), cols.astype(np.int32)


def random_sparse_lp(m, n, k, seed, structure="window", window=50, ub_fraction=0.0):
    """Feasible-and-bounded-by-construction LP:
    A: k nnz/col, values N(0,1); x* ~ U(.5,1.5), b = A x*; y* ~ N(0,1), z* ~ U(0,1),
    c = A'y* + z*; bounds x >= 0 (a fraction `ub_fraction` of variables also get x <= 2);
    all constraints are equalities. max|A_ij| and ||c||_inf stay far below 100 so MadNLP's
    gradient-based scaling is the identity."""
    rng = np.random.default_rng(seed)
    rows, cols = _sparsity(rng, m, n, k, structure, window)
    vals = rng.standard_normal(len(rows))
    # keep |A_ij| away from 0 so no structural entry is numerically void
    vals = np.sign(vals) * np.maximum(np.abs(vals), 1e-2)
    xs = rng.uniform(0.5, 1.5, n)
    b = np.zeros(m)
    np.add.at(b, rows, vals * xs[cols])
    ys = rng.standard_normal(m)
    zs = rng.uniform(0.0, 1.0, n)
    c = zs.copy()
    np.add.at(c, cols, vals * ys[rows])
    uvar = np.full(n, np.inf)
    if ub_fraction > 0:
        pick = rng.random(n) < ub_fraction
        uvar[pick] = 2.0
    return QuadraticModel(
        c=c, Hrows=[], Hcols=[], Hvals=[], Arows=rows, Acols=cols, Avals=vals,
        lcon=b, ucon=b.copy(), lvar=np.zeros(n), uvar=uvar, x0=np.zeros(n),
        name=f"lp_{structure}_m{m}_n{n}_k{k}_s{seed}",
        meta=dict(m=m, n=n, k=k, seed=seed, structure=structure, window=window),
    )


def random_sparse_qp(m, n, k, seed, structure="window", window=50, q_offdiag=2):
    """Convex QP (config C3): A as in random_sparse_lp; Q sparse, symmetric, strictly
    diagonally dominant (hence PSD) with `q_offdiag` off-diagonals per column placed within
    the same locality window in variable space."""
    rng = np.random.default_rng(seed)
    lp = random_sparse_lp(m, n, k, seed, structure, window)
    hr, hc, hv = [np.arange(n, dtype=np.int64)], [np.arange(n, dtype=np.int64)], []
    diag = np.full(n, 1e-3)
    if q_offdiag > 0:
        j = np.repeat(np.arange(n, dtype=np.int64), q_offdiag)
        wv = max(1, min(window, n - 1))
        i = j + rng.integers(1, wv + 1, size=len(j))
        keep = i < n
        i, j = i[keep], j[keep]
        # merge duplicates (i,j)
        key = i * n + j
        _, first = np.unique(key, return_index=True)
        i, j = i[first], j[first]
        v = 0.1 * rng.standard_normal(len(i))
        np.add.at(diag, i, np.abs(v))
        np.add.at(diag, j, np.abs(v))
        hr.append(i), hc.append(j), hv.append(v)
    hv.insert(0, diag + rng.uniform(0.0, 1.0, n))
    Hrows, Hcols, Hvals = np.concatenate(hr), np.concatenate(hc), np.concatenate(hv)
    # shift c so the LP-feasible x* stays dual feasible: c_qp = c_lp - Q x*  is not needed for
    # boundedness (Q PSD, x >= 0 region, feasible set non-empty) -> keep c.
    return QuadraticModel(
        c=lp.c, Hrows=Hrows, Hcols=Hcols, Hvals=Hvals, Arows=lp.Arows, Acols=lp.Acols,
        Avals=lp.Avals, lcon=lp.lcon, ucon=lp.ucon, lvar=lp.lvar, uvar=lp.uvar, x0=lp.x0,
        name=f"qp_{structure}_m{m}_n{n}_k{k}_s{seed}",
        meta=dict(m=m, n=n, k=k, seed=seed, structure=structure, window=window),
    )


# The five BASELINE.json configs. C2's structure is stated explicitly: uniformly random rows
# make chol(A A') ~90% dense (1.8e10 nonzeros at m=2e5), so the named shape is
# generated with windowed rows (+-50 around the column's home row), as SURVEY 7.4 measured.
def config_c1(seed=1):
    return random_sparse_lp(2_000, 10_000, 5, seed, structure="uniform")


def config_c2(seed=2, scale=1.0):
    m, n = int(200_000 * scale), int(1_000_000 * scale)
    return random_sparse_lp(m, n, 8, seed, structure="window", window=50)


def mesh_sparse_lp(grid, cols_per_cell, k, radius, seed):
    """LP whose rows are the cells of a grid x grid mesh: every column has k nonzeros in distinct cells within +-radius
    (both directions) of its home cell. A A' then is a 2-D mesh graph with a thick stencil, whose nested-dissection
    separators grow like grid * 2 * radius: fronts with thousands of columns, the regime where the factorization is
    bound by the FP64 tensor pipe instead of latency (the band structure of config_c2 keeps every front below 300
    columns). Feasible and bounded by construction like random_sparse_lp."""
    rng = np.random.default_rng(seed)
    m = grid * grid
    n = m * cols_per_cell
    home = np.repeat(np.arange(m, dtype=np.int64), cols_per_cell)
    hy, hx = home // grid, home % grid
    w = 2 * radius + 1
    assert w * w >= k
    keys = rng.random((n, w * w))
    off = np.argpartition(keys, k - 1, axis=1)[:, :k]
    dy, dx = off // w - radius, off % w - radius
    cy = np.clip(hy[:, None] + dy, 0, grid - 1)
    cx = np.clip(hx[:, None] + dx, 0, grid - 1)
    rows = np.sort(cy * grid + cx, axis=1)
    # clipping at the mesh boundary can repeat a cell inside a column: keep the first occurrence of each
    keep = np.ones_like(rows, dtype=bool)
    keep[:, 1:] = rows[:, 1:] != rows[:, :-1]
    cols = np.repeat(np.arange(n, dtype=np.int64), k).reshape(n, k)
    rows, cols = rows[keep], cols[keep]
    vals = rng.standard_normal(len(rows))
    vals = np.sign(vals) * np.maximum(np.abs(vals), 1e-2)
    xs = rng.uniform(0.5, 1.5, n)
    b = np.zeros(m)
    np.add.at(b, rows, vals * xs[cols])
    ys = rng.standard_normal(m)
    zs = rng.uniform(0.0, 1.0, n)
    c = zs.copy()
    np.add.at(c, cols, vals * ys[rows])
    return QuadraticModel(c=c, Hrows=[], Hcols=[], Hvals=[], Arows=rows, Acols=cols, Avals=vals, lcon=b, ucon=b.copy(),
                          lvar=np.zeros(n), uvar=np.full(n, np.inf), x0=np.zeros(n),
                          name=f"lp_mesh_g{grid}_c{cols_per_cell}_k{k}_r{radius}_s{seed}",
                          meta=dict(m=m, n=n, k=k, seed=seed, structure="mesh", grid=grid, radius=radius))


def config_c2_mesh(seed=6, scale=1.0):
    """The harder sibling of config_c2 (same 8 nnz/col, 5 columns per row): 2-D mesh locality, root separator ~2000 columns."""
    return mesh_sparse_lp(max(8, int(round(256 * np.sqrt(scale)))), 5, 8, 3, seed)


def config_c3(seed=3, scale=1.0):
    m, n = int(150_000 * scale), int(500_000 * scale)
    return random_sparse_qp(m, n, 5, seed, structure="window", window=50)


def config_c5(index, seed=5):
    return random_sparse_lp(500, 2_000, 5, seed * 100_003 + index, structure="uniform")


def mixed_bounds_lp(m, n, k, seed):
    """Small LP exercising every bound / constraint kind the reference handles: equality, range,
    upper-only and lower-only rows (-> slack variables), and variables that are lower-bounded,
    boxed, upper-only or free. Feasible (built around x*) and bounded (c = A_eq' y* + z*, z* > 0 on
    the variables with an infinite upper bound, free variables get exactly A_eq' y*)."""
    rng = np.random.default_rng(seed)
    rows, cols = _sparsity(rng, m, n, k, "uniform", 0)
    vals = rng.standard_normal(len(rows))
    xs = rng.uniform(0.5, 1.5, n)
    b = np.zeros(m)
    np.add.at(b, rows, vals * xs[cols])
    kind = rng.integers(0, 4, m)                 # 0 equality, 1 range, 2 upper-only, 3 lower-only
    kind[: max(1, m // 3)] = 0
    lcon = np.where(kind == 0, b, np.where(kind == 1, b - 1.0, np.where(kind == 2, -np.inf, b - 0.5)))
    ucon = np.where(kind == 0, b, np.where(kind == 1, b + 1.0, np.where(kind == 2, b + 0.5, np.inf)))
    vk = rng.integers(0, 4, n)                   # 0 x>=0, 1 boxed, 2 upper-only, 3 free
    vk[rng.random(n) < 0.5] = 0
    lvar = np.where(vk == 0, 0.0, np.where(vk == 1, -2.0, -np.inf))
    uvar = np.where(vk == 0, np.inf, np.where(vk == 1, 3.0, np.where(vk == 2, 2.5, np.inf)))
    ys = np.where(kind == 0, rng.standard_normal(m), 0.0)
    c = np.zeros(n)
    np.add.at(c, cols, vals * ys[rows])
    zs = rng.uniform(0.1, 1.0, n)
    c += np.where(vk == 0, zs, np.where(vk == 2, -zs, np.where(vk == 1, rng.standard_normal(n), 0.0)))
    return QuadraticModel(c=c, Hrows=[], Hcols=[], Hvals=[], Arows=rows, Acols=cols, Avals=vals, lcon=lcon, ucon=ucon,
                          lvar=lvar, uvar=uvar, x0=np.zeros(n), name=f"mixed_lp_m{m}_n{n}_s{seed}")


def badly_scaled_lp(m, n, k, seed):
    """mixed_bounds_lp with a third of the rows (inequality rows among them) multiplied by 1e3: max |A_ij| > 100 on those rows,
    so MadNLP's constraint scaling (con_scale_i = 100 / max_j |A_ij|) is active, on slack rows too."""
    qp = mixed_bounds_lp(m, n, k, seed)
    big = np.zeros(m, dtype=bool)
    big[::3] = True
    f = np.where(big, 1.0e3, 1.0)
    return QuadraticModel(c=qp.c, Hrows=[], Hcols=[], Hvals=[], Arows=qp.Arows, Acols=qp.Acols, Avals=qp.Avals * f[qp.Arows],
                          lcon=qp.lcon * f, ucon=qp.ucon * f, lvar=qp.lvar, uvar=qp.uvar, x0=qp.x0, name=f"badscale_lp_m{m}_n{n}_s{seed}")


def degenerate_lp(m, n, k, seed, n_dup=0, cond=1.0):
    """LP that stresses the linear algebra instead of being a friendly random instance:
      * primal degenerate: only m // 2 entries of x* are positive (fewer than m), the rest sit on their bound;
      * dual degenerate: a tenth of the zero entries also has z* = 0 (no strict complementarity there);
      * `cond` > 1 multiplies the columns by factors log-uniform in [1/cond, 1], so A D A' gets ill-conditioned early;
      * `n_dup` > 0 appends copies of the first rows (consistent right-hand sides): A loses full row rank. The normal
        equations are then singular (NormalKKTSystem needs full row rank, like the reference); K2 with LDL' and the
        dual regularization still solves the LP.
    Feasible and bounded by construction (b = A x*, c = A'y* + z*)."""
    rng = np.random.default_rng(seed)
    rows, cols = _sparsity(rng, m, n, k, "uniform", 0)
    vals = rng.standard_normal(len(rows))
    vals = np.sign(vals) * np.maximum(np.abs(vals), 1e-2)
    if cond > 1.0:
        vals = vals * np.exp(rng.uniform(-np.log(cond), 0.0, n))[cols]
    xs = np.zeros(n)
    pos = rng.permutation(n)[: max(1, m // 2)]
    xs[pos] = rng.uniform(0.5, 1.5, len(pos))
    zs = np.where(xs > 0, 0.0, rng.uniform(0.1, 1.0, n))
    zero = np.flatnonzero(xs == 0)
    zs[zero[rng.random(len(zero)) < 0.1]] = 0.0
    if n_dup > 0:
        extra = [(m + d, c_, v_) for d in range(n_dup) for c_, v_ in zip(cols[rows == d], vals[rows == d])]
        rows = np.concatenate([rows, np.array([e[0] for e in extra], dtype=rows.dtype)])
        cols = np.concatenate([cols, np.array([e[1] for e in extra], dtype=cols.dtype)])
        vals = np.concatenate([vals, np.array([e[2] for e in extra])])
    mm = m + n_dup
    b = np.zeros(mm)
    np.add.at(b, rows, vals * xs[cols])
    ys = rng.standard_normal(mm)
    c = zs.copy()
    np.add.at(c, cols, vals * ys[rows])
    return QuadraticModel(c=c, Hrows=[], Hcols=[], Hvals=[], Arows=rows, Acols=cols, Avals=vals, lcon=b, ucon=b.copy(),
                          lvar=np.zeros(n), uvar=np.full(n, np.inf), x0=np.zeros(n),
                          name=f"degenerate_lp_m{m}_n{n}_k{k}_s{seed}_dup{n_dup}_cond{cond:g}",
                          meta=dict(m=mm, n=n, k=k, seed=seed, n_dup=n_dup, cond=cond, objective=float(c @ xs)))


def bound_constrained_qp(n, seed):
    """Convex QP with bounds only (m = 0), like MadNLPTests.DenseDummyQP(x0; m=0) in test/runtests.jl:64."""
    rng = np.random.default_rng(seed)
    Q = rng.standard_normal((n, n))
    Q = Q @ Q.T + np.eye(n)
    rows, cols = np.tril_indices(n)
    return QuadraticModel(c=rng.standard_normal(n), Hrows=rows, Hcols=cols, Hvals=Q[rows, cols], Arows=[], Acols=[], Avals=[],
                          lcon=[], ucon=[], lvar=np.zeros(n), uvar=np.full(n, 2.0), x0=np.ones(n), name=f"boxqp_n{n}_s{seed}")


def block_angular_lp(K, grid_w, grid_h, n_link, seed):
    """Multicommodity-flow LP (config C4): K commodities on one directed grid graph
    (V = grid_w*grid_h nodes, 4-neighbour arcs in both directions). Per commodity N' x_k = b_k with N the
    node-arc incidence matrix minus one node row (the rows sum to zero and the normal path has no dual
    regularization); n_link linking rows sum_k x_k[e] + s_e = cap_e with explicit slack columns, placed LAST.
    Rows: [commodity 0 | ... | commodity K-1 | linking]; columns: [arcs of commodity 0 | ... | slacks]."""
    rng = np.random.default_rng(seed)
    V = grid_w * grid_h
    idx = np.arange(V).reshape(grid_h, grid_w)
    tail = np.concatenate([idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()])
    head = np.concatenate([idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()])
    E = len(tail)
    link = np.sort(rng.choice(E, size=min(n_link, E), replace=False))
    n_link = len(link)
    mrow = V - 1                                        # node V-1 dropped
    rows, cols, vals = [], [], []
    b = np.zeros(K * mrow + n_link)
    link_load = np.zeros(n_link)
    for kc in range(K):
        r0, c0 = kc * mrow, kc * E
        keep_t, keep_h = tail < mrow, head < mrow
        rows += [r0 + tail[keep_t], r0 + head[keep_h]]
        cols += [c0 + np.flatnonzero(keep_t), c0 + np.flatnonzero(keep_h)]
        vals += [np.ones(keep_t.sum()), -np.ones(keep_h.sum())]
        # feasible by construction: supplies are the divergence of a strictly positive flow x*_k, so they are
        # generic (a unit source/sink pair makes the optimal tree solution massively primal degenerate, and the
        # normal equations of a degenerate LP break down near the solution)
        xs = rng.uniform(0.5, 1.5, E)
        np.add.at(b, r0 + tail[keep_t], xs[keep_t])
        np.add.at(b, r0 + head[keep_h], -xs[keep_h])
        link_load += xs[link]
        rows.append(K * mrow + np.arange(n_link))
        cols.append(c0 + link)
        vals.append(np.ones(n_link))
    rows.append(K * mrow + np.arange(n_link))
    cols.append(K * E + np.arange(n_link))
    vals.append(np.ones(n_link))
    b[K * mrow:] = link_load + rng.uniform(0.1, 1.0, n_link)     # cap_e = sum_k x*_k[e] + s*_e
    n = K * E + n_link
    c = np.concatenate([rng.uniform(1.0, 2.0, K * E), np.zeros(n_link)])
    return QuadraticModel(c=c, Hrows=[], Hcols=[], Hvals=[], Arows=np.concatenate(rows), Acols=np.concatenate(cols),
                          Avals=np.concatenate(vals), lcon=b, ucon=b.copy(), lvar=np.zeros(n), uvar=np.full(n, np.inf),
                          x0=np.zeros(n), name=f"mcf_K{K}_{grid_w}x{grid_h}_L{n_link}_s{seed}",
                          meta=dict(K=K, V=V, E=E, n_link=n_link, n_border=n_link))


def config_c4(seed=4, scale=1.0):
    """m ~ 2e6: K = 64 commodities on a 176 x 176 grid (V = 30 976, E = 123 200), 2 048 linking rows."""
    K = max(2, int(round(64 * scale)))
    side = max(6, int(round(176 * np.sqrt(min(1.0, scale)))))
    return block_angular_lp(K, side, side, max(8, int(2048 * min(1.0, scale))), seed)
