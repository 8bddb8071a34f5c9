// Numeric supernodal multifrontal Cholesky / LDL^T, level-scheduled triangular solves and
// iterative refinement. Replaces cuDSS factorization / refactorization / solve as reached
// through MadNLP.factorize!(linear_solver) and MadNLP.solve!(linear_solver, x)
// (reference call sites: src/linear_solver.jl:10, src/KKT/normalkkt.jl:210).
//
// Data layout (all FP64, column-major):
//   panel of supernode s : (k+r) x k at L + lp[s], ld = k+r   (k columns, r rows below)
//   update matrix of s   : r x r     at U + up[s], ld = r      (lower triangle used)
//   inverse diagonal blocks: 64 x 64 at Dinv + 4096*(dinv_off[s] + jb/64)  (inv(L11 block), lower)
//   LDL^T only: W + wp[s], (k+r) x NB scratch holding the unscaled panel block L*D
//
// Execution model: ONE persistent cooperative kernel per factorization and ONE per solve.
// The schedule is a list of phases (extend-add | diagonal block | panel TRSM | trailing update
// for every (elimination-tree level, 64-column block step)); CTAs stride over the tasks of a
// phase and meet at a grid-wide barrier. A level-by-level multi-launch version of the same
// schedule spent most of its time in ~60 us launch/latency floors (profiles/launches_r01_baseline.csv).
// The dense work runs on the FP64 tensor pipe (mma.sync m8n8k4 -> DMMA): both the trailing
// update C -= X Y' and the panel solve X = R inv(L11)' are 64x64 tile GEMMs staged k-major in
// shared memory.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "diag_block.cuh"

namespace cg = cooperative_groups;

namespace mipm {

namespace {

constexpr int NB = 64;          // block width of the dense partial factorization
constexpr int LDS = 65;         // smem leading dimension for NB x NB blocks (odd: conflict-free rows)
constexpr int EA_COLS = 32;     // parent-front columns per extend-add task
constexpr int TILE = 64;        // GEMM tile
constexpr int XS = 68;          // smem row stride of staged operands (68 % 16 == 4: conflict-free DMMA loads)
constexpr int SMEM_DOUBLES = 2 * TILE * XS;
static_assert(2 * NB * LDS + 4 * NB <= SMEM_DOUBLES, "diag task: S + scratch + Sinv must fit the tile buffers");
static_assert(NB == mipm_diag::DB && LDS == mipm_diag::DLD, "diag_block.cuh is written for 64 x 64 blocks, ld 65");
constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;   // 69,632 B
enum { PH_EA = 0, PH_DIAG = 1, PH_TRSM = 2, PH_UPDATE = 3, PH_LEAF = 4, PH_FRONT = 5 };
constexpr int SL_K = 8;         // small leaf front: no children, at most SL_K columns ...
constexpr int SL_N = 32;        // ... and at most SL_N rows: one warp does the whole front in registers

// Everything a task needs to know about a front, in one 64-byte record (4 x 16-byte loads)
// instead of six dependent index lookups.
struct __align__(16) FrontInfo {
    int32_t k, r, c0, nchild;
    int64_t lp, up;
    int64_t wp, dinv;
    int64_t rowp, childp;
};

static_assert(sizeof(FrontInfo) == 64, "FrontInfo must match Handle::FI64");

struct FactorParams {
    const FrontInfo *fi;
    const int32_t *child_idx, *rel_idx;
    const int32_t *sched;
    const int64_t *phases;      // 8 x int64 per phase: type, jb, n_tasks, off_tasks, 0, 0, 0, 0
    int n_phases;               // phases [phase_begin, n_phases) are executed by this launch
    int phase_begin;
    double *L, *U, *W, *Dinv;
    int *info;
    const int32_t *ea_first;        // per front: first extend-add task record of its own (PH_FRONT levels), count
    const int32_t *ea_count;
    int *work_counter;              // one dynamic task counter per phase (PH_FRONT phases)
    unsigned long long *phase_ns;   // device-side time of every phase (one entry per phase)
    double piv_tol;
    const int *vmap;                // virtual CTA id per blockIdx.x (see build_cta_map), or null
    int *probe;                     // non-null: write %smid per block and return (setup-time placement probe)
};

struct SolveParams {
    const FrontInfo *fi;
    const int32_t *child_idx, *rel_idx, *row_idx, *perm;
    const int32_t *sched;
    int fwd_begin, fwd_end;     // forward sweep over levels [fwd_begin, fwd_end)
    int do_gather, do_backward; // stage control (distributed solves pause before the root level)
    int root_mode;              // front_forward mode for the last level (0 unless staged)
    const int64_t *lvl;         // 2 x int64 per level: off_all, n_all (level 0: regular fronts only)
    int64_t leaf_off;           // small leaf fronts (level 0), one warp each
    int n_leaf;
    int n_levels;
    int64_t n, n_u;
    const double *L, *Dinv;
    double *xp, *uvec;
    const double *b_in;         // gathered through perm at the start
    double *x_out;              // scattered through perm at the end
    int accumulate;             // x_out[perm] += xp instead of =
    const int *vmap;            // virtual CTA id per blockIdx.x, or null
    int *probe;                 // non-null: placement probe only
    int prefetch;               // L2 prefetch of the next level's fronts when it has at most this many (0 = off)
    // fronts with many children (K2: one tiny leaf child per primal variable) fold their children's update vectors in
    // through a transposed map: per destination row of the front, the update-vector slots that land on it, in child
    // order. gat_off[2 s] = offset of the front's N + 1 pointers in gat_ptr (-1: walk the children instead),
    // gat_off[2 s + 1] = offset of its source list in gat_src.
    const int64_t *gat_off;
    const int32_t *gat_ptr, *gat_src;
    unsigned long long *lvl_ns; // optional (MIPM_SOLVE_LOG): device time of gather, every forward level, every backward level
};

__device__ __forceinline__ int read_smid()
{
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return (int)s;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stage a (nrows x ncols) column-major block into smem k-major: dst[kk * XS + rr], zero padded to 64 x 64.
__device__ __forceinline__ void stage_tile(double *dst, const double *__restrict__ src, int64_t ld, int nrows, int ncols)
{
    double v[(TILE * TILE) / 256];          // all 16 global loads in flight before the first shared store
#pragma unroll
    for (int i = 0; i < (TILE * TILE) / 256; ++i) {
        int idx = threadIdx.x + i * 256;
        int rr = idx & 63, kk = idx >> 6;
        v[i] = (rr < nrows && kk < ncols) ? src[(int64_t)kk * ld + rr] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < (TILE * TILE) / 256; ++i) {
        int idx = threadIdx.x + i * 256;
        dst[(idx >> 6) * XS + (idx & 63)] = v[i];
    }
}

// Two operand tiles at once: 32 loads in flight per thread (one global round trip instead of two).
__device__ __forceinline__ void stage_two(double *dx, const double *__restrict__ sx, int64_t ldx, int nrx, int ncx,
                                          double *dy, const double *__restrict__ sy, int64_t ldy, int nry, int ncy)
{
    double v[(TILE * TILE) / 256], w[(TILE * TILE) / 256];
#pragma unroll
    for (int i = 0; i < (TILE * TILE) / 256; ++i) {
        int idx = threadIdx.x + i * 256;
        int rr = idx & 63, kk = idx >> 6;
        v[i] = (rr < nrx && kk < ncx) ? sx[(int64_t)kk * ldx + rr] : 0.0;
        w[i] = (rr < nry && kk < ncy) ? sy[(int64_t)kk * ldy + rr] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < (TILE * TILE) / 256; ++i) {
        int idx = threadIdx.x + i * 256;
        dx[(idx >> 6) * XS + (idx & 63)] = v[i];
        dy[(idx >> 6) * XS + (idx & 63)] = w[i];
    }
}

// acc(64x64) += Xs' * Ys over K: 8 warps as 2 (rows) x 4 (cols), each 32 x 16 = 4 x 2 DMMA fragments.
__device__ __forceinline__ void mma_64x64(double (&acc)[4][2][2], const double *Xs, const double *Ys, int K)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, l3 = lane & 3;
    const double *xa = Xs + l3 * XS + wm * 32 + g;
    const double *yb = Ys + l3 * XS + wn * 16 + g;
    for (int kk = 0; kk < K; kk += 4) {
        double a[4], b[2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = xa[kk * XS + mi * 8];
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) b[ni] = yb[kk * XS + ni * 8];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// Visit the 16 accumulator entries of this thread with their (row, col) inside the 64 x 64 tile.
template <typename F>
__device__ __forceinline__ void acc_foreach(const double (&acc)[4][2][2], F f)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, l3 = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) f(wm * 32 + mi * 8 + g, wn * 16 + ni * 8 + l3 * 2 + e, acc[mi][ni][e]);
}

// ------------------------------------------------------------------ tasks of the factorization
__global__ void __launch_bounds__(256)
k_scatter_a(int64_t nnz, const int64_t *__restrict__ a2l, const double *__restrict__ Ax, double *__restrict__ L)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) L[a2l[p]] = Ax[p];
}

// Extend-add: task = (parent s, parent-front columns [q0, q1)). The CTA walks the children of
// s in a fixed order and adds the part of each child's update matrix that lands in its column
// range, so every destination entry is owned by exactly one CTA -> deterministic sums.
__device__ void task_extend_add(const FactorParams &p, const int32_t *tasks, int task)
{
    // task record: (parent s, q0, q1, offset of the per-child [b0, b1) ranges computed on the host)
    const int4 t4 = *reinterpret_cast<const int4 *>(tasks + 4 * (int64_t)task);
    const int s = t4.x;
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r, N = f.k + f.r;
    double *P = p.L + f.lp;
    double *Us = p.U + f.up;
    // per task: the number of children that reach this column range, then (child position, b0, b1) for each of them
    // (a front of a K2 system can have thousands of tiny leaf children, almost all of them outside any one range)
    const int32_t *ranges = p.sched + t4.w;
    const int nrel = ranges[0];
    // Each warp owns a slice of the task's parent columns and walks ALL children for it, in order: destinations of
    // different warps are disjoint, so no CTA barrier separates the children (with hundreds of tiny leaf children per
    // task the barriers were most of the time) and the summation order per entry is still the child order.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cw = (t4.z - t4.y + 7) >> 3;
    const int lo = t4.y + warp * cw, hi = min(t4.z, lo + cw);
    if (lo >= hi) return;
    for (int e = 0; e < nrel; ++e) {
        const int ci = ranges[1 + 3 * e], b0 = ranges[2 + 3 * e], b1 = ranges[3 + 3 * e];   // b1 - b0 <= 32
        const int c = p.child_idx[f.childp + ci];
        const FrontInfo fc = p.fi[c];
        const int rc = fc.r;
        const int32_t *rel = p.rel_idx + fc.rowp;
        const double *Uc = p.U + fc.up;
        // rel is increasing: the child columns that land in [lo, hi) are a contiguous run of [b0, b1)
        const int bl = b0 + lane;
        const int tl = (bl < b1) ? rel[bl] : -1;
        const unsigned mine = __ballot_sync(0xffffffffu, tl >= lo && tl < hi);
        if (mine == 0u) continue;
        const int bfirst = b0 + __ffs(mine) - 1, bend = bfirst + __popc(mine);
        for (int b = bfirst; b < bend; ++b) {
            const int tb = __shfl_sync(0xffffffffu, tl, b - b0);
            const double *src = Uc + (int64_t)b * rc;
            double *dst = (tb < k) ? (P + (int64_t)tb * N) : (Us + (int64_t)(tb - k) * r - k);
            // rows of one child column land on distinct parent rows: four independent read-modify-writes in flight
            // per lane (the plain loop is a chain of dependent global round trips, the compiler cannot prove
            // the destinations distinct)
            int a = b + lane;
            for (; a + 96 < rc; a += 128) {
                const int r0 = rel[a], r1 = rel[a + 32], r2 = rel[a + 64], r3 = rel[a + 96];
                const double s0 = src[a], s1 = src[a + 32], s2 = src[a + 64], s3 = src[a + 96];
                const double d0 = dst[r0], d1 = dst[r1], d2 = dst[r2], d3 = dst[r3];
                dst[r0] = d0 + s0; dst[r1] = d1 + s1; dst[r2] = d2 + s2; dst[r3] = d3 + s3;
            }
            for (; a < rc; a += 32) dst[rel[a]] += src[a];
        }
    }
}

// Small leaf front (no children, k <= SL_K, k + r <= SL_N): one warp, lane i owns row i of the
// panel in registers; pivots and multipliers travel by shuffles; the update matrix is written
// directly (a leaf receives nothing, so U = -L21 D L21'). K2 systems have ~10^5 of these
// (one per primal variable), which the 64 x 64 tile machinery would handle at < 1% efficiency.
template <bool LDL>
__device__ void leaf_factor(const FactorParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r, N = f.k + f.r;
    const int lane = threadIdx.x & 31;
    double *P = p.L + f.lp;
    double *U = p.U + f.up;
    double pr[SL_K], dd[SL_K];
#pragma unroll
    for (int c = 0; c < SL_K; ++c) pr[c] = (lane < N && c < k && lane >= c) ? P[(int64_t)c * N + lane] : 0.0;
    int nbad = 0, ntiny = 0, nneg = 0;
#pragma unroll
    for (int j = 0; j < SL_K; ++j) {
        dd[j] = 1.0;
        if (j < k) {
            double d = __shfl_sync(0xffffffffu, pr[j], j);
            if (!LDL) {
                if (!(d > 0.0) || !(d < 1.0e300)) { nbad++; d = 1.0; }
            } else {
                if (!(fabs(d) <= 1.0e300)) { nbad++; d = 1.0; }
                else if (fabs(d) < p.piv_tol) { ntiny++; d = (d < 0.0) ? -p.piv_tol : p.piv_tol; }
                nneg += (d < 0.0);
            }
            dd[j] = d;
            const double colj = pr[j];                       // unscaled entry (lane, j)
            const double l = (lane > j) ? colj / d : 0.0;
#pragma unroll
            for (int c = j + 1; c < SL_K; ++c) {
                if (c < k) {
                    const double t = __shfl_sync(0xffffffffu, colj, c);   // unscaled entry (c, j)
                    if (lane >= c) pr[c] = fma(-l, t, pr[c]);
                }
            }
            pr[j] = (lane > j) ? l : ((lane == j) ? d : 0.0);
        }
    }
    if (!LDL) {                                              // L = L_unit * sqrt(D)
#pragma unroll
        for (int j = 0; j < SL_K; ++j) {
            const double sq = sqrt(dd[j]);
            pr[j] = (lane == j) ? sq : pr[j] * sq;
            dd[j] = 1.0;
        }
    }
#pragma unroll
    for (int c = 0; c < SL_K; ++c)
        if (c < k && lane < N && lane >= c) P[(int64_t)c * N + lane] = pr[c];
    // update matrix of the leaf: U(i, c) = -sum_j L(i, j) d_j L(c, j), k <= c <= i < N
    double ld[SL_K];
#pragma unroll
    for (int j = 0; j < SL_K; ++j) ld[j] = (j < k) ? pr[j] * dd[j] : 0.0;
    for (int c = k; c < N; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < SL_K; ++j) {
            const double lc = __shfl_sync(0xffffffffu, pr[j], c);
            acc = fma(ld[j], lc, acc);
        }
        if (lane >= c && lane < N) U[(int64_t)(c - k) * r + (lane - k)] = -acc;
    }
    if (lane == 0) {
        if (nbad) atomicMax(&p.info[0], 1);
        if (ntiny) atomicAdd(&p.info[2], ntiny);
        if (LDL && nneg) atomicAdd(&p.info[1], nneg);
    }
}

// Diagonal block: factor the nb x nb block at column jb and invert the triangular factor (diag_block.cuh:
// panel-blocked, one warp on the 16 x 16 serial part, everything else CTA-wide); both are written out.
// Cholesky: a non-positive pivot sets info[0] and is replaced by 1 so the run stays finite (the
// host then retries with more regularization like src/linear_solver.jl:6-17).
// LDL^T: unit-lower L11 with D on the diagonal; |pivot| < piv_tol is replaced by +-piv_tol.
template <bool LDL>
__device__ void task_diag(const FactorParams &p, int s, int jb, double *smem)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r;
    const int nb = min(NB, k - jb);
    double *P = p.L + f.lp + (int64_t)jb * N + jb;
    double *Dv = p.Dinv + (f.dinv + (jb >> 6)) * (int64_t)(NB * NB);
    mipm_diag::diag_block<LDL>(P, N, nb, Dv, p.piv_tol, p.info, smem);
}

// Panel TRSM as a tile GEMM with the inverted diagonal block: X = R * inv(L11)'  (64 rows per task).
template <bool LDL>
__device__ void task_trsm(const FactorParams &p, int s, int lc, int jb, double *smem)
{
    double *Xs = smem, *Ys = smem + TILE * XS;
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r;
    const int nb = min(NB, k - jb);
    const int row0 = jb + nb + lc * TILE;
    const int nrow = min(TILE, N - row0);
    double *P = p.L + f.lp;
    double *R = P + (int64_t)jb * N + row0;
    const double *Dv = p.Dinv + (f.dinv + (jb >> 6)) * (int64_t)(NB * NB);
    stage_tile(Xs, R, N, nrow, nb);
    stage_tile(Ys, Dv, NB, NB, nb);
    __syncthreads();
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    mma_64x64(acc, Xs, Ys, (nb + 3) & ~3);
    double *Ws = LDL ? (p.W + f.wp + row0) : nullptr;
    acc_foreach(acc, [&](int rr, int cc, double val) {
        if (rr < nrow && cc < nb) {
            if (!LDL) {
                R[(int64_t)cc * N + rr] = val;
            } else {
                Ws[(int64_t)cc * N + rr] = val;
                R[(int64_t)cc * N + rr] = val / P[(int64_t)(jb + cc) * N + jb + cc];
            }
        }
    });
}

// Trailing update on the FP64 tensor pipe: C(tile) -= X(rows, 0:nb) * Y(cols, 0:nb)'.
template <bool LDL>
__device__ void task_update(const FactorParams &p, int s, int lt, int jb, double *smem)
{
    double *Xs = smem, *Ys = smem + TILE * XS;
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r;
    const int N = k + r;
    const int nb = min(NB, k - jb);
    const int j1 = jb + nb;
    const int nt1 = (k > j1) ? (k - j1 + TILE - 1) / TILE : 0;
    int tr = (int)((sqrt(8.0 * (double)lt + 1.0) - 1.0) * 0.5);
    while (tr * (tr + 1) / 2 > lt) --tr;
    while ((tr + 1) * (tr + 2) / 2 <= lt) ++tr;
    const int tc = lt - tr * (tr + 1) / 2;
    int row0, rend, col0, cend;
    if (tr < nt1) { row0 = j1 + TILE * tr; rend = min(row0 + TILE, k); }
    else { row0 = k + TILE * (tr - nt1); rend = min(row0 + TILE, N); }
    if (tc < nt1) { col0 = j1 + TILE * tc; cend = min(col0 + TILE, k); }
    else { col0 = k + TILE * (tc - nt1); cend = min(col0 + TILE, N); }
    const int nrow = rend - row0, ncol = cend - col0;
    double *P = p.L + f.lp;
    const double *Y = P + (int64_t)jb * N + col0;
    const double *X = LDL ? (p.W + f.wp + row0) : (P + (int64_t)jb * N + row0);
    if (LDL || tr != tc) stage_two(Xs, X, N, nrow, nb, Ys, Y, N, ncol, nb);
    else { stage_tile(Xs, X, N, nrow, nb); Ys = Xs; }
    __syncthreads();
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    mma_64x64(acc, Xs, Ys, (nb + 3) & ~3);
    double *C;
    int64_t ldc;
    if (col0 < k) { C = P + (int64_t)col0 * N + row0; ldc = N; }
    else { C = p.U + f.up + (int64_t)(col0 - k) * r + (row0 - k); ldc = r; }
    // read all 16 destination entries first, then subtract and store: as read-modify-writes in sequence they are a
    // chain of dependent global round trips (the compiler must assume the stores alias the next load)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, l3 = lane & 3;
        double cv[4][2][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int rr = wm * 32 + mi * 8 + g, cc = wn * 16 + ni * 8 + l3 * 2 + e;
                    const bool on = rr < nrow && cc < ncol && (row0 + rr) >= (col0 + cc);
                    cv[mi][ni][e] = on ? C[(int64_t)cc * ldc + rr] : 0.0;
                }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int rr = wm * 32 + mi * 8 + g, cc = wn * 16 + ni * 8 + l3 * 2 + e;
                    const bool on = rr < nrow && cc < ncol && (row0 + rr) >= (col0 + cc);
                    if (on) C[(int64_t)cc * ldc + rr] = cv[mi][ni][e] - acc[mi][ni][e];
                }
    }
}

// Whole-front task for levels with many more fronts than CTAs: one CTA assembles the front from its
// children and runs every block step of its partial factorization with CTA-local barriers only. A level
// then is ONE phase instead of 1 + 3 x (block steps), no CTA idles at grid barriers while another
// front's serial diagonal block finishes, and the front stays hot in L2.
template <bool LDL>
__device__ void task_front(const FactorParams &p, int s, double *smem)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r, N = f.k + f.r;
    const int n_ea = p.ea_count[s];
    const int32_t *ea = p.sched + p.ea_first[s];
    for (int t = 0; t < n_ea; ++t) task_extend_add(p, ea, t);     // ends with a CTA barrier per child
    __syncthreads();
    for (int jb = 0; jb < k; jb += NB) {
        const int nb = min(NB, k - jb), j1 = jb + nb;
        task_diag<LDL>(p, s, jb, smem);
        __syncthreads();
        const int ntr = (N - j1 + TILE - 1) / TILE;
        for (int lc = 0; lc < ntr; ++lc) { task_trsm<LDL>(p, s, lc, jb, smem); __syncthreads(); }
        const int nt1 = (k > j1) ? (k - j1 + TILE - 1) / TILE : 0, nt2 = (r + TILE - 1) / TILE;
        const int nup = (nt1 + nt2) * (nt1 + nt2 + 1) / 2;
        for (int lt = 0; lt < nup; ++lt) { task_update<LDL>(p, s, lt, jb, smem); __syncthreads(); }
    }
}

template <bool LDL>
__global__ void __launch_bounds__(256, 3) k_factor_persistent(FactorParams p)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];
    unsigned long long t0 = 0;
    if (p.probe) { if (threadIdx.x == 0) p.probe[blockIdx.x] = read_smid(); return; }
    // Tasks are dealt to VIRTUAL CTA ids: ids 0..#SM-1 sit on distinct SMs, the next #SM ids are each SM's second
    // resident CTA, and so on, so a phase with few latency-bound tasks (diag) gets one SM per task. The hardware
    // places the first blockIdx values three to an SM (tools/smid_map.cu), which doubled those phases.
    const int vid = p.vmap ? p.vmap[blockIdx.x] : (int)blockIdx.x;
    const bool timer = (blockIdx.x == 0 && threadIdx.x == 0);
    if (timer) t0 = globaltimer_ns();
    for (int ph = p.phase_begin; ph < p.n_phases; ++ph) {
        const int64_t *d = p.phases + 8 * (int64_t)ph;
        const int type = (int)d[0], jb = (int)d[1], n_tasks = (int)d[2];
        const int32_t *A = p.sched + d[3];
        if (type == PH_FRONT) {
            // dynamic distribution (fronts are sorted by decreasing work on the host)
            __shared__ int next_task;
            for (;;) {
                if (threadIdx.x == 0) next_task = atomicAdd(&p.work_counter[ph], 1);
                __syncthreads();
                const int task = next_task;
                __syncthreads();
                if (task >= n_tasks) break;
                task_front<LDL>(p, A[task], smem);
                __syncthreads();
            }
        } else
        for (int task = vid; task < n_tasks; task += gridDim.x) {
            if (type == PH_LEAF) {
                const int li = task * 8 + (threadIdx.x >> 5);       // one warp per leaf front
                if (li < jb) leaf_factor<LDL>(p, A[li]);             // jb carries the number of leaves
            } else if (type == PH_EA) {
                task_extend_add(p, A, task);
            } else if (type == PH_DIAG) {
                task_diag<LDL>(p, A[task], jb, smem);
            } else {
                const int2 t2 = *reinterpret_cast<const int2 *>(A + 2 * (int64_t)task);   // (front, local tile)
                if (type == PH_TRSM) task_trsm<LDL>(p, t2.x, t2.y, jb, smem);
                else task_update<LDL>(p, t2.x, t2.y, jb, smem);
            }
            __syncthreads();
        }
        grid.sync();
        if (timer) {
            unsigned long long t1 = globaltimer_ns();
            p.phase_ns[ph] = t1 - t0;
            t0 = t1;
        }
    }
}

// Micro-benchmark hook for the same tile code: C (n x n, lower tiles) -= X X' with K = kdim.
__global__ void __launch_bounds__(256, 2)
k_bench_syrk(int n, int kdim, double *__restrict__ C, int64_t ldc, const double *__restrict__ X, int64_t ldx)
{
    extern __shared__ double smem[];
    double *Xs = smem, *Ys = smem + TILE * XS;
    const int lt = blockIdx.x;
    int tr = (int)((sqrt(8.0 * (double)lt + 1.0) - 1.0) * 0.5);
    while (tr * (tr + 1) / 2 > lt) --tr;
    while ((tr + 1) * (tr + 2) / 2 <= lt) ++tr;
    const int tc = lt - tr * (tr + 1) / 2;
    const int row0 = tr * TILE, col0 = tc * TILE;
    const int nrow = min(TILE, n - row0), ncol = min(TILE, n - col0);
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int k0 = 0; k0 < kdim; k0 += NB) {
        const int nb = min(NB, kdim - k0);
        stage_tile(Xs, X + (int64_t)k0 * ldx + row0, ldx, nrow, nb);
        stage_tile(Ys, X + (int64_t)k0 * ldx + col0, ldx, ncol, nb);
        __syncthreads();
        mma_64x64(acc, Xs, Ys, (nb + 3) & ~3);
        __syncthreads();
    }
    acc_foreach(acc, [&](int rr, int cc, double val) {
        if (rr < nrow && cc < ncol && (row0 + rr) >= (col0 + cc)) C[(int64_t)(col0 + cc) * ldc + row0 + rr] -= val;
    });
}

// ------------------------------------------------------------------ triangular solves
// One persistent cooperative kernel: gather through perm, forward sweep level by level (children's
// update vectors are summed by the parent in a fixed order), backward sweep from the root down,
// scatter through perm. One CTA per front per level; the 64 x 64 diagonal blocks are applied
// through their stored inverses (mat-vec), so nothing in a front is sequential.
constexpr int XR_MAX = 1536;    // ancestor entries of x cached in shared memory by the backward sweep

// L2 prefetch of a front's panel and inverted diagonal blocks. The solves are latency-bound: a front is a chain of
// dependent global loads, and the factor (hundreds of MB) does not stay in L2. CTAs with no task at a level fetch the
// next level's fronts when that level is small (near the root; prefetching a wide level only thrashes L2), so the
// dependent loads hit L2 instead of HBM.
constexpr int PREFETCH_MAX_FRONTS = 128;
__device__ __forceinline__ void prefetch_front(const SolveParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const char *base = reinterpret_cast<const char *>(p.L + f.lp);
    const int64_t bytes = (int64_t)(f.k + f.r) * f.k * 8;
    for (int64_t off = (int64_t)threadIdx.x * 128; off < bytes; off += 256 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    const char *dv = reinterpret_cast<const char *>(p.Dinv + f.dinv * (int64_t)(NB * NB));
    const int64_t bytes2 = (int64_t)((f.k + NB - 1) / NB) * NB * NB * 8;
    for (int64_t off = (int64_t)threadIdx.x * 128; off < bytes2; off += 256 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(dv + off));
}

// mode 0: whole front; 1: only fold the children's update vectors in (distributed solves: the root
// segment is all-reduced after this); 2: skip that part (it was done in the previous stage)
template <bool LDL>
__device__ void front_forward(const SolveParams &p, int s, double *smem, int mode)
{
    double *xb = smem, *yb = smem + NB;
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r;
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    double *u = p.uvec + f.rowp;
    const int tid = threadIdx.x;
    const int64_t go = p.gat_off[2 * (int64_t)s];
    if (go >= 0 && mode != 2) {
        const int32_t *gp = p.gat_ptr + go;
        const int32_t *gs = p.gat_src + p.gat_off[2 * (int64_t)s + 1];
        for (int t = tid; t < N; t += 256) {
            const int q0 = gp[t], q1 = gp[t + 1];
            if (q1 > q0) {
                double acc = 0.0;
                for (int q = q0; q < q1; ++q) acc += p.uvec[gs[q]];
                if (t < k) x1[t] += acc; else u[t - k] += acc;
            }
        }
        __syncthreads();
    }
    for (int ci = 0; ci < f.nchild && mode != 2 && go < 0; ++ci) {
        const int c = p.child_idx[f.childp + ci];
        const FrontInfo fc = p.fi[c];
        const int32_t *rel = p.rel_idx + fc.rowp;
        const double *uc = p.uvec + fc.rowp;
        for (int a = tid; a < fc.r; a += 256) {
            int t = rel[a];
            if (t < k) x1[t] += uc[a]; else u[t - k] += uc[a];
        }
        __syncthreads();
    }
    if (mode == 1) return;
    for (int jb = 0; jb < k; jb += NB) {
        const int nb = min(NB, k - jb);
        const double *Dv = p.Dinv + (f.dinv + (jb >> 6)) * (int64_t)(NB * NB);
        if (tid < NB) xb[tid] = (tid < nb) ? x1[jb + tid] : 0.0;
        __syncthreads();
        {   // y = inv(L11 block) * xb : row rr by the 4 threads (rr, q), columns pp = q, q+4, ...
            const int rr = tid >> 2, q = tid & 3;
            double acc = 0.0;
            double dvv[NB / 4];                  // the 16 loads of this row first (entries above the diagonal are stored zeros)
#pragma unroll
            for (int t = 0; t < NB / 4; ++t) dvv[t] = Dv[(q + 4 * t) * NB + rr];
#pragma unroll
            for (int t = 0; t < NB / 4; ++t) acc = fma(dvv[t], xb[q + 4 * t], acc);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (q == 0) { yb[rr] = acc; if (rr < nb) x1[jb + rr] = acc; }
        }
        __syncthreads();
        for (int i = jb + nb + tid; i < N; i += 256) {
            // 16 independent loads in flight per thread (the panel lives in HBM: latency-bound otherwise)
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const double *col = P + (int64_t)jb * N + i;
            int j = 0;
            for (; j + 16 <= nb; j += 16) {
                double v[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = col[(int64_t)(j + t) * N];
#pragma unroll
                for (int t = 0; t < 16; ++t) acc[t & 3] = fma(v[t], yb[j + t], acc[t & 3]);
            }
            for (; j < nb; ++j) acc[j & 3] = fma(col[(int64_t)j * N], yb[j], acc[j & 3]);
            const double tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
            if (i < k) x1[i] -= tot; else u[i - k] -= tot;
        }
        __syncthreads();
    }
}

template <bool LDL>
__device__ void front_backward(const SolveParams &p, int s, double *smem)
{
    double *S = smem;                 // inverse block, col-major ld LDS
    double *wb = smem + NB * LDS;     // 64
    double *xr = wb + NB;             // XR_MAX: x at the front's below-diagonal rows
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r, N = f.k + f.r;
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    const int32_t *rows = p.row_idx + f.rowp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cached = r <= XR_MAX;
    if (cached) for (int i = tid; i < r; i += 256) xr[i] = p.xp[rows[i]];
    const int nblk = (k + NB - 1) / NB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int jb = b * NB;
        const int nb = min(NB, k - jb);
        const double *Dv = p.Dinv + (f.dinv + b) * (int64_t)(NB * NB);
        for (int idx = tid; idx < NB * NB; idx += 256) S[(idx >> 6) * LDS + (idx & 63)] = Dv[idx];
        if (tid < NB) wb[tid] = 0.0;
        __syncthreads();
        // w[q] = y[q] (/ D[q]) - sum_{i >= jb+nb} L[i][q] * xfull[i]; each warp owns 8 consecutive
        // columns and walks the rows once for all of them (8 independent loads per lane in flight)
        {
            const int q0 = warp * 8;
            if (q0 < nb) {
                double acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = 0.0;
                const double *col = P + (int64_t)(jb + q0) * N;
                int i = jb + nb + lane;
                for (; i + 32 < N; i += 64) {     // two row-chunks per trip: 16 loads in flight per lane
                    const int i2 = i + 32;
                    double xv = (i < k) ? x1[i] : (cached ? xr[i - k] : p.xp[rows[i - k]]);
                    double xw = (i2 < k) ? x1[i2] : (cached ? xr[i2 - k] : p.xp[rows[i2 - k]]);
                    double v[8], w[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const bool on = q0 + c < nb;
                        v[c] = on ? col[(int64_t)c * N + i] : 0.0;
                        w[c] = on ? col[(int64_t)c * N + i2] : 0.0;
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] = fma(w[c], xw, fma(v[c], xv, acc[c]));
                }
                for (; i < N; i += 32) {
                    double xv = (i < k) ? x1[i] : (cached ? xr[i - k] : p.xp[rows[i - k]]);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (q0 + c < nb) acc[c] = fma(col[(int64_t)c * N + i], xv, acc[c]);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
                }
                if (lane < 8 && q0 + lane < nb) {
                    double a = 0.0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) if (lane == c) a = acc[c];
                    double y = x1[jb + q0 + lane];
                    if (LDL) y = y / col[(int64_t)lane * N + jb + q0 + lane];
                    wb[q0 + lane] = y - a;
                }
            }
        }
        __syncthreads();
        {   // x = inv(L11 block)' * w : column cc by the 4 threads (cc, q)
            const int cc = tid >> 2, q = tid & 3;
            double acc = 0.0;
            for (int rr = cc + q; rr < NB; rr += 4) acc = fma(S[cc * LDS + rr], wb[rr], acc);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (q == 0 && cc < nb) x1[jb + cc] = acc;
        }
        __syncthreads();
    }
}

// Small leaf fronts in the solves: one warp per front, L11 (k <= SL_K) applied by direct substitution.
template <bool LDL>
__device__ void leaf_forward(const SolveParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r;
    const int lane = threadIdx.x & 31;
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    double y[SL_K];
#pragma unroll
    for (int j = 0; j < SL_K; ++j) {
        y[j] = 0.0;
        if (j < k) {
            double t = x1[j];
#pragma unroll
            for (int q = 0; q < j; ++q) t = fma(-P[(int64_t)q * N + j], y[q], t);
            if (!LDL) t = t / P[(int64_t)j * N + j];
            y[j] = t;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < SL_K; ++j) if (j < k) x1[j] = y[j];
    }
    if (lane >= k && lane < N) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < SL_K; ++j) if (j < k) acc = fma(P[(int64_t)j * N + lane], y[j], acc);
        p.uvec[f.rowp + lane - k] = -acc;          // a leaf has no children: its update vector starts from zero
    }
}

template <bool LDL>
__device__ void leaf_backward(const SolveParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r;
    const int lane = threadIdx.x & 31;
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    const double xv = (lane >= k && lane < N) ? p.xp[p.row_idx[f.rowp + lane - k]] : 0.0;
    double w[SL_K];
#pragma unroll
    for (int j = 0; j < SL_K; ++j) {
        double part = (j < k && lane >= k && lane < N) ? P[(int64_t)j * N + lane] * xv : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        double yj = (j < k) ? x1[j] : 0.0;
        if (LDL && j < k) yj = yj / P[(int64_t)j * N + j];
        w[j] = yj - part;
    }
#pragma unroll
    for (int j = SL_K - 1; j >= 0; --j) {
        if (j < k) {
            double t = w[j];
#pragma unroll
            for (int q = j + 1; q < SL_K; ++q) if (q < k) t = fma(-P[(int64_t)j * N + q], w[q], t);
            if (!LDL) t = t / P[(int64_t)j * N + j];
            w[j] = t;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < SL_K; ++j) if (j < k) x1[j] = w[j];
    }
}

template <bool LDL>
__global__ void __launch_bounds__(256, 3) k_solve_persistent(SolveParams p)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double smem[NB * LDS + NB + XR_MAX];
    if (p.probe) { if (threadIdx.x == 0) p.probe[blockIdx.x] = read_smid(); return; }
    const int vid = p.vmap ? p.vmap[blockIdx.x] : (int)blockIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * 256 + threadIdx.x, gsz = (int64_t)gridDim.x * 256;
    if (p.do_gather) {
        for (int64_t i = gtid; i < p.n; i += gsz) p.xp[i] = p.b_in[p.perm[i]];
        for (int64_t i = gtid; i < p.n_u; i += gsz) p.uvec[i] = 0.0;
        grid.sync();
    }
    const bool timer = (p.lvl_ns != nullptr && blockIdx.x == 0 && threadIdx.x == 0);
    unsigned long long tprev = 0;
    int tslot = 0;
    if (timer) tprev = globaltimer_ns();
    const int32_t *leaves = p.sched + p.leaf_off;
    const int n_leaf_groups = (p.n_leaf + 7) / 8;
    for (int l = p.fwd_begin; l < p.fwd_end; ++l) {
        const int32_t *fr = p.sched + p.lvl[2 * l];
        const int nf = (int)p.lvl[2 * l + 1];
        const int extra = (l == 0) ? n_leaf_groups : 0;      // small leaves ride along with level 0
        if (p.prefetch && l + 1 < p.fwd_end && vid >= nf + extra && (int)p.lvl[2 * (l + 1) + 1] <= p.prefetch) {
            const int32_t *nx = p.sched + p.lvl[2 * (l + 1)];
            const int nnx = (int)p.lvl[2 * (l + 1) + 1], idle = (int)gridDim.x - (nf + extra);
            for (int j = vid - (nf + extra); j < nnx; j += idle) prefetch_front(p, nx[j]);
        }
        for (int t = vid; t < nf + extra; t += gridDim.x) {
            if (t < extra) {
                const int li = t * 8 + (threadIdx.x >> 5);
                if (li < p.n_leaf) leaf_forward<LDL>(p, leaves[li]);
            } else {
                front_forward<LDL>(p, fr[t - extra], smem, (l == p.n_levels - 1) ? p.root_mode : 0);
            }
            __syncthreads();
        }
        grid.sync();
        if (timer) { unsigned long long t1 = globaltimer_ns(); p.lvl_ns[tslot++] = t1 - tprev; tprev = t1; }
    }
    if (!p.do_backward) return;
    for (int l = p.n_levels - 1; l >= 0; --l) {
        const int32_t *fr = p.sched + p.lvl[2 * l];
        const int nf = (int)p.lvl[2 * l + 1];
        const int extra = (l == 0) ? n_leaf_groups : 0;
        if (p.prefetch && l > 0 && vid >= nf + extra && (int)p.lvl[2 * (l - 1) + 1] <= p.prefetch) {
            const int32_t *nx = p.sched + p.lvl[2 * (l - 1)];
            const int nnx = (int)p.lvl[2 * (l - 1) + 1], idle = (int)gridDim.x - (nf + extra);
            for (int j = vid - (nf + extra); j < nnx; j += idle) prefetch_front(p, nx[j]);
        }
        for (int t = vid; t < nf + extra; t += gridDim.x) {
            if (t < extra) {
                const int li = t * 8 + (threadIdx.x >> 5);
                if (li < p.n_leaf) leaf_backward<LDL>(p, leaves[li]);
            } else {
                front_backward<LDL>(p, fr[t - extra], smem);
            }
            __syncthreads();
        }
        grid.sync();
        if (timer) { unsigned long long t1 = globaltimer_ns(); p.lvl_ns[tslot++] = t1 - tprev; tprev = t1; }
    }
    if (p.accumulate) { for (int64_t i = gtid; i < p.n; i += gsz) p.x_out[p.perm[i]] += p.xp[i]; }
    else { for (int64_t i = gtid; i < p.n; i += gsz) p.x_out[p.perm[i]] = p.xp[i]; }
}

// r = b - K x with K symmetric, given by its full CSR index into the caller's lower-CSC values.
__global__ void __launch_bounds__(256)
k_sym_residual(int64_t n, const int64_t *__restrict__ ptr, const int32_t *__restrict__ col,
               const int64_t *__restrict__ vpos, const double *__restrict__ val, const double *__restrict__ x,
               const double *__restrict__ b, double *__restrict__ rout)
{
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    double acc = 0.0;
    for (int64_t q = ptr[row] + lane; q < ptr[row + 1]; q += 32) acc = fma(__ldg(val + vpos[q]), __ldg(x + col[q]), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rout[row] = b[row] - acc;
}

inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)std::max<int64_t>(1, (n + per_block - 1) / per_block); }

}  // namespace

// ---------------------------------------------------------------------------------------
// Placement probe: launch the persistent kernel in probe mode (same function, block size, shared memory and grid
// as the real launches, so the block scheduler places it the same way), read back the SM of every block and number
// the blocks so that ids [0, #SM) are the first resident CTA of each SM, [#SM, 2 #SM) the second, ...
// The map is a permutation of the block indices whatever the probe returns, so it can only affect speed.
static int build_cta_map(Handle *h, bool factor)
{
    const int grid = factor ? h->grid_factor : h->grid_solve;
    DBuf<int> &dmap = factor ? h->d_vmap_factor : h->d_vmap_solve;
    if (std::getenv("MIPM_NO_VMAP")) { dmap.release(); return MIPM_OK; }
    DBuf<int> d_probe;
    MIPM_CUDA(h, d_probe.alloc((size_t)grid));
    MIPM_CUDA(h, cudaMemsetAsync(d_probe.p, 0xff, (size_t)grid * sizeof(int), h->stream));
    const bool ldl = (h->sym.kind == MIPM_LDL);
    if (factor) {
        FactorParams p;
        std::memset(&p, 0, sizeof(p));
        p.probe = d_probe.p;
        void *args[] = {&p};
        const void *fn = ldl ? (const void *)k_factor_persistent<true> : (const void *)k_factor_persistent<false>;
        MIPM_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), args, SMEM_BYTES, h->stream));
    } else {
        SolveParams p;
        std::memset(&p, 0, sizeof(p));
        p.probe = d_probe.p;
        void *args[] = {&p};
        const void *fn = ldl ? (const void *)k_solve_persistent<true> : (const void *)k_solve_persistent<false>;
        MIPM_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), args, 0, h->stream));
    }
    std::vector<int> smid((size_t)grid);
    MIPM_CUDA(h, cudaMemcpyAsync(smid.data(), d_probe.p, (size_t)grid * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    // slot of each block on its SM (in block order), then sort by (slot, smid, block)
    std::vector<int> slot((size_t)grid), order((size_t)grid), vmap((size_t)grid);
    {
        std::vector<std::pair<int, int>> seen;      // (smid, count), tiny
        for (int b = 0; b < grid; ++b) {
            int c = -1;
            for (auto &e : seen) if (e.first == smid[(size_t)b]) { c = e.second++; break; }
            if (c < 0) { seen.push_back({smid[(size_t)b], 1}); c = 0; }
            slot[(size_t)b] = c;
        }
    }
    for (int b = 0; b < grid; ++b) order[(size_t)b] = b;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (slot[(size_t)a] != slot[(size_t)b]) return slot[(size_t)a] < slot[(size_t)b];
        if (smid[(size_t)a] != smid[(size_t)b]) return smid[(size_t)a] < smid[(size_t)b];
        return a < b;
    });
    for (int v = 0; v < grid; ++v) vmap[(size_t)order[(size_t)v]] = v;
    MIPM_CUDA(h, dmap.upload(vmap, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    return MIPM_OK;
}

int ls_device_setup(Handle *h)
{
    const LsSymbolic &S = h->sym;
    const int ns = S.ns;
    MIPM_CUDA(h, cudaSetDevice(h->device));
    // ---- build the schedule: per level [extend-add], then per block step diag / trsm / update.
    // Every phase owns a flat array of task records in `sched` (no searching on the device):
    //   DIAG: (front)   TRSM / UPDATE: (front, local tile)   EA: (parent, q0, q1, offset of child ranges)
    std::vector<int32_t> sched;
    std::vector<int64_t> phases, lvl;
    std::vector<int64_t> wp((size_t)ns + 1, 0), dinv_off((size_t)ns + 1, 0);
    std::vector<FrontInfo> finfo((size_t)std::max(ns, 1));
    std::vector<char> small((size_t)std::max(ns, 1), 0);
    const bool use_leaf = std::getenv("MIPM_NO_LEAF") == nullptr;
    for (int s = 0; s < ns; ++s) {
        int64_t k = S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s];
        int64_t r = S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s];
        small[(size_t)s] = use_leaf && S.child_ptr[(size_t)s + 1] == S.child_ptr[(size_t)s] && k <= SL_K && k + r <= SL_N;
        // small leaves need neither the LDL^T scratch panel nor inverted diagonal blocks
        wp[(size_t)s + 1] = wp[(size_t)s] + ((S.kind == MIPM_LDL && !small[(size_t)s]) ? (k + r) * NB : 0);
        dinv_off[(size_t)s + 1] = dinv_off[(size_t)s] + (small[(size_t)s] ? 0 : (k + NB - 1) / NB);
        FrontInfo &f = finfo[(size_t)s];
        f.k = (int32_t)k; f.r = (int32_t)r; f.c0 = S.sn_ptr[(size_t)s];
        f.nchild = (int32_t)(S.child_ptr[(size_t)s + 1] - S.child_ptr[(size_t)s]);
        f.lp = S.lp[(size_t)s]; f.up = S.up[(size_t)s]; f.wp = wp[(size_t)s]; f.dinv = dinv_off[(size_t)s];
        f.rowp = S.row_ptr[(size_t)s]; f.childp = S.child_ptr[(size_t)s];
    }
    auto align4 = [&]() { while (sched.size() % 4) sched.push_back(0); };
    auto push_phase = [&](int type, int jb, int64_t n_tasks, int64_t off_tasks) {
        int64_t d[8] = {type, jb, n_tasks, off_tasks, 0, 0, 0, 0};
        phases.insert(phases.end(), d, d + 8);
    };
    int64_t leaf_off = 0;
    int n_leaf = 0;
    int root_phase_begin = -1;
    // grid size is needed to decide which levels run whole-front tasks
    int fuse_min = 1 << 30;
    int grid_estimate = 444;
    {
        DeviceInfo prop0;
        if (device_info(h->device, prop0) != MIPM_OK) return fail(h, MIPM_ERR_CUDA, "cudaGetDeviceProperties failed");
        // Measured on B200 (tools/sweep_fuse.py): per-CTA tile latency, not the grid barriers, bounds
        // the wide levels, so whole-front tasks are neutral on C2 and slower on C3's small fronts.
        // Off by default; MIPM_FRONT_FUSE_MIN=<n> enables them for levels with at least n fronts.
        grid_estimate = prop0.sm_count * 3;       // __launch_bounds__(256, 3)
        if (const char *e = std::getenv("MIPM_FRONT_FUSE_MIN")) fuse_min = std::max(1, atoi(e));
    }
    std::vector<int32_t> ea_first((size_t)std::max(ns, 1), 0), ea_count((size_t)std::max(ns, 1), 0);
    for (int l = 0; l < S.n_levels; ++l) {
        const int64_t f0 = S.level_ptr[(size_t)l], f1 = S.level_ptr[(size_t)l + 1];
        if (l == 0) {            // small leaf fronts: their own list, one PH_LEAF phase, one warp each
            leaf_off = (int64_t)sched.size();
            for (int64_t t = f0; t < f1; ++t) {
                int s = S.level_sn[(size_t)t];
                if (small[(size_t)s]) { sched.push_back(s); n_leaf++; }
            }
            if (n_leaf) push_phase(PH_LEAF, n_leaf, (n_leaf + 7) / 8, leaf_off);
        }
        lvl.push_back((int64_t)sched.size());
        int64_t n_reg = 0;
        int kmax = 0;
        for (int64_t t = f0; t < f1; ++t) {
            int s = S.level_sn[(size_t)t];
            if (small[(size_t)s]) continue;
            sched.push_back(s);
            n_reg++;
            kmax = std::max(kmax, S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s]);
        }
        lvl.push_back(n_reg);
        const bool fuse_level = n_reg >= fuse_min && !(S.root_sn >= 0 && l == S.n_levels - 1);
        if (S.root_sn >= 0 && l == S.n_levels - 1) root_phase_begin = (int)(phases.size() / 8);   // moved past the EA phase below
        // extend-add: per-child column ranges first, then the task records that point at them
        std::vector<int32_t> ea;
        // columns per task: EA_COLS, narrower near the root where a level has fewer tasks than CTAs (an extend-add
        // task is a chain of dependent global round trips, so the level costs one task's latency)
        int ea_cols = EA_COLS;
        for (;;) {
            int64_t nt = 0;
            for (int64_t t = f0; t < f1; ++t) {
                const FrontInfo &f = finfo[(size_t)S.level_sn[(size_t)t]];
                if (f.nchild > 0) nt += (f.k + f.r + ea_cols / 2 - 1) / (ea_cols / 2);
            }
            if (ea_cols <= 8 || nt > grid_estimate) break;
            ea_cols /= 2;
        }
        for (int64_t t = f0; t < f1; ++t) {
            int s = S.level_sn[(size_t)t];
            const FrontInfo &f = finfo[(size_t)s];
            if (f.nchild == 0) continue;
            int N = f.k + f.r;
            ea_first[(size_t)s] = (int32_t)(ea.size() / 4);            // index inside this level's record array
            for (int q0 = 0; q0 < N; q0 += ea_cols) {
                int q1 = std::min(N, q0 + ea_cols);
                int64_t off_r = (int64_t)sched.size();
                sched.push_back(0);                                    // count, patched below
                int nrel = 0;
                for (int ci = 0; ci < f.nchild; ++ci) {
                    int c = S.child_idx[(size_t)(f.childp + ci)];
                    const int32_t *rel = S.rel_idx.data() + S.row_ptr[(size_t)c];
                    int rc = finfo[(size_t)c].r;
                    int b0 = (int)(std::lower_bound(rel, rel + rc, q0) - rel);
                    int b1 = (int)(std::lower_bound(rel, rel + rc, q1) - rel);
                    if (b1 > b0) {
                        sched.push_back(ci);
                        sched.push_back(b0);
                        sched.push_back(b1);
                        ++nrel;
                    }
                }
                if (nrel == 0) { sched.resize((size_t)off_r); continue; }
                sched[(size_t)off_r] = nrel;
                if (off_r > INT32_MAX) return fail(h, MIPM_ERR_ARG, "schedule too large");
                ea.push_back(s); ea.push_back(q0); ea.push_back(q1); ea.push_back((int32_t)off_r);
                ea_count[(size_t)s]++;
            }
        }
        if (!ea.empty()) {
            align4();
            const int64_t off_ea = (int64_t)sched.size();
            if (!fuse_level) push_phase(PH_EA, 0, (int64_t)ea.size() / 4, off_ea);
            if (S.root_sn >= 0 && l == S.n_levels - 1) root_phase_begin = (int)(phases.size() / 8);
            sched.insert(sched.end(), ea.begin(), ea.end());
            for (int64_t t = f0; t < f1; ++t) {       // turn per-level record indices into offsets into sched
                int s = S.level_sn[(size_t)t];
                if (ea_count[(size_t)s]) ea_first[(size_t)s] = (int32_t)(off_ea + 4 * (int64_t)ea_first[(size_t)s]);
            }
        }
        if (fuse_level) {
            std::vector<int32_t> fr;
            for (int64_t t = f0; t < f1; ++t) {
                int s = S.level_sn[(size_t)t];
                if (!small[(size_t)s]) fr.push_back(s);
            }
            std::stable_sort(fr.begin(), fr.end(), [&](int32_t a, int32_t b) {
                const FrontInfo &fa = finfo[(size_t)a], &fb = finfo[(size_t)b];
                double wa = (double)fa.k * (fa.k + fa.r) * (fa.k + fa.r), wb = (double)fb.k * (fb.k + fb.r) * (fb.k + fb.r);
                return wa > wb;
            });
            align4();
            push_phase(PH_FRONT, 0, (int64_t)fr.size(), (int64_t)sched.size());
            sched.insert(sched.end(), fr.begin(), fr.end());
            continue;
        }
        for (int jb = 0; jb < kmax; jb += NB) {
            std::vector<int32_t> td, tt, tu;
            for (int64_t t = f0; t < f1; ++t) {
                int s = S.level_sn[(size_t)t];
                const FrontInfo &f = finfo[(size_t)s];
                if (f.k <= jb || small[(size_t)s]) continue;
                td.push_back(s);
                int nb = std::min(NB, f.k - jb), j1 = jb + nb, N = f.k + f.r;
                int ntr = (N - j1 + TILE - 1) / TILE;
                for (int i = 0; i < ntr; ++i) { tt.push_back(s); tt.push_back(i); }
                int64_t nt1 = (f.k > j1) ? (f.k - j1 + TILE - 1) / TILE : 0, nt2 = (f.r + TILE - 1) / TILE;
                int64_t ntl = nt1 + nt2, nup = ntl * (ntl + 1) / 2;
                if (nup > (1 << 28)) return fail(h, MIPM_ERR_ARG, "front too large for the tile schedule");
                for (int i = 0; i < (int)nup; ++i) { tu.push_back(s); tu.push_back(i); }
            }
            if (sched.size() + td.size() + tt.size() + tu.size() + 16 > (size_t)INT32_MAX)
                return fail(h, MIPM_ERR_ARG, "schedule too large");
            align4();
            push_phase(PH_DIAG, jb, (int64_t)td.size(), (int64_t)sched.size());
            sched.insert(sched.end(), td.begin(), td.end());
            if (!tt.empty()) {
                align4();
                push_phase(PH_TRSM, jb, (int64_t)tt.size() / 2, (int64_t)sched.size());
                sched.insert(sched.end(), tt.begin(), tt.end());
            }
            if (!tu.empty()) {
                align4();
                push_phase(PH_UPDATE, jb, (int64_t)tu.size() / 2, (int64_t)sched.size());
                sched.insert(sched.end(), tu.begin(), tu.end());
            }
        }
    }
    h->leaf_off = leaf_off;
    h->n_leaf = n_leaf;
    h->root_phase_begin = root_phase_begin;
    h->n_phases = (int)(phases.size() / 8);
    h->n_launch_factor = 2;   // scatter + persistent kernel (plus three memsets)
    // ---- cooperative grid sizes
    DeviceInfo prop;
    if (device_info(h->device, prop) != MIPM_OK) return fail(h, MIPM_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (!prop.cooperative) return fail(h, MIPM_ERR_CUDA, "device does not support cooperative launch");
    MIPM_CUDA(h, cudaFuncSetAttribute(k_factor_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MIPM_CUDA(h, cudaFuncSetAttribute(k_factor_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MIPM_CUDA(h, cudaFuncSetAttribute(k_bench_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int occ_f = 0, occ_s = 0;
    if (S.kind == MIPM_LDL) {
        MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_factor_persistent<true>, 256, SMEM_BYTES));
        MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_solve_persistent<true>, 256, 0));
    } else {
        MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_factor_persistent<false>, 256, SMEM_BYTES));
        MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_solve_persistent<false>, 256, 0));
    }
    if (occ_f < 1 || occ_s < 1) return fail(h, MIPM_ERR_CUDA, "persistent kernels do not fit on an SM");
    h->grid_factor = prop.sm_count * occ_f;
    h->grid_solve = prop.sm_count * std::min(occ_s, 4);
    if (h->grid_limit > 0) {        // small systems solved side by side on one GPU (mipm_set_grid_limit)
        h->grid_factor = std::min(h->grid_factor, h->grid_limit);
        h->grid_solve = std::min(h->grid_solve, h->grid_limit);
    }
    // ---- uploads and workspaces
    cudaStream_t st = h->stream;
    MIPM_CUDA(h, h->d_sched.upload(sched, st));
    MIPM_CUDA(h, h->d_finfo.alloc(finfo.size()));
    MIPM_CUDA(h, cudaMemcpyAsync(h->d_finfo.p, finfo.data(), finfo.size() * sizeof(FrontInfo), cudaMemcpyHostToDevice, st));
    MIPM_CUDA(h, h->d_phases.upload(phases, st));
    MIPM_CUDA(h, h->d_ea_first.upload(ea_first, st));
    MIPM_CUDA(h, h->d_ea_count.upload(ea_count, st));
    MIPM_CUDA(h, h->d_work_counter.alloc((size_t)h->n_phases + 8));
    MIPM_CUDA(h, h->d_lvl.upload(lvl, st));
    MIPM_CUDA(h, h->d_sn_ptr.upload(S.sn_ptr, st));
    MIPM_CUDA(h, h->d_sn_parent.upload(S.sn_parent, st));
    MIPM_CUDA(h, h->d_row_ptr.upload(S.row_ptr, st));
    MIPM_CUDA(h, h->d_row_idx.upload(S.row_idx, st));
    MIPM_CUDA(h, h->d_rel_idx.upload(S.rel_idx, st));
    MIPM_CUDA(h, h->d_perm.upload(S.perm, st));
    MIPM_CUDA(h, h->d_lp.upload(S.lp, st));
    MIPM_CUDA(h, h->d_up.upload(S.up, st));
    MIPM_CUDA(h, h->d_wp.upload(wp, st));
    MIPM_CUDA(h, h->d_dinv_off.upload(dinv_off, st));
    MIPM_CUDA(h, h->d_child_ptr.upload(S.child_ptr, st));
    MIPM_CUDA(h, h->d_child_idx.upload(S.child_idx, st));
    MIPM_CUDA(h, h->d_a2l.upload(S.a2l, st));
    MIPM_CUDA(h, h->d_full_ptr.upload(S.full_ptr, st));
    MIPM_CUDA(h, h->d_full_col.upload(S.full_col, st));
    MIPM_CUDA(h, h->d_full_val.upload(S.full_val, st));
    MIPM_CUDA(h, h->d_L.alloc((size_t)std::max<int64_t>(S.nnz_l, 1)));
    h->L_cur = h->d_L.p;
    h->l_prezeroed = false;
    h->d_L2.release();
    if (S.root_sn < 0 && S.nnz_l > 0 && (size_t)S.nnz_l * sizeof(double) <= ((size_t)16 << 30) && !std::getenv("MIPM_SINGLE_L")) {
        if (h->d_L2.alloc((size_t)S.nnz_l) != cudaSuccess) { h->d_L2.release(); (void)cudaGetLastError(); }   // optional
    }
    MIPM_CUDA(h, h->d_U.alloc((size_t)std::max<int64_t>(S.update_doubles, 1)));
    MIPM_CUDA(h, h->d_W.alloc((size_t)std::max<int64_t>(wp[(size_t)ns], 1)));
    MIPM_CUDA(h, h->d_Dinv.alloc((size_t)std::max<int64_t>(dinv_off[(size_t)ns], 1) * NB * NB));
    MIPM_CUDA(h, h->d_xp.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_uvec.alloc((size_t)std::max<int64_t>(S.row_ptr[(size_t)ns], 1)));
    MIPM_CUDA(h, h->d_b.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_r.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_phase_ns.alloc((size_t)h->n_phases + 8));
    {
        // transposed child maps for the forward solve (fronts with more than GATHER_MIN_CHILDREN children)
        constexpr int GATHER_MIN_CHILDREN = 4;
        std::vector<int64_t> gat_off((size_t)2 * std::max(ns, 1), -1);
        std::vector<int32_t> gat_ptr, gat_src, cnt;
        const bool slots_fit = S.row_ptr[(size_t)ns] < (int64_t)INT32_MAX;
        for (int s = 0; s < ns && slots_fit; ++s) {
            const FrontInfo &f = finfo[(size_t)s];
            if (f.nchild <= GATHER_MIN_CHILDREN) continue;
            const int N = f.k + f.r;
            cnt.assign((size_t)N + 1, 0);
            for (int ci = 0; ci < f.nchild; ++ci) {
                const int c = S.child_idx[(size_t)(f.childp + ci)];
                const int32_t *rel = S.rel_idx.data() + S.row_ptr[(size_t)c];
                for (int a = 0; a < finfo[(size_t)c].r; ++a) cnt[(size_t)rel[a] + 1]++;
            }
            for (int t = 0; t < N; ++t) cnt[(size_t)t + 1] += cnt[(size_t)t];
            gat_off[(size_t)2 * s] = (int64_t)gat_ptr.size();
            gat_off[(size_t)2 * s + 1] = (int64_t)gat_src.size();
            gat_ptr.insert(gat_ptr.end(), cnt.begin(), cnt.end());
            const size_t base = gat_src.size();
            gat_src.resize(base + (size_t)cnt[(size_t)N]);
            for (int ci = 0; ci < f.nchild; ++ci) {             // child order, then row order: the summation order
                const int c = S.child_idx[(size_t)(f.childp + ci)];
                const int64_t rp = S.row_ptr[(size_t)c];
                const int32_t *rel = S.rel_idx.data() + rp;
                for (int a = 0; a < finfo[(size_t)c].r; ++a) gat_src[base + (size_t)cnt[(size_t)rel[a]]++] = (int32_t)(rp + a);
            }
        }
        if (gat_ptr.empty()) { gat_ptr.push_back(0); gat_src.push_back(0); }
        MIPM_CUDA(h, h->d_gat_off.upload(gat_off, st));
        MIPM_CUDA(h, h->d_gat_ptr.upload(gat_ptr, st));
        MIPM_CUDA(h, h->d_gat_src.upload(gat_src, st));
    }
    MIPM_CUDA(h, cudaStreamSynchronize(st));
    {
        int rc = build_cta_map(h, true);
        if (rc == MIPM_OK) rc = build_cta_map(h, false);
        if (rc != MIPM_OK) return rc;
    }
    if (!h->side) {
        MIPM_CUDA(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        MIPM_CUDA(h, cudaEventCreateWithFlags(&h->ev_factor_done, cudaEventDisableTiming));
        MIPM_CUDA(h, cudaEventCreateWithFlags(&h->ev_u_zero, cudaEventDisableTiming));
    } else {
        MIPM_CUDA(h, cudaStreamSynchronize(h->side));
    }
    h->u_prezeroed = false;
    h->factorized = false;
    return MIPM_OK;
}

// stage -1: everything; stage 0: assembly + all phases below the root (border) front, leaving the local
// Schur contribution in the root panel; stage 1: the root front's own factorization (after the all-reduce).
int ls_factorize_staged(Handle *h, const double *d_nzval, int stage)
{
    const LsSymbolic &S = h->sym;
    cudaStream_t st = h->stream;
    MIPM_CUDA(h, cudaSetDevice(h->device));
    if (stage >= 0 && h->root_phase_begin < 0) return fail(h, MIPM_ERR_STATE, "staged factorization needs mipm_ls_analyze_border");
    const int ph0 = (stage == 1) ? h->root_phase_begin : 0;
    const int ph1 = (stage == 0) ? h->root_phase_begin : h->n_phases;
    if (stage != 1) {
    if (h->d_L2.p) {
        // two factor buffers: this factorization writes the one the previous solves did not use; it was zero-filled on
        // the side stream while they ran
        h->L_cur = (h->L_cur == h->d_L.p) ? h->d_L2.p : h->d_L.p;
        if (h->l_prezeroed) { MIPM_CUDA(h, cudaStreamWaitEvent(st, h->ev_u_zero, 0)); }
        else { MIPM_CUDA(h, cudaMemsetAsync(h->L_cur, 0, (size_t)S.nnz_l * sizeof(double), st)); }
    } else {
        MIPM_CUDA(h, cudaMemsetAsync(h->d_L.p, 0, (size_t)std::max<int64_t>(S.nnz_l, 1) * sizeof(double), st));
    }
    // The update matrices are only live inside the factorization kernel, so their zero-fill for the
    // NEXT factorization runs on a side stream behind this one (it overlaps the latency-bound solves).
    if (h->u_prezeroed) {
        MIPM_CUDA(h, cudaStreamWaitEvent(st, h->ev_u_zero, 0));
    } else {
        MIPM_CUDA(h, cudaMemsetAsync(h->d_U.p, 0, (size_t)std::max<int64_t>(S.update_doubles, 1) * sizeof(double), st));
    }
    MIPM_CUDA(h, cudaMemsetAsync(h->d_info.p, 0, 4 * sizeof(int), st));
    MIPM_CUDA(h, cudaMemsetAsync(h->d_work_counter.p, 0, ((size_t)h->n_phases + 8) * sizeof(int), st));
    if (S.nnz_a > 0) {
        k_scatter_a<<<grid_for(S.nnz_a, 256), 256, 0, st>>>(S.nnz_a, h->d_a2l.p, d_nzval, h->L_cur);
        MIPM_CHECK_LAUNCH(h);
    }
    }
    if (ph1 > ph0) {
        FactorParams p;
        p.fi = (const FrontInfo *)h->d_finfo.p; p.child_idx = h->d_child_idx.p; p.rel_idx = h->d_rel_idx.p;
        p.sched = h->d_sched.p; p.phases = h->d_phases.p; p.n_phases = ph1; p.phase_begin = ph0;
        p.L = h->L_cur; p.U = h->d_U.p; p.W = h->d_W.p; p.Dinv = h->d_Dinv.p; p.info = h->d_info.p;
        p.phase_ns = h->d_phase_ns.p;
        p.ea_first = h->d_ea_first.p; p.ea_count = h->d_ea_count.p; p.work_counter = h->d_work_counter.p;
        p.piv_tol = 1e-13;   // LDL^T: absolute floor on |pivot|
        p.vmap = h->d_vmap_factor.p; p.probe = nullptr;
        void *args[] = {&p};
        const void *fn = (S.kind == MIPM_LDL) ? (const void *)k_factor_persistent<true> : (const void *)k_factor_persistent<false>;
        MIPM_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)h->grid_factor), dim3(256), args, SMEM_BYTES, st));
        h->launches++;
    }
    if (stage == 0) { h->d_nzval = d_nzval; return MIPM_OK; }
    if (h->side) {
        MIPM_CUDA(h, cudaEventRecord(h->ev_factor_done, st));
        MIPM_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_factor_done, 0));
        MIPM_CUDA(h, cudaMemsetAsync(h->d_U.p, 0, (size_t)std::max<int64_t>(S.update_doubles, 1) * sizeof(double), h->side));
        if (h->d_L2.p) {     // the buffer of the PREVIOUS factor: every solve that read it precedes ev_factor_done in stream order
            double *idle = (h->L_cur == h->d_L.p) ? h->d_L2.p : h->d_L.p;
            MIPM_CUDA(h, cudaMemsetAsync(idle, 0, (size_t)S.nnz_l * sizeof(double), h->side));
            h->l_prezeroed = true;
        }
        MIPM_CUDA(h, cudaEventRecord(h->ev_u_zero, h->side));
        h->u_prezeroed = true;
    }
    if (stage != 1) h->d_nzval = d_nzval;
    h->factorized = true;
    return MIPM_OK;
}

int ls_factorize_impl(Handle *h, const double *d_nzval) { return ls_factorize_staged(h, d_nzval, -1); }

// stage -1: whole solve; 0: gather + forward sweep below the root level; 1: root level forward, backward sweep, scatter
static int solve_once(Handle *h, const double *b_in, double *x_out, int accumulate, int stage = -1)
{
    const LsSymbolic &S = h->sym;
    SolveParams p;
    p.fi = (const FrontInfo *)h->d_finfo.p; p.child_idx = h->d_child_idx.p; p.rel_idx = h->d_rel_idx.p; p.row_idx = h->d_row_idx.p;
    p.perm = h->d_perm.p; p.sched = h->d_sched.p; p.lvl = h->d_lvl.p; p.n_levels = S.n_levels;
    p.leaf_off = h->leaf_off; p.n_leaf = h->n_leaf;
    p.n = S.n; p.n_u = S.row_ptr[(size_t)S.ns];
    p.L = h->L_cur; p.Dinv = h->d_Dinv.p; p.xp = h->d_xp.p; p.uvec = h->d_uvec.p;
    p.b_in = b_in; p.x_out = x_out; p.accumulate = accumulate;
    p.vmap = h->d_vmap_solve.p; p.probe = nullptr;
    p.gat_off = h->d_gat_off.p; p.gat_ptr = h->d_gat_ptr.p; p.gat_src = h->d_gat_src.p;
    static const bool solve_log = std::getenv("MIPM_SOLVE_LOG") != nullptr;
    static const int prefetch_max = std::getenv("MIPM_NO_PREFETCH") ? 0
                                    : (std::getenv("MIPM_PREFETCH_MAX") ? atoi(std::getenv("MIPM_PREFETCH_MAX")) : PREFETCH_MAX_FRONTS);
    p.prefetch = prefetch_max;
    DBuf<unsigned long long> d_lvl_ns;
    p.lvl_ns = nullptr;
    if (solve_log) {
        MIPM_CUDA(h, d_lvl_ns.alloc((size_t)2 * S.n_levels + 2));
        MIPM_CUDA(h, cudaMemsetAsync(d_lvl_ns.p, 0, ((size_t)2 * S.n_levels + 2) * sizeof(unsigned long long), h->stream));
        p.lvl_ns = d_lvl_ns.p;
    }
    p.do_gather = (stage != 1);
    p.fwd_begin = (stage == 1) ? S.n_levels - 1 : 0;
    p.fwd_end = S.n_levels;
    p.root_mode = (stage == 0) ? 1 : ((stage == 1) ? 2 : 0);
    p.do_backward = (stage != 0);
    void *args[] = {&p};
    const void *fn = (S.kind == MIPM_LDL) ? (const void *)k_solve_persistent<true> : (const void *)k_solve_persistent<false>;
    MIPM_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)h->grid_solve), dim3(256), args, 0, h->stream));
    h->launches++;
    if (solve_log) {        // diagnostic only: synchronises
        std::vector<unsigned long long> ns((size_t)2 * S.n_levels + 2);
        MIPM_CUDA(h, cudaMemcpyAsync(ns.data(), d_lvl_ns.p, ns.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
        std::fprintf(stderr, "solve levels (us): fwd");
        int slot = 0;
        for (int l = p.fwd_begin; l < p.fwd_end; ++l) std::fprintf(stderr, " %d:%lld[%lld]", l, (long long)(ns[(size_t)slot++] / 1000), (long long)S.level_ptr[(size_t)l + 1] - (long long)S.level_ptr[(size_t)l]);
        if (p.do_backward) {
            std::fprintf(stderr, " | bwd");
            for (int l = S.n_levels - 1; l >= 0; --l) std::fprintf(stderr, " %d:%lld", l, (long long)(ns[(size_t)slot++] / 1000));
        }
        std::fprintf(stderr, "\n");
    }
    return MIPM_OK;
}

int ls_solve_impl(Handle *h, double *d_x, int ir_steps)
{
    const LsSymbolic &S = h->sym;
    cudaStream_t st = h->stream;
    MIPM_CUDA(h, cudaSetDevice(h->device));
    const int64_t n = S.n;
    if (n == 0) return MIPM_OK;
    // b is needed after x is overwritten (refinement) and the solve reads b through perm while
    // writing x through perm: always work from a copy of the right-hand side
    MIPM_CUDA(h, cudaMemcpyAsync(h->d_b.p, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    int rc = solve_once(h, h->d_b.p, d_x, 0);
    if (rc != MIPM_OK) return rc;
    for (int it = 0; it < ir_steps; ++it) {
        k_sym_residual<<<grid_for(n * 32, 256), 256, 0, st>>>(n, h->d_full_ptr.p, h->d_full_col.p, h->d_full_val.p, h->d_nzval,
                                                            d_x, h->d_b.p, h->d_r.p);
        MIPM_CHECK_LAUNCH(h);
        rc = solve_once(h, h->d_r.p, d_x, 1);
        if (rc != MIPM_OK) return rc;
    }
    return MIPM_OK;
}

int ls_solve_staged(Handle *h, double *d_x, int stage)
{
    const LsSymbolic &S = h->sym;
    MIPM_CUDA(h, cudaSetDevice(h->device));
    if (h->root_phase_begin < 0) return fail(h, MIPM_ERR_STATE, "staged solve needs mipm_ls_analyze_border");
    if (S.n == 0) return MIPM_OK;
    if (stage == 0) {
        MIPM_CUDA(h, cudaMemcpyAsync(h->d_b.p, d_x, (size_t)S.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return solve_once(h, h->d_b.p, d_x, 0, 0);
    }
    return solve_once(h, h->d_b.p, d_x, 0, 1);
}

}  // namespace mipm

extern "C" int mipm_ls_analyze_border(mipm_handle hh, int64_t n, const int32_t *colptr, const int32_t *rowval, int index_base,
                                      int kind, int64_t n_border)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    if (h && !h->host_only) use_handle(h);
    if (!h || n < 0 || !colptr || n_border < 1 || n_border > n || (kind != MIPM_CHOLESKY && kind != MIPM_LDL))
        return fail(h, MIPM_ERR_ARG, "bad argument");
    int64_t nnz = colptr[n] - index_base;
    std::vector<int32_t> cp((size_t)n + 1), ri((size_t)std::max<int64_t>(nnz, 0));
    for (int64_t j = 0; j <= n; ++j) cp[(size_t)j] = colptr[j] - index_base;
    for (int64_t q = 0; q < nnz; ++q) ri[(size_t)q] = rowval[q] - index_base;
    LsOptions opt;
    opt.kind = kind;
    opt.ordering = MIPM_ORDER_ND;
    opt.n_border = n_border;
    if (const char *s = std::getenv("MIPM_ND_LEAF")) opt.nd_leaf = std::max(1, atoi(s));
    h->has_ls = false;
    h->factorized = false;
    std::string e = ls_analyze(n, cp.data(), ri.data(), opt, nullptr, h->sym);
    if (!e.empty()) return fail(h, MIPM_ERR_ARG, e);
    h->has_ls = true;
    if (!h->host_only) return ls_device_setup(h);
    return MIPM_OK;
}

extern "C" int mipm_ls_factorize_stage(mipm_handle hh, const double *d_nzval, int stage)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze_border has not been called");
    if (stage != 0 && stage != 1) return fail(h, MIPM_ERR_ARG, "stage must be 0 or 1");
    if (stage == 0 && !d_nzval && h->sym.nnz_a > 0) return fail(h, MIPM_ERR_ARG, "null values");
    return ls_factorize_staged(h, d_nzval, stage);
}

extern "C" int mipm_ls_solve_stage(mipm_handle hh, double *d_x, int stage)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls || !h->factorized) return fail(h, MIPM_ERR_STATE, "solve before factorize");
    if (stage != 0 && stage != 1) return fail(h, MIPM_ERR_ARG, "stage must be 0 or 1");
    if (!d_x && h->sym.n > 0) return fail(h, MIPM_ERR_ARG, "null argument");
    return ls_solve_staged(h, d_x, stage);
}

extern "C" int mipm_ls_root_info(mipm_handle hh, double **d_root_panel, int64_t *n_root, double **d_root_rhs)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls || h->sym.root_sn < 0) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze_border has not been called");
    const LsSymbolic &S = h->sym;
    const int rs = S.root_sn;
    if (d_root_panel) *d_root_panel = h->d_L.p + S.lp[(size_t)rs];
    if (n_root) *n_root = S.sn_ptr[(size_t)rs + 1] - S.sn_ptr[(size_t)rs];
    if (d_root_rhs) *d_root_rhs = h->d_xp.p + S.sn_ptr[(size_t)rs];
    return MIPM_OK;
}

extern "C" int mipm_ls_factorize_profile(mipm_handle hh, const double *d_nzval, double *ms, double *work, int64_t *launches)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze has not been called");
    if (!ms || !work || !launches) return fail(h, MIPM_ERR_ARG, "null argument");
    const LsSymbolic &S = h->sym;
    cudaEvent_t e0, e1;
    MIPM_CUDA(h, cudaEventCreate(&e0));
    MIPM_CUDA(h, cudaEventCreate(&e1));
    MIPM_CUDA(h, cudaEventRecord(e0, h->stream));
    int rc = ls_factorize_impl(h, d_nzval);
    if (rc != MIPM_OK) return rc;
    MIPM_CUDA(h, cudaEventRecord(e1, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    float total = 0.f;
    cudaEventElapsedTime(&total, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::vector<unsigned long long> ns((size_t)h->n_phases + 8, 0);
    std::vector<int64_t> ph((size_t)h->n_phases * 8 + 8, 0);
    if (h->n_phases) {
        MIPM_CUDA(h, cudaMemcpy(ns.data(), h->d_phase_ns.p, (size_t)h->n_phases * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        MIPM_CUDA(h, cudaMemcpy(ph.data(), h->d_phases.p, (size_t)h->n_phases * 8 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    // classes: 0 zero-fill + scatter (= total - in-kernel phases), 1 extend-add, 2 diag, 3 trsm, 4 update
    for (int c = 0; c < 5; ++c) { ms[c] = 0.0; work[c] = 0.0; launches[c] = 0; }
    double inside = 0.0;
    FILE *logf = nullptr;
    if (const char *pth = std::getenv("MIPM_PHASE_LOG")) logf = std::fopen(pth, "w");
    if (logf) std::fprintf(logf, "phase,type,jb,n_tasks,us\n");
    for (int i = 0; i < h->n_phases; ++i) {
        int type = (int)ph[(size_t)i * 8];
        double t = (double)ns[(size_t)i] * 1e-6;
        // small-leaf fronts are counted with the diagonal-block class, whole-front phases with the update class
        const int cls = (type == PH_LEAF) ? 2 : ((type == PH_FRONT) ? 4 : type + 1);
        ms[cls] += t;
        launches[cls] += 1;
        inside += t;
        if (logf) std::fprintf(logf, "%d,%d,%d,%lld,%.2f\n", i, type, (int)ph[(size_t)i * 8 + 1], (long long)ph[(size_t)i * 8 + 2], t * 1e3);
    }
    if (logf) std::fclose(logf);
    ms[0] = std::max(0.0, (double)total - inside);
    launches[0] = 4;
    work[0] = 8.0 * (double)(S.nnz_l + S.update_doubles) + 24.0 * (double)S.nnz_a;
    for (int s = 0; s < S.ns; ++s) {
        double k = S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s];
        double r = (double)(S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s]);
        if (S.sn_parent[(size_t)s] >= 0) work[1] += 24.0 * r * (r + 1) / 2 + 4.0 * r;
        for (double jb = 0; jb < k; jb += NB) {
            double nb = std::min<double>(NB, k - jb), T = k + r - jb - nb;
            work[2] += nb * nb * nb / 3.0;
            work[3] += nb * nb * T;
            work[4] += nb * T * (T + 1);
        }
    }
    return MIPM_OK;
}

extern "C" int mipm_bench_syrk(mipm_handle hh, int64_t n, int64_t k, double *d_C, int64_t ldc, const double *d_X, int64_t ldx)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n <= 0 || k <= 0 || !d_C || !d_X || ldc < n || ldx < n) return fail(h, MIPM_ERR_ARG, "bad argument");
    int64_t nt = (n + TILE - 1) / TILE;
    int64_t tiles = nt * (nt + 1) / 2;
    if (tiles > INT32_MAX) return fail(h, MIPM_ERR_ARG, "too many tiles");
    MIPM_CUDA(h, cudaFuncSetAttribute(k_bench_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    k_bench_syrk<<<(unsigned)tiles, 256, SMEM_BYTES, h->stream>>>((int)n, (int)k, d_C, ldc, d_X, ldx);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}
