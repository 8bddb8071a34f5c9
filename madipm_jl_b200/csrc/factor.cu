// Numeric supernodal multifrontal Cholesky / LDL^T. Replaces cuDSS factorization / refactorization as reached through
// MadNLP.factorize!(linear_solver) (reference call site: src/linear_solver.jl:10). The triangular solves are in solve.cu.
//
// Execution model: ONE persistent launch per factorization that executes a TASK GRAPH. The host analysis flattens the
// factorization into a topologically ordered list of tasks (front = supernode of the elimination tree):
//     EA     extend-add of the children's update matrices into a slice of the front's columns
//     DIAG   64-column diagonal block: left-looking update with the finished columns of the current super-panel,
//            factorization, explicit inverse of the triangular factor
//     PANEL  64 rows x 64 columns of the panel below a diagonal block: left-looking update, then X inv(L11)'
//     TRAIL  64 x 64 tile of the trailing matrix (remaining panel columns or the update matrix U):
//            C -= L[rows, K] D L[cols, K]' accumulated over a whole super-panel K (up to 256+ columns) in registers
//     LEAF   eight small leaf fronts, one warp each, entirely in registers
// CTAs draw tasks in list order from an atomic ticket counter; a task waits (acquire spin) until a per-front progress
// counter reaches the number of earlier completions it depends on, and bumps the counter (release) when done; the task
// that completes a front bumps its parent's counter. Because every task only depends on tasks earlier in the list and
// every CTA of the (cooperative) launch is resident, this cannot deadlock, and nothing ever waits for the whole grid:
// the round-1 kernel ran the same work as 116 phases separated by grid-wide barriers and re-read / re-wrote each
// trailing matrix once per 64 columns (profiles/factor_phases_r01_final.csv).
//
// Dense work runs on the FP64 tensor pipe (mma.sync m8n8k4 -> DMMA.8x8x4; tcgen05 has no FP64 kind). Operand tiles are
// moved global -> shared by TMA bulk copies (cp.async.bulk + mbarrier complete_tx, 16 columns per stage, 4 stages),
// so the K loop overlaps copies and math without staging through registers.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "diag_block.cuh"
#include "front.cuh"

namespace mipm {

namespace {

constexpr int EA_COLS = 32;                     // parent-front columns per extend-add task
constexpr int KC = 16;                          // columns per pipeline stage
constexpr int NSTAGE = 4;
constexpr int STAGE_DOUBLES = 2 * KC * XS;      // A and B operand of one stage
constexpr int DIAG_DOUBLES = 2 * NB * mipm_diag::DLD + 4 * NB;      // diag task: S + scratch + Sinv
constexpr int SMEM_DOUBLES = (NSTAGE * STAGE_DOUBLES > DIAG_DOUBLES) ? NSTAGE * STAGE_DOUBLES : DIAG_DOUBLES;   // 8960 doubles = 71,680 B
constexpr int SMEM_BYTES = SMEM_DOUBLES * 8 + NSTAGE * KC * 8 + 64;   // + pivots of each stage (LDL^T) + mbarrier
static_assert(2 * TILE * XS <= SMEM_DOUBLES, "TRSM epilogue: X and inv(L11) tiles must fit the stage buffers");
static_assert(NB == mipm_diag::DB, "diag_block.cuh is written for 64 x 64 blocks");
enum { T_EA = 0, T_DIAG = 1, T_PANEL = 2, T_TRAIL = 3, T_LEAF = 4, T_NCLASS = 5 };

#ifndef MIPM_FACTOR_OCC
#define MIPM_FACTOR_OCC 3
#endif
constexpr int FACTOR_OCC = MIPM_FACTOR_OCC;     // CTAs of the task kernel per SM (register budget 65536 / 256 / FACTOR_OCC)

struct __align__(16) Task {
    int32_t type, front;
    int32_t a, b, c, d;         // EA: q0, q1, offset of the child ranges | DIAG: jb, K0 | PANEL: jb, K0, row0 | TRAIL: K0, K1, row0, col0
    int32_t need;               // completions on the front's chain counter (children + own chain tasks) this task waits for
    int32_t need_b;             // look-ahead fronts: completions on the second counter (deferred trailing tiles) it waits for
};
constexpr int T_DEFERRED = 0x100;   // type flag: the task counts on the second counter
static_assert(sizeof(Task) == 32, "Task must match Handle::T32");

struct FactorParams {
    const FrontInfo *fi;
    const int32_t *child_idx, *rel_idx;
    const int32_t *sched;           // extend-add child ranges, small-leaf list
    const Task *tasks;
    int task_begin, task_end;       // this launch executes tasks [task_begin, task_end)
    double *L, *U, *Dinv, *Dg;
    int *info;
    int *prog;                      // per front: completions so far on the chain counter (children included)
    int *prog_b, *done;             // per front: completions of deferred trailing tiles; of all own tasks
    int *ticket;
    unsigned long long *prof;       // 8 per CTA: busy ns per task class, wait ns, first start, last end
    unsigned long long *front_ns;   // optional: completion time of every front
    unsigned long long *trace;      // optional: 4 per task (wait start, start, end, SM id)
    double piv_tol;
};

// ------------------------------------------------------------------ PTX helpers (mbarrier, TMA bulk copy, acquire load)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA bulk copy global -> shared (SASS: UBLKCP); src, dst and bytes must be multiples of 16
__device__ __forceinline__ void bulk_g2s(double *dst, const double *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Operand pipeline of a CTA: NSTAGE stages of KC columns, filled by 16-byte cp.async (LDGSTS) from every thread and
// consumed one stage per iteration with a single CTA barrier (per-column TMA bulk copies of 512 B were tried first:
// the copy engine then needs ~1.5 us per stage, profiles/r02_notes.md). ephase: parity of the epilogue's mbarrier.
struct Pipe {
    double *buf;        // NSTAGE x (A[KC][XS], B[KC][XS])
    double *dsm;        // NSTAGE x KC pivots (LDL^T)
    uint64_t *bar;      // mbarrier of the epilogue's bulk copy (inverse diagonal block)
    uint32_t ephase;
};

__device__ __forceinline__ void cp_async16(double *dst, const double *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One operand of one stage: columns [kcol, kcol + nkc) x (32 or 33) 16-byte segments from the even row `src` points at;
// the columns up to nk4 (K tail) are zero-filled by the same instruction (source size 0).
__device__ __forceinline__ void issue_operand(double *sdst, const double *src, int64_t ld, int kcol, int nkc, int nk4, int segs)
{
    const int total = nk4 * segs;
    for (int i = threadIdx.x; i < total; i += 256) {
        const int col = i / segs, seg = i - col * segs;
        const bool on = col < nkc;
        cp_async16(sdst + col * XS + 2 * seg, src + (int64_t)(kcol + (on ? col : 0)) * ld + 2 * seg, on ? 16 : 0);
    }
}

template <bool LDL>
__device__ __forceinline__ void issue_chunk(const Pipe &pp, int stage, const double *A, const double *B, bool same, int64_t ld,
                                            int kcol, int nkc, int segA, int segB, const double *dv)
{
    double *sa = pp.buf + stage * STAGE_DOUBLES, *sb = sa + KC * XS;
    const int nk4 = (nkc + 3) & ~3;
    issue_operand(sa, A, ld, kcol, nkc, nk4, segA);
    if (!same) issue_operand(sb, B, ld, kcol, nkc, nk4, segB);
    if (LDL) {
        const int j = threadIdx.x;
        if (j < KC) {
            double d = 0.0;
            if (j < nkc) d = dv[kcol + j];
            pp.dsm[stage * KC + j] = d;
        }
    }
}

// acc(64x64) += A' B over nk4 (multiple of 4) staged columns: 8 warps as 2 (rows) x 4 (cols), each 32 x 16 = 4 x 2 DMMA
// fragments; fragment (mi, ni) is skipped when bit mi*2+ni of fmask is clear (outside the tile or above the diagonal).
template <bool SCALE>
__device__ __forceinline__ void mma_chunk(double (&acc)[4][2][2], const double *sa, const double *sb, int nk4, const double *dsc,
                                          uint32_t fmask)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, l3 = lane & 3;
    const double *xa = sa + l3 * XS + wm * 32 + g;
    const double *yb = sb + l3 * XS + wn * 16 + g;
    if (fmask == 0u) return;
    auto step = [&](int kk) {
        double a[4], b[2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = xa[kk * XS + mi * 8];
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) b[ni] = yb[kk * XS + ni * 8];
        if (SCALE) {
            const double d = dsc[kk + l3];
            b[0] *= d;
            b[1] *= d;
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
                if (fmask & (1u << (mi * 2 + ni))) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    };
    if (SCALE) {
#pragma unroll 1       // (ptxas 12.9 crashes on the LDL^T instantiation with this loop unrolled)
        for (int kk = 0; kk < nk4; kk += 4) step(kk);
    } else {
#pragma unroll 4
        for (int kk = 0; kk < nk4; kk += 4) step(kk);
    }
}

// Fragments of this warp that intersect [0, nrow) x [0, ncol) and, on a diagonal tile, the lower triangle.
__device__ __forceinline__ uint32_t frag_mask(int nrow, int ncol, bool diag)
{
    const int warp = threadIdx.x >> 5, wm = warp >> 2, wn = warp & 3;
    uint32_t m = 0;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const int r0 = wm * 32 + mi * 8, c0 = wn * 16 + ni * 8;
            if (r0 < nrow && c0 < ncol && (!diag || r0 + 7 >= c0)) m |= 1u << (mi * 2 + ni);
        }
    return m;
}

// acc += A[rows, K0:K1] (D) B[cols, K0:K1]' with the operands streamed through the stage ring. A / B point at the even
// row at or before the tile's first row in column 0; oa / ob = 1 when that first row is odd.
template <bool LDL>
__device__ __forceinline__ void gemm_tile(double (&acc)[4][2][2], Pipe &pp, const double *A, const double *B, bool same, int64_t ld,
                                          int K0, int K1, int oa, int ob, const double *dv, uint32_t fmask)
{
    const int n = (K1 - K0 + KC - 1) / KC;
    const int segA = 32 + oa, segB = 32 + ob;
    for (int i = 0; i < NSTAGE - 1; ++i) {
        if (i < n) issue_chunk<LDL>(pp, i, A, B, same, ld, K0 + i * KC, min(KC, K1 - K0 - i * KC), segA, segB, dv);
        cp_async_commit();
    }
    for (int c = 0; c < n; ++c) {
        cp_async_wait<NSTAGE - 2>();        // this thread's part of chunk c has landed
        __syncthreads();                    // ... everyone's has, and everyone is done with chunk c - 1
        const int cn = c + NSTAGE - 1;
        if (cn < n) issue_chunk<LDL>(pp, cn % NSTAGE, A, B, same, ld, K0 + cn * KC, min(KC, K1 - K0 - cn * KC), segA, segB, dv);
        cp_async_commit();
        const int stage = c % NSTAGE;
        const int nk4 = (min(KC, K1 - K0 - c * KC) + 3) & ~3;
        const double *sa = pp.buf + stage * STAGE_DOUBLES, *sb = same ? sa : sa + KC * XS;
        mma_chunk<LDL>(acc, sa + oa, sb + ob, nk4, pp.dsm + stage * KC, fmask);
    }
    cp_async_wait<0>();
    __syncthreads();                        // the stage buffers are free (epilogues reuse them)
}

// Visit the 16 accumulator entries of this thread with their (row, col) inside the 64 x 64 tile.
template <typename F>
__device__ __forceinline__ void acc_foreach(double (&acc)[4][2][2], F f)
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

__device__ __forceinline__ void acc_zero(double (&acc)[4][2][2])
{
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

// out = X inv(L11)' for the panel solve: inv(L11) is lower triangular, so output columns [16 wn, 16 wn + 16) only need
// K < 16 wn + 16. The two warps that share a scheduler (w and w + 4) take column blocks wn and 3 - wn, so every tensor
// pipe gets the same 16 + 64 or 32 + 48 columns of K; f(row, col, value) visits the 16 results of the thread.
template <typename F>
__device__ __forceinline__ void trsm_product(double (&acc)[4][2][2], const double *Xs, const double *Ys, int nk4, int nrow, int ncol, F f)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = wm ? 3 - (warp & 3) : (warp & 3), g = lane >> 2, l3 = lane & 3;
    const double *xa = Xs + l3 * XS + wm * 32 + g;
    const double *yb = Ys + l3 * XS + wn * 16 + g;
    const int kend = min(nk4, wn * 16 + 16);
    if (wm * 32 < nrow && wn * 16 < ncol) {
#pragma unroll 1
        for (int kk = 0; kk < kend; kk += 4) {
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
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) f(wm * 32 + mi * 8 + g, wn * 16 + ni * 8 + l3 * 2 + e, acc[mi][ni][e]);
}

// ------------------------------------------------------------------ tasks of the factorization
__global__ void __launch_bounds__(256)
k_build_a2l(int64_t n, int64_t nnz, const int32_t *__restrict__ colptr, const int32_t *__restrict__ rowval,
            const int32_t *__restrict__ iperm, const int32_t *__restrict__ col2sn, const int32_t *__restrict__ sn_ptr,
            const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ row_idx, const int64_t *__restrict__ lp,
            int64_t *__restrict__ a2l, int *__restrict__ bad)
{
    // destination of every input nonzero in the panel storage (the host version is ls_build_a2l, ls_symbolic.cpp)
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int64_t lo = 0, hi = n;                 // column of position p: last j with colptr[j] <= p
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (colptr[mid] <= p) lo = mid; else hi = mid;
    }
    const int32_t a = iperm[rowval[p]], b = iperm[lo];
    const int32_t r = max(a, b), c = min(a, b);
    const int32_t s = col2sn[c];
    const int32_t c0 = sn_ptr[s], c1 = sn_ptr[s + 1];
    const int64_t k = c1 - c0, r0 = row_ptr[s], nr = row_ptr[s + 1] - r0;
    int64_t tt;
    if (r < c1) {
        tt = r - c0;
    } else {
        int64_t x = 0, y = nr;              // lower bound of r in the front's row list
        while (x < y) {
            const int64_t mid = (x + y) >> 1;
            if (row_idx[r0 + mid] < r) x = mid + 1; else y = mid;
        }
        if (x >= nr || row_idx[r0 + x] != r) { *bad = 1; a2l[p] = 0; return; }
        tt = k + x;
    }
    a2l[p] = lp[s] + (int64_t)(c - c0) * ((k + nr + 1) & ~(int64_t)1) + tt;
}

__global__ void __launch_bounds__(256)
k_scatter_a(int64_t nnz, const int64_t *__restrict__ a2l, const double *__restrict__ Ax, double *__restrict__ L)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) L[a2l[p]] = Ax[p];
}

// Extend-add: task = (parent s, parent-front columns [q0, q1)). The CTA walks the children of s in a fixed order and adds
// the part of each child's update matrix that lands in its column range, so every destination entry is owned by exactly
// one CTA -> deterministic sums.
__device__ void task_extend_add(const FactorParams &p, const FrontInfo &f, int q0, int q1, int off)
{
    const int k = f.k, r = f.r, ld = front_ld(f.k, f.r);
    double *P = p.L + f.lp;
    double *Us = p.U + f.up;
    // per task: the number of children that reach this column range, then (child position, b0, b1) for each of them
    // (a front of a K2 system can have thousands of tiny leaf children, almost all of them outside any one range)
    const int32_t *ranges = p.sched + off;
    const int nrel = ranges[0];
    // Each warp owns a slice of the task's parent columns and walks ALL children for it, in order: destinations of
    // different warps are disjoint, so no CTA barrier separates the children and the summation order per entry is still
    // the child order.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cw = (q1 - q0 + 7) >> 3;
    const int lo = q0 + warp * cw, hi = min(q1, lo + cw);
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
            double *dst = (tb < k) ? (P + (int64_t)tb * ld) : (Us + (int64_t)(tb - k) * r - k);
            // rows of one child column land on distinct parent rows: four independent read-modify-writes in flight
            // per lane (the plain loop is a chain of dependent global round trips)
            int a = b + lane;
            for (; a + 224 < rc; a += 256) {            // eight in flight: tall children (border rows of block-angular fronts)
                int rr[8];
                double sv[8], dv[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) { rr[t] = rel[a + 32 * t]; sv[t] = src[a + 32 * t]; }
#pragma unroll
                for (int t = 0; t < 8; ++t) dv[t] = dst[rr[t]];
#pragma unroll
                for (int t = 0; t < 8; ++t) dst[rr[t]] = dv[t] + sv[t];
            }
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

// Small leaf front (no children, k <= SL_K, k + r <= SL_N): one warp, lane i owns row i of the panel in registers;
// pivots and multipliers travel by shuffles; the update matrix is written directly (a leaf receives nothing, so
// U = -L21 D L21'). K2 systems have ~10^5 of these (one per primal variable).
template <bool LDL>
__device__ void leaf_factor(const FactorParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const int lane = threadIdx.x & 31;
    double *P = p.L + f.lp;
    double *U = p.U + f.up;
    double pr[SL_K], dd[SL_K];
#pragma unroll
    for (int c = 0; c < SL_K; ++c) pr[c] = (lane < N && c < k && lane >= c) ? P[(int64_t)c * ld + lane] : 0.0;
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
        if (c < k && lane < N && lane >= c) P[(int64_t)c * ld + lane] = pr[c];
    // update matrix of the leaf: U(i, c) = -sum_j L(i, j) d_j L(c, j), k <= c <= i < N
    double ldv[SL_K];
#pragma unroll
    for (int j = 0; j < SL_K; ++j) ldv[j] = (j < k) ? pr[j] * dd[j] : 0.0;
    for (int c = k; c < N; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < SL_K; ++j) {
            const double lc = __shfl_sync(0xffffffffu, pr[j], c);
            acc = fma(ldv[j], lc, acc);
        }
        if (lane >= c && lane < N) U[(int64_t)(c - k) * r + (lane - k)] = -acc;
    }
    if (lane == 0) {
        if (nbad) atomicMax(&p.info[0], 1);
        if (ntiny) atomicAdd(&p.info[2], ntiny);
        if (LDL && nneg) atomicAdd(&p.info[1], nneg);
    }
    __syncwarp();
    if (lane == 0 && f.parent >= 0) {                        // a leaf is complete as soon as its warp is done
        __threadfence();
        atomicAdd(&p.prog[f.parent], 1);
    }
}

// Diagonal block at column jb: S = P[jb.., jb..] - sum over the finished columns [K0, jb) of the super-panel (left-looking,
// accumulated in registers), then the 64 x 64 factorization + inverse of diag_block.cuh; both are written out.
// Cholesky: a non-positive pivot sets info[0] and is replaced by 1 so the run stays finite (the host then retries with
// more regularization like src/linear_solver.jl:6-17). LDL^T: unit-lower L11 with D on the diagonal; |pivot| < piv_tol is
// replaced by +-piv_tol; the pivots also go to Dg for the scaled updates.
template <bool LDL>
__device__ void task_diag(const FactorParams &p, Pipe &pp, const FrontInfo &f, int jb, int K0, double *smem)
{
    const int k = f.k, ld = front_ld(f.k, f.r);
    const int nb = min(NB, k - jb);
    double *P = p.L + f.lp;
    double *Pd = P + (int64_t)jb * ld + jb;
    double *Dv = p.Dinv + (f.dinv + (jb >> 6)) * (int64_t)(NB * XS);
    bool pre = false;
    if (jb > K0) {
        double acc[4][2][2];
        acc_zero(acc);
        gemm_tile<LDL>(acc, pp, P + jb, P + jb, true, ld, K0, jb, 0, 0, p.Dg + f.c0, frag_mask(nb, nb, true));
        double *S = smem;
        acc_foreach(acc, [&](int rr, int cc, double val) {
            S[cc * mipm_diag::DLD + rr] = (rr < nb && cc < nb && rr >= cc) ? Pd[(int64_t)cc * ld + rr] - val : ((rr == cc) ? 1.0 : 0.0);
        });
        pre = true;
    }
    mipm_diag::diag_block<LDL>(Pd, ld, nb, Dv, XS, p.piv_tol, p.info, smem, pre);
    if (LDL) {
        __syncthreads();
        if (threadIdx.x < nb) p.Dg[f.c0 + jb + threadIdx.x] = Pd[(int64_t)threadIdx.x * ld + threadIdx.x];
    }
}

// 64 rows of the panel below the diagonal block at jb: X = P[rows, jb..] - sum over [K0, jb), then L = X inv(L11)' as a
// second tile product against the explicitly inverted diagonal block (LDL^T: divided by the pivots).
template <bool LDL>
__device__ void task_panel(const FactorParams &p, Pipe &pp, const FrontInfo &f, int jb, int K0, int row0, double *smem)
{
    const int k = f.k, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const int nb = min(NB, k - jb);
    const int nrow = min(TILE, N - row0);
    double *P = p.L + f.lp;
    double *R = P + (int64_t)jb * ld + row0;
    const double *Dv = p.Dinv + (f.dinv + (jb >> 6)) * (int64_t)(NB * XS);
    const uint32_t fm = frag_mask(nrow, nb, false);
    double acc[4][2][2];
    acc_zero(acc);
    if (jb > K0) gemm_tile<LDL>(acc, pp, P + (row0 & ~1), P + jb, false, ld, K0, jb, row0 & 1, 0, p.Dg + f.c0, fm);
    double *Xs = smem, *Ys = smem + TILE * XS;
    uint64_t *ebar = pp.bar;
    if (threadIdx.x == 0) {             // inverse block -> Ys: stored with the padded stride XS, so ONE TMA bulk copy (34 KB)
        mbar_expect_tx(ebar, NB * XS * 8);
        bulk_g2s(Ys, Dv, NB * XS * 8, ebar);
    }
    acc_foreach(acc, [&](int rr, int cc, double val) {
        Xs[cc * XS + rr] = (rr < nrow && cc < nb) ? R[(int64_t)cc * ld + rr] - val : 0.0;
    });
    mbar_wait(ebar, pp.ephase);
    pp.ephase ^= 1u;
    __syncthreads();
    acc_zero(acc);
    const double *dv = p.Dg + f.c0 + jb;
    trsm_product(acc, Xs, Ys, (nb + 3) & ~3, nrow, nb, [&](int rr, int cc, double val) {
        if (rr < nrow && cc < nb) R[(int64_t)cc * ld + rr] = LDL ? val / dv[cc] : val;
    });
}

// Trailing update on the FP64 tensor pipe: C(tile) -= L[rows, K0:K1] D L[cols, K0:K1]', the whole super-panel accumulated
// in registers, C read and written once. C is a tile of the remaining panel columns or of the update matrix.
template <bool LDL>
__device__ void task_trail(const FactorParams &p, Pipe &pp, const FrontInfo &f, int K0, int K1, int row0, int col0)
{
    const int k = f.k, r = f.r, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const int nrow = min(TILE, ((row0 < k) ? k : N) - row0), ncol = min(TILE, ((col0 < k) ? k : N) - col0);
    const bool diag = (row0 == col0);
    double *P = p.L + f.lp;
    const uint32_t fm = frag_mask(nrow, ncol, diag);
    double acc[4][2][2];
    acc_zero(acc);
    gemm_tile<LDL>(acc, pp, P + (row0 & ~1), P + (col0 & ~1), diag, ld, K0, K1, row0 & 1, col0 & 1, p.Dg + f.c0, fm);
    double *C;
    int64_t ldc;
    if (col0 < k) { C = P + (int64_t)col0 * ld + row0; ldc = ld; }
    else { C = p.U + f.up + (int64_t)(col0 - k) * r + (row0 - k); ldc = r; }
    // read all 16 destination entries first, then subtract and store: as read-modify-writes in sequence they are a
    // chain of dependent global round trips (the compiler must assume the stores alias the next load)
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
                const bool on = rr < nrow && cc < ncol && (!diag || rr >= cc);
                cv[mi][ni][e] = on ? C[(int64_t)cc * ldc + rr] : 0.0;
            }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int rr = wm * 32 + mi * 8 + g, cc = wn * 16 + ni * 8 + l3 * 2 + e;
                const bool on = rr < nrow && cc < ncol && (!diag || rr >= cc);
                if (on) C[(int64_t)cc * ldc + rr] = cv[mi][ni][e] - acc[mi][ni][e];
            }
}

template <bool LDL>
__global__ void __launch_bounds__(256, FACTOR_OCC) k_factor_tasks(FactorParams p)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ int s_ticket;
    __shared__ unsigned long long s_busy[8];      // thread 0: ns per task class, [5] = dependency wait
    Pipe pp;
    pp.buf = smem;
    pp.dsm = smem + SMEM_DOUBLES;
    pp.bar = reinterpret_cast<uint64_t *>(pp.dsm + NSTAGE * KC);
    pp.ephase = 0;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) s_busy[i] = 0;
        mbar_init(pp.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        s_ticket = p.task_begin + atomicAdd(p.ticket, 1);
    }
    __syncthreads();
    unsigned long long t_first = 0, t_last = 0;
    if (tid == 0) t_first = globaltimer_ns();
    for (;;) {
        const int t = s_ticket;
        if (t >= p.task_end) break;
        const Task tk = p.tasks[t];
        int next = 0;
        unsigned long long t0 = 0, t1 = 0;
        FrontInfo f;
        const int ttype = tk.type & 0xff;
        if (ttype != T_LEAF) {
            f = p.fi[tk.front];
            if (tid == 0) {
                t0 = globaltimer_ns();
                while (ld_acquire(p.prog + tk.front) < tk.need) __nanosleep(40);
                if (tk.need_b > 0) { while (ld_acquire(p.prog_b + tk.front) < tk.need_b) __nanosleep(40); }
                asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies below read what other CTAs wrote
                t1 = globaltimer_ns();
                next = p.task_begin + atomicAdd(p.ticket, 1);      // in flight while the task runs
            }
        } else if (tid == 0) {
            t0 = t1 = globaltimer_ns();
            next = p.task_begin + atomicAdd(p.ticket, 1);
        }
        __syncthreads();
        switch (ttype) {
        case T_LEAF: {
            const int li = tk.a + (tid >> 5);
            if ((tid >> 5) < tk.b) leaf_factor<LDL>(p, p.sched[li]);
            break;
        }
        case T_EA: task_extend_add(p, f, tk.a, tk.b, tk.c); break;
        case T_DIAG: task_diag<LDL>(p, pp, f, tk.a, tk.b, smem); break;
        case T_PANEL: task_panel<LDL>(p, pp, f, tk.a, tk.b, tk.c, smem); break;
        default: task_trail<LDL>(p, pp, f, tk.a, tk.b, tk.c, tk.d); break;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes of this task before later bulk copies
        __syncthreads();
        if (tid == 0) {
            if (ttype != T_LEAF) {
                __threadfence();
                atomicAdd(((tk.type & T_DEFERRED) ? p.prog_b : p.prog) + tk.front, 1);
                const int old = atomicAdd(p.done + tk.front, 1);
                if (old + 1 == f.total) {
                    if (p.front_ns) p.front_ns[tk.front] = globaltimer_ns();
                    if (f.parent >= 0) {
                        __threadfence();
                        atomicAdd(p.prog + f.parent, 1);
                    }
                }
            }
            t_last = globaltimer_ns();
            s_busy[5] += t1 - t0;
            s_busy[ttype] += t_last - t1;
            if (p.trace) {
                unsigned long long *tr = p.trace + 4 * (size_t)t;
                tr[0] = t0; tr[1] = t1; tr[2] = t_last; tr[3] = (unsigned long long)read_smid();
            }
            s_ticket = next;
        }
        __syncthreads();
    }
    if (tid == 0 && p.prof) {
        unsigned long long *o = p.prof + 8 * (size_t)blockIdx.x;
        for (int c = 0; c < 6; ++c) o[c] = s_busy[c];
        o[6] = t_first;
        o[7] = t_last ? t_last : t_first;
    }
}

// Micro-benchmark of the same pipelined tile code: C (n x n, lower tiles) -= X X' with K = kdim (X: n x kdim, ld even).
__global__ void __launch_bounds__(256, FACTOR_OCC)
k_bench_syrk(int n, int kdim, double *C, int64_t ldc, const double *X, int64_t ldx)
{
    extern __shared__ __align__(128) double smem[];
    Pipe pp;
    pp.buf = smem;
    pp.dsm = smem + SMEM_DOUBLES;
    pp.bar = reinterpret_cast<uint64_t *>(pp.dsm + NSTAGE * KC);
    pp.ephase = 0;
    if (threadIdx.x == 0) {
        mbar_init(pp.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    const int nt = (n + TILE - 1) / TILE, ntiles = nt * (nt + 1) / 2;
    for (int lt = blockIdx.x; lt < ntiles; lt += gridDim.x) {
        int tr = (int)((sqrt(8.0 * (double)lt + 1.0) - 1.0) * 0.5);
        while (tr * (tr + 1) / 2 > lt) --tr;
        while ((tr + 1) * (tr + 2) / 2 <= lt) ++tr;
        const int tc = lt - tr * (tr + 1) / 2;
        const int row0 = tr * TILE, col0 = tc * TILE;
        const int nrow = min(TILE, n - row0), ncol = min(TILE, n - col0);
        const bool diag = (tr == tc);
        const uint32_t fm = frag_mask(nrow, ncol, diag);
        double acc[4][2][2];
        acc_zero(acc);
        gemm_tile<false>(acc, pp, X + row0, X + col0, diag, ldx, 0, kdim, 0, 0, nullptr, fm);
        acc_foreach(acc, [&](int rr, int cc, double val) {
            if (rr < nrow && cc < ncol && (row0 + rr) >= (col0 + cc)) C[(int64_t)(col0 + cc) * ldc + row0 + rr] -= val;
        });
    }
}

inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)std::max<int64_t>(1, (n + per_block - 1) / per_block); }

}  // namespace

int ls_device_setup(Handle *h)
{
    const LsSymbolic &S = h->sym;
    const int ns = S.ns;
    const bool tlog = std::getenv("MIPM_ANALYZE_LOG") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto tlog_stage = [&](const char *name) {
        if (!tlog) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "device setup: %s %.3f s\n", name, std::chrono::duration<double>(now - t_prev).count());
        t_prev = now;
    };
    MIPM_CUDA(h, cudaSetDevice(h->device));
    // a re-analysis must not release buffers the side stream is still zero-filling
    if (h->side) MIPM_CUDA(h, cudaStreamSynchronize(h->side));
    // ---- fronts
    std::vector<int64_t> dinv_off((size_t)ns + 1, 0);
    std::vector<FrontInfo> finfo((size_t)std::max(ns, 1));
    std::vector<char> small((size_t)std::max(ns, 1), 0);
    const bool use_leaf = std::getenv("MIPM_NO_LEAF") == nullptr;
    for (int s = 0; s < ns; ++s) {
        int64_t k = S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s];
        int64_t r = S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s];
        // (the border root of a distributed factorization is never a small leaf: its factorization is a stage of its own)
        small[(size_t)s] = use_leaf && s != S.root_sn && S.child_ptr[(size_t)s + 1] == S.child_ptr[(size_t)s] && k <= SL_K && k + r <= SL_N;
        dinv_off[(size_t)s + 1] = dinv_off[(size_t)s] + (small[(size_t)s] ? 0 : (k + NB - 1) / NB);
        FrontInfo &f = finfo[(size_t)s];
        f.k = (int32_t)k; f.r = (int32_t)r; f.c0 = S.sn_ptr[(size_t)s];
        f.nchild = (int32_t)(S.child_ptr[(size_t)s + 1] - S.child_ptr[(size_t)s]);
        f.lp = S.lp[(size_t)s]; f.up = S.up[(size_t)s]; f.dinv = dinv_off[(size_t)s];
        f.rowp = S.row_ptr[(size_t)s]; f.childp = S.child_ptr[(size_t)s];
        f.parent = S.sn_parent[(size_t)s]; f.total = 0;
    }
    DeviceInfo prop;
    if (device_info(h->device, prop) != MIPM_OK) return fail(h, MIPM_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (!prop.cooperative) return fail(h, MIPM_ERR_CUDA, "device does not support cooperative launch");
    const int grid_estimate = prop.sm_count * FACTOR_OCC;
    // super-panel width: columns whose updates are accumulated in registers before the trailing matrix is touched.
    // Measured on C1, C2, its mesh variant and C4 (profiles/r02_notes.md): 64 (one block step per panel) is as fast or
    // faster than 128 / 256 everywhere -- the operand re-reads of a right-looking update hit L2, while a wider panel puts
    // its left-looking K loops on the dependency chain of the front. Wider panels stay available for sweeps.
    // Exception: a level with at least as many wide fronts as the grid has CTAs (a stacked batch of dense blocks) is
    // throughput bound, not chain bound -- there the wide panel's saving in C traffic wins (C5: 0.278 -> 0.249 s).
    int SP_narrow = 64, SP_wide = 256;
    int ea_task_factor = 4, stagger = 1;
    int la_min_k = 512;          // fronts at least this wide factor with look-ahead (second progress counter)
    if (const char *e = std::getenv("MIPM_LOOKAHEAD_MIN")) la_min_k = std::max(2 * NB, atoi(e));
    const bool fuse_diag = std::getenv("MIPM_NO_FUSE_DIAG") == nullptr;
    if (const char *e = std::getenv("MIPM_EA_FACTOR")) ea_task_factor = std::max(1, atoi(e));
    if (const char *e = std::getenv("MIPM_NO_STAGGER")) stagger = (atoi(e) == 0);
    if (const char *e = std::getenv("MIPM_SUPER_PANEL")) SP_narrow = SP_wide = std::max(NB, (atoi(e) / NB) * NB);
    if (const char *e = std::getenv("MIPM_SUPER_PANEL_WIDE")) SP_wide = std::max(NB, (atoi(e) / NB) * NB);
    // ---- task list: level by level (children before parents); inside a level breadth-first over the fronts' task
    // groups, so that consecutive tickets rarely wait on each other. `sched` keeps the extend-add child ranges, the
    // small-leaf list and the per-level front lists of the solves.
    std::vector<int32_t> sched;
    std::vector<int64_t> lvl;
    std::vector<Task> tasks;
    int64_t leaf_off = 0;
    int n_leaf = 0;
    int root_task_begin = -1;
    auto mk = [](int type, int front, int a, int b, int c, int d) {
        Task t;
        t.type = type; t.front = front; t.a = a; t.b = b; t.c = c; t.d = d; t.need = 0; t.need_b = 0;
        return t;
    };
    for (int l = 0; l < S.n_levels; ++l) {
        const int64_t f0 = S.level_ptr[(size_t)l], f1 = S.level_ptr[(size_t)l + 1];
        if (l == 0) {            // small leaf fronts: their own list, one warp each, eight per task
            leaf_off = (int64_t)sched.size();
            for (int64_t t = f0; t < f1; ++t) {
                int s = S.level_sn[(size_t)t];
                if (small[(size_t)s]) { sched.push_back(s); n_leaf++; }
            }
            for (int i = 0; i < n_leaf; i += 8) tasks.push_back(mk(T_LEAF, -1, (int)leaf_off + i, std::min(8, n_leaf - i), 0, 0));
        }
        lvl.push_back((int64_t)sched.size());
        int64_t n_reg = 0;
        for (int64_t t = f0; t < f1; ++t) {
            int s = S.level_sn[(size_t)t];
            if (small[(size_t)s]) continue;
            sched.push_back(s);
            n_reg++;
        }
        lvl.push_back(n_reg);
        int64_t n_wide = 0;
        for (int64_t t = f0; t < f1; ++t) n_wide += finfo[(size_t)S.level_sn[(size_t)t]].k > 2 * NB;
        const int SP = (n_wide >= grid_estimate) ? SP_wide : SP_narrow;
        // columns per extend-add task: EA_COLS, narrower near the root where a level has fewer tasks than CTAs (an
        // extend-add task is a chain of dependent global round trips, so the level costs one task's latency)
        int ea_cols = EA_COLS;
        for (;;) {
            int64_t nt = 0;
            for (int64_t t = f0; t < f1; ++t) {
                const FrontInfo &f = finfo[(size_t)S.level_sn[(size_t)t]];
                if (f.nchild > 0) nt += (f.k + f.r + ea_cols / 2 - 1) / (ea_cols / 2);
            }
            if (ea_cols <= 4 || nt > ea_task_factor * grid_estimate) break;
            ea_cols /= 2;
        }
        std::vector<std::vector<std::vector<Task>>> groups;       // [front in level][group][task]
        size_t max_groups = 0;
        for (int64_t t = f0; t < f1; ++t) {
            const int s = S.level_sn[(size_t)t];
            if (small[(size_t)s]) continue;
            FrontInfo &f = finfo[(size_t)s];
            const int k = f.k, N = f.k + f.r;
            std::vector<std::vector<Task>> G;
            // extend-add: per-child column ranges first, then the task records that point at them
            if (f.nchild > 0) {
                std::vector<Task> ea;
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
                    ea.push_back(mk(T_EA, s, q0, q1, (int32_t)off_r, 0));
                }
                if (!ea.empty()) G.push_back(std::move(ea));
            }
            const size_t n_ea_groups = G.size();
            auto trailing_starts = [&](int J1) {
                std::vector<int> st2;
                for (int x = J1; x < k; x += TILE) st2.push_back(x);
                for (int x = k; x < N; x += TILE) st2.push_back(x);
                return st2;
            };
            const bool lookahead = k >= la_min_k && (k + SP - 1) / SP >= 2;
            if (!lookahead) {
                // The first diagonal block of panel p + 1 applies panel p's update to its own tile itself (left-looking over
                // [J0, J1): task_diag with K0 = J0) instead of waiting for a trailing-update task to do it: it is emitted at
                // the head of panel p's trailing group, so it starts as soon as the panel tiles of p are done and runs
                // beside the other trailing tiles (one hand-off and one tile round trip less per 64 columns of the chain).
                bool diag_emitted = false;
                for (int J0 = 0; J0 < k;) {
                    int J1 = (k - J0 <= SP + NB) ? k : J0 + SP;
                    for (int jb = J0; jb < J1; jb += NB) {
                        if (!(jb == J0 && diag_emitted)) G.push_back({mk(T_DIAG, s, jb, J0, 0, 0)});
                        const int j1 = jb + std::min(NB, k - jb);
                        std::vector<Task> pt;
                        for (int row0 = j1; row0 < N; row0 += TILE) pt.push_back(mk(T_PANEL, s, jb, J0, row0, 0));
                        if (!pt.empty()) G.push_back(std::move(pt));
                    }
                    // trailing tiles: tile starts over the remaining panel columns [J1, k) and over the rows below [k, N)
                    std::vector<int> starts = trailing_starts(J1);
                    if ((int64_t)starts.size() * ((int64_t)starts.size() + 1) / 2 > (1 << 26)) return fail(h, MIPM_ERR_ARG, "front too large for the tile schedule");
                    std::vector<Task> tt;
                    const bool fuse = fuse_diag && J1 < k;
                    if (fuse) tt.push_back(mk(T_DIAG, s, J1, J0, 0, 0));
                    for (size_t tc = 0; tc < starts.size(); ++tc)
                        for (size_t tr = tc; tr < starts.size(); ++tr) {
                            if (fuse && tc == 0 && tr == 0) continue;          // the tile of the next diagonal block
                            tt.push_back(mk(T_TRAIL, s, J0, J1, starts[tr], starts[tc]));
                        }
                    diag_emitted = fuse;
                    if (!tt.empty()) G.push_back(std::move(tt));
                    J0 = J1;
                }
                int done = f.nchild;
                for (auto &g : G) {
                    for (auto &t2 : g) t2.need = done;
                    done += (int)g.size();
                }
                f.total = done - f.nchild;
            } else {
                // Look-ahead for wide fronts: the trailing tiles of panel p are split into the block columns of panel p + 1
                // (LA: on the chain counter) and the rest (REST: on the second counter). Panel p + 1 only waits for LA(p),
                // so its diagonal blocks and panel solves run while REST(p) is still being applied; tiles that update the
                // same C tile stay ordered through the second counter. List order per panel p:
                //   DIAG(p, first step) | REST(p-1) | remaining steps of panel p | LA(p)          (... | REST(last))
                std::vector<std::array<int, 2>> panels;
                for (int J0 = 0; J0 < k;) {
                    int J1 = (k - J0 <= SP + NB) ? k : J0 + SP;
                    panels.push_back({J0, J1});
                    J0 = J1;
                }
                int cntA = f.nchild, cntB = 0;           // completions so far on the two counters
                int needB_chain = 0;                     // REST completed through panel p-2 (what panel p's chain needs)
                int restB_prev = 0;                      // REST completed through panel p-1
                for (auto &g : G) { for (auto &t2 : g) t2.need = cntA; cntA += (int)g.size(); }      // extend-add
                std::vector<Task> rest_prev;             // REST(p-1), emitted after the first DIAG of panel p
                bool diag_emitted = false;               // the first DIAG of this panel already sits at the head of LA(p-1)
                for (size_t pi = 0; pi < panels.size(); ++pi) {
                    const int J0 = panels[pi][0], J1 = panels[pi][1];
                    bool first = true;
                    for (int jb = J0; jb < J1; jb += NB) {
                        if (!(jb == J0 && diag_emitted)) {
                            Task dg = mk(T_DIAG, s, jb, J0, 0, 0);
                            dg.need = cntA; dg.need_b = needB_chain;
                            G.push_back({dg});
                            cntA += 1;
                        }
                        if (first && !rest_prev.empty()) {
                            G.push_back(std::move(rest_prev));
                            rest_prev.clear();
                        }
                        first = false;
                        const int j1 = jb + std::min(NB, k - jb);
                        std::vector<Task> pt;
                        for (int row0 = j1; row0 < N; row0 += TILE) {
                            Task q = mk(T_PANEL, s, jb, J0, row0, 0);
                            q.need = cntA; q.need_b = needB_chain;
                            pt.push_back(q);
                        }
                        if (!pt.empty()) { cntA += (int)pt.size(); G.push_back(std::move(pt)); }
                    }
                    const int chain_done = cntA;         // everything of panel p
                    std::vector<int> starts = trailing_starts(J1);
                    if ((int64_t)starts.size() * ((int64_t)starts.size() + 1) / 2 > (1 << 26)) return fail(h, MIPM_ERR_ARG, "front too large for the tile schedule");
                    const int la_end = (pi + 1 < panels.size()) ? panels[pi + 1][1] : J1;     // columns [J1, la_end) = next panel
                    std::vector<Task> la, rest;
                    // fused diagonal update (see the branch above): the next panel's first diagonal block takes the place of
                    // the look-ahead tile (J1, J1) with the same two dependency counts
                    const bool fuse = fuse_diag && J1 < k;
                    if (fuse) {
                        Task dg = mk(T_DIAG, s, J1, J0, 0, 0);
                        dg.need = chain_done; dg.need_b = restB_prev;
                        la.push_back(dg);
                    }
                    diag_emitted = fuse;
                    for (size_t tc = 0; tc < starts.size(); ++tc)
                        for (size_t tr = tc; tr < starts.size(); ++tr) {
                            if (fuse && tc == 0 && tr == 0) continue;
                            Task q = mk(T_TRAIL, s, J0, J1, starts[tr], starts[tc]);
                            q.need = chain_done;
                            q.need_b = restB_prev;       // the same C tile was last touched by REST(p-1)
                            if (starts[tc] < la_end && starts[tc] < k) la.push_back(q);
                            else { q.type |= T_DEFERRED; rest.push_back(q); }
                        }
                    needB_chain = restB_prev;            // panel p+1 needs REST through p-1 ... (see below)
                    if (!la.empty()) { cntA += (int)la.size(); G.push_back(std::move(la)); }
                    cntB += (int)rest.size();
                    // panel p + 1 reads columns that REST(p - 1) updated, not REST(p): its need_b is the count BEFORE this panel's REST
                    restB_prev = cntB;
                    rest_prev = std::move(rest);
                }
                if (!rest_prev.empty()) G.push_back(std::move(rest_prev));
                f.total = (cntA - f.nchild) + cntB;
            }
            if (s == S.root_sn) {
                // staged (distributed) factorization: the root's extend-add belongs to stage 0, the rest to stage 1;
                // the root is alone on the last level, so its groups are appended in order below
                root_task_begin = (int)tasks.size();
                for (size_t g = 0; g < n_ea_groups; ++g) root_task_begin += (int)G[g].size();
            }
            max_groups = std::max(max_groups, G.size());
            groups.push_back(std::move(G));
        }
        // Emission order inside the level: front i starts `i mod max_groups` slots late, so that any stretch of the list
        // mixes latency-bound tasks (EA, DIAG) of some fronts with tensor-pipe tasks (PANEL, TRAIL) of others instead of
        // running ~1000 diagonal blocks at once, three to an SM. Within a front the group order is kept.
        const size_t n_slots = max_groups ? 2 * max_groups : 0;
        for (size_t slot = 0; slot < n_slots; ++slot)
            for (size_t i = 0; i < groups.size(); ++i) {
                const size_t phi = stagger ? (i % max_groups) : 0;
                if (slot < phi) continue;
                const size_t g = slot - phi;
                if (g < groups[i].size()) tasks.insert(tasks.end(), groups[i][g].begin(), groups[i][g].end());
            }
        if (tasks.size() > (size_t)(1 << 30) || sched.size() > (size_t)INT32_MAX - 16) return fail(h, MIPM_ERR_ARG, "schedule too large");
    }
    tlog_stage("task list");
    h->leaf_off = leaf_off;
    h->n_leaf = n_leaf;
    h->root_task_begin = root_task_begin;
    h->n_tasks = (int)tasks.size();
    h->n_launch_factor = 2;   // scatter + task kernel (plus memsets)
    // ---- grid sizes
    MIPM_CUDA(h, cudaFuncSetAttribute(k_factor_tasks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MIPM_CUDA(h, cudaFuncSetAttribute(k_factor_tasks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MIPM_CUDA(h, cudaFuncSetAttribute(k_bench_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int occ_f = 0;
    if (S.kind == MIPM_LDL) MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_factor_tasks<true>, 256, SMEM_BYTES));
    else MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_factor_tasks<false>, 256, SMEM_BYTES));
    if (occ_f < 1) return fail(h, MIPM_ERR_CUDA, "factorization kernel does not fit on an SM");
    h->grid_factor = prop.sm_count * occ_f;
    if (h->grid_limit > 0) h->grid_factor = std::min(h->grid_factor, h->grid_limit);   // small systems side by side (mipm_set_grid_limit)
    h->grid_factor = std::max(1, std::min(h->grid_factor, h->n_tasks));
    // ---- uploads and workspaces
    cudaStream_t st = h->stream;
    MIPM_CUDA(h, h->d_sched.upload(sched, st));
    MIPM_CUDA(h, h->d_finfo.alloc(finfo.size()));
    MIPM_CUDA(h, cudaMemcpyAsync(h->d_finfo.p, finfo.data(), finfo.size() * sizeof(FrontInfo), cudaMemcpyHostToDevice, st));
    MIPM_CUDA(h, h->d_tasks.alloc(std::max<size_t>(tasks.size(), 1)));
    if (!tasks.empty()) MIPM_CUDA(h, cudaMemcpyAsync(h->d_tasks.p, tasks.data(), tasks.size() * sizeof(Task), cudaMemcpyHostToDevice, st));
    MIPM_CUDA(h, h->d_prog.alloc((size_t)3 * std::max(ns, 1) + 4));
    MIPM_CUDA(h, h->d_prof.alloc((size_t)8 * (size_t)h->grid_factor));
    MIPM_CUDA(h, h->d_lvl.upload(lvl, st));
    MIPM_CUDA(h, h->d_row_idx.upload(S.row_idx, st));
    MIPM_CUDA(h, h->d_rel_idx.upload(S.rel_idx, st));
    MIPM_CUDA(h, h->d_perm.upload(S.perm, st));
    MIPM_CUDA(h, h->d_child_idx.upload(S.child_idx, st));
    DBuf<int32_t> t_colptr, t_rowval, t_iperm, t_col2sn, t_snptr;
    DBuf<int64_t> t_rowptr, t_lp;
    DBuf<int> t_bad;
    const bool device_a2l = S.a2l.empty() && S.nnz_a > 0;
    if (device_a2l) {
        // scatter map built on the device from the pattern (the host copy would be 8 bytes per nonzero to compute,
        // page in and upload)
        MIPM_CUDA(h, t_colptr.upload(S.in_colptr, st));
        MIPM_CUDA(h, t_rowval.upload(S.in_rowval, st));
        MIPM_CUDA(h, t_iperm.upload(S.iperm, st));
        MIPM_CUDA(h, t_col2sn.upload(S.col2sn, st));
        MIPM_CUDA(h, t_snptr.upload(S.sn_ptr, st));
        MIPM_CUDA(h, t_rowptr.upload(S.row_ptr, st));
        MIPM_CUDA(h, t_lp.upload(S.lp, st));
        MIPM_CUDA(h, t_bad.alloc(1));
        MIPM_CUDA(h, cudaMemsetAsync(t_bad.p, 0, sizeof(int), st));
        MIPM_CUDA(h, h->d_a2l.alloc((size_t)S.nnz_a));
        k_build_a2l<<<grid_for(S.nnz_a, 256), 256, 0, st>>>(S.n, S.nnz_a, t_colptr.p, t_rowval.p, t_iperm.p, t_col2sn.p, t_snptr.p,
                                                            t_rowptr.p, h->d_row_idx.p, t_lp.p, h->d_a2l.p, t_bad.p);
        MIPM_CHECK_LAUNCH(h);
    } else {
        MIPM_CUDA(h, h->d_a2l.upload(S.a2l, st));
    }
    h->d_full_ptr.release();        // refinement operator: built and uploaded on first use (ls_solve_impl)
    h->d_full_col.release();
    h->d_full_val.release();
    tlog_stage("uploads");
    // + 64 doubles: a tile's bulk copies may read past the end of the last panel (rows that are never used)
    MIPM_CUDA(h, h->d_L.alloc((size_t)std::max<int64_t>(S.nnz_l, 1) + 64));
    h->L_cur = h->d_L.p;
    h->l_prezeroed = false;
    h->d_L2.release();
    if (S.root_sn < 0 && S.nnz_l > 0 && (size_t)S.nnz_l * sizeof(double) <= ((size_t)16 << 30) && !std::getenv("MIPM_SINGLE_L")) {
        if (h->d_L2.alloc((size_t)S.nnz_l + 64) != cudaSuccess) { h->d_L2.release(); (void)cudaGetLastError(); }   // optional
    }
    MIPM_CUDA(h, h->d_U.alloc((size_t)std::max<int64_t>(S.update_doubles, 1)));
    MIPM_CUDA(h, h->d_Dinv.alloc((size_t)std::max<int64_t>(dinv_off[(size_t)ns], 1) * NB * XS));
    MIPM_CUDA(h, h->d_Dg.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_xp.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_uvec.alloc((size_t)std::max<int64_t>(S.row_ptr[(size_t)ns], 1)));
    MIPM_CUDA(h, h->d_b.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_r.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, cudaMemsetAsync(h->d_L.p + std::max<int64_t>(S.nnz_l, 1), 0, 64 * sizeof(double), st));
    if (h->d_L2.p) MIPM_CUDA(h, cudaMemsetAsync(h->d_L2.p + S.nnz_l, 0, 64 * sizeof(double), st));
    tlog_stage("workspace allocation");
    {
        int rc = ls_solve_setup(h, finfo.data(), small.data());
        if (rc != MIPM_OK) return rc;
    }
    int a2l_bad = 0;
    if (device_a2l) MIPM_CUDA(h, cudaMemcpyAsync(&a2l_bad, t_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    MIPM_CUDA(h, cudaStreamSynchronize(st));
    if (a2l_bad) return fail(h, MIPM_ERR_ARG, "input entry outside the symbolic structure (internal error)");
    tlog_stage("solve setup + sync");
    if (!h->side) {
        MIPM_CUDA(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        MIPM_CUDA(h, cudaEventCreateWithFlags(&h->ev_factor_done, cudaEventDisableTiming));
        MIPM_CUDA(h, cudaEventCreateWithFlags(&h->ev_u_zero, cudaEventDisableTiming));
    }
    h->u_prezeroed = false;
    h->factorized = false;
    return MIPM_OK;
}

// stage -1: everything; stage 0: assembly + all tasks below the root (border) front's own factorization, leaving the
// local Schur contribution in the root panel; stage 1: the root front's own factorization (after the all-reduce).
int ls_factorize_staged(Handle *h, const double *d_nzval, int stage)
{
    const LsSymbolic &S = h->sym;
    cudaStream_t st = h->stream;
    MIPM_CUDA(h, cudaSetDevice(h->device));
    if (stage >= 0 && h->root_task_begin < 0) return fail(h, MIPM_ERR_STATE, "staged factorization needs mipm_ls_analyze_border");
    const int t0 = (stage == 1) ? h->root_task_begin : 0;
    const int t1 = (stage == 0) ? h->root_task_begin : h->n_tasks;
    if (stage != 1) {
        if (h->d_L2.p) {
            // two factor buffers: this factorization writes the one the previous solves did not use; it was zero-filled
            // on the side stream while they ran
            h->L_cur = (h->L_cur == h->d_L.p) ? h->d_L2.p : h->d_L.p;
            if (h->l_prezeroed) { MIPM_CUDA(h, cudaStreamWaitEvent(st, h->ev_u_zero, 0)); }
            else { MIPM_CUDA(h, cudaMemsetAsync(h->L_cur, 0, (size_t)S.nnz_l * sizeof(double), st)); }
        } else {
            MIPM_CUDA(h, cudaMemsetAsync(h->d_L.p, 0, (size_t)std::max<int64_t>(S.nnz_l, 1) * sizeof(double), st));
        }
        // The update matrices are only live inside the factorization kernel, so their zero-fill for the NEXT
        // factorization runs on a side stream behind this one (it overlaps the latency-bound solves).
        if (h->u_prezeroed) {
            MIPM_CUDA(h, cudaStreamWaitEvent(st, h->ev_u_zero, 0));
        } else {
            MIPM_CUDA(h, cudaMemsetAsync(h->d_U.p, 0, (size_t)std::max<int64_t>(S.update_doubles, 1) * sizeof(double), st));
        }
        MIPM_CUDA(h, cudaMemsetAsync(h->d_info.p, 0, 4 * sizeof(int), st));
        MIPM_CUDA(h, cudaMemsetAsync(h->d_prog.p, 0, ((size_t)3 * std::max(S.ns, 1) + 4) * sizeof(int), st));
        if (S.nnz_a > 0) {
            k_scatter_a<<<grid_for(S.nnz_a, 256), 256, 0, st>>>(S.nnz_a, h->d_a2l.p, d_nzval, h->L_cur);
            MIPM_CHECK_LAUNCH(h);
        }
    } else {
        MIPM_CUDA(h, cudaMemsetAsync(h->d_prog.p + 3 * (size_t)std::max(S.ns, 1), 0, sizeof(int), st));      // the ticket counter only
    }
    if (t1 > t0) {
        FactorParams p;
        p.fi = (const FrontInfo *)h->d_finfo.p; p.child_idx = h->d_child_idx.p; p.rel_idx = h->d_rel_idx.p;
        p.sched = h->d_sched.p; p.tasks = (const Task *)h->d_tasks.p; p.task_begin = t0; p.task_end = t1;
        p.L = h->L_cur; p.U = h->d_U.p; p.Dinv = h->d_Dinv.p; p.Dg = h->d_Dg.p; p.info = h->d_info.p;
        p.prog = h->d_prog.p; p.prog_b = h->d_prog.p + std::max(S.ns, 1); p.done = h->d_prog.p + 2 * (size_t)std::max(S.ns, 1);
        p.ticket = h->d_prog.p + 3 * (size_t)std::max(S.ns, 1);
        p.prof = h->d_prof.p; p.front_ns = h->d_front_ns.p; p.trace = h->d_trace.p;
        p.piv_tol = h->ldl_definite ? 1e-300 : 1e-13;   // LDL^T: absolute floor on |pivot| (K2: the dual regularization is 1e-10)
        void *args[] = {&p};
        const void *fn = (S.kind == MIPM_LDL) ? (const void *)k_factor_tasks<true> : (const void *)k_factor_tasks<false>;
        // cooperative launch: guarantees that every CTA is resident, which the dependency spins rely on
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

}  // namespace mipm

extern "C" int mipm_ls_analyze_border(mipm_handle hh, int64_t n, const int32_t *colptr, const int32_t *rowval, int index_base,
                                      int kind, int64_t n_border)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    if (h && !h->host_only) use_handle(h);
    if (!h || n < 0 || !colptr || n_border < 1 || n_border > n || (kind != MIPM_CHOLESKY && kind != MIPM_LDL && kind != MIPM_LDL_DEFINITE))
        return fail(h, MIPM_ERR_ARG, "bad argument");
    LsOptions opt;
    opt.kind = (kind == MIPM_LDL_DEFINITE) ? MIPM_CHOLESKY : kind;
    opt.ordering = MIPM_ORDER_ND;
    opt.n_border = n_border;
    opt.host_a2l = h->host_only;
    if (const char *s = std::getenv("MIPM_ND_LEAF")) opt.nd_leaf = std::max(1, atoi(s));
    h->has_ls = false;
    h->factorized = false;
    std::string e = ls_analyze(n, colptr, rowval, index_base, opt, nullptr, h->sym);
    if (!e.empty()) return fail(h, MIPM_ERR_ARG, e);
    h->ldl_definite = (kind == MIPM_LDL_DEFINITE);
    if (h->ldl_definite) h->sym.kind = MIPM_LDL;
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

// One timed factorization; per-class CTA-busy time from the kernel's own %globaltimer accounting.
//   ms[0] zero-fill + scatter (= event time - kernel span), ms[1..4] extend-add / diag (+ small leaves) / panel / trailing
//   update: busy time summed over CTAs divided by the grid size (the classes and ms[5] = dependency wait add up to the
//   part of the kernel span a CTA spent working or spinning), ms[6] = kernel span, ms[7] = grid size.
extern "C" int mipm_ls_factorize_profile(mipm_handle hh, const double *d_nzval, double *ms, double *work, int64_t *launches)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze has not been called");
    if (!ms || !work || !launches) return fail(h, MIPM_ERR_ARG, "null argument");
    const LsSymbolic &S = h->sym;
    const bool want_fronts = std::getenv("MIPM_PHASE_LOG") != nullptr;
    if (want_fronts && !h->d_front_ns.p) MIPM_CUDA(h, h->d_front_ns.alloc((size_t)std::max(S.ns, 1)));
    const char *trace_path = std::getenv("MIPM_TASK_TRACE");
    if (trace_path && !h->d_trace.p) MIPM_CUDA(h, h->d_trace.alloc((size_t)4 * std::max(h->n_tasks, 1)));
    cudaEvent_t e0, e1;
    MIPM_CUDA(h, cudaEventCreate(&e0));
    MIPM_CUDA(h, cudaEventCreate(&e1));
    MIPM_CUDA(h, cudaMemsetAsync(h->d_prof.p, 0, (size_t)8 * h->grid_factor * sizeof(unsigned long long), h->stream));
    MIPM_CUDA(h, cudaEventRecord(e0, h->stream));
    int rc = ls_factorize_impl(h, d_nzval);
    if (rc != MIPM_OK) return rc;
    MIPM_CUDA(h, cudaEventRecord(e1, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    float total = 0.f;
    cudaEventElapsedTime(&total, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::vector<unsigned long long> pr((size_t)8 * h->grid_factor, 0);
    MIPM_CUDA(h, cudaMemcpy(pr.data(), h->d_prof.p, pr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int c = 0; c < 8; ++c) { ms[c] = 0.0; work[c] = 0.0; launches[c] = 0; }
    unsigned long long tmin = ~0ull, tmax = 0;
    double busy[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < h->grid_factor; ++b) {
        const unsigned long long *o = pr.data() + (size_t)8 * b;
        if (o[6] == 0) continue;
        for (int c = 0; c < 6; ++c) busy[c] += (double)o[c];
        tmin = std::min(tmin, o[6]);
        tmax = std::max(tmax, o[7]);
    }
    const double g = (double)h->grid_factor * 1e6;
    ms[1] = busy[T_EA] / g;
    ms[2] = (busy[T_DIAG] + busy[T_LEAF]) / g;
    ms[3] = busy[T_PANEL] / g;
    ms[4] = busy[T_TRAIL] / g;
    ms[5] = busy[5] / g;
    ms[6] = (tmax > tmin) ? (double)(tmax - tmin) * 1e-6 : 0.0;
    ms[7] = (double)h->grid_factor;
    ms[0] = std::max(0.0, (double)total - ms[6]);
    launches[0] = 4;
    launches[6] = 1;
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
    if (trace_path) {
        // one line per task: index, type, front, a, b, c, d, need, wait start, start, end (us from the kernel start), SM
        std::vector<unsigned long long> tr((size_t)4 * std::max(h->n_tasks, 1), 0);
        std::vector<Task> tk((size_t)std::max(h->n_tasks, 1));
        MIPM_CUDA(h, cudaMemcpy(tr.data(), h->d_trace.p, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        MIPM_CUDA(h, cudaMemcpy(tk.data(), h->d_tasks.p, (size_t)h->n_tasks * sizeof(Task), cudaMemcpyDeviceToHost));
        if (FILE *tf = std::fopen(trace_path, "w")) {
            std::fprintf(tf, "task,type,front,a,b,c,d,need,wait_us,start_us,end_us,sm\n");
            for (int t = 0; t < h->n_tasks; ++t) {
                const Task &q = tk[(size_t)t];
                const unsigned long long *o = tr.data() + 4 * (size_t)t;
                std::fprintf(tf, "%d,%d,%d,%d,%d,%d,%d,%d,%.2f,%.2f,%.2f,%llu\n", t, q.type, q.front, q.a, q.b, q.c, q.d, q.need,
                             (double)(o[0] - tmin) * 1e-3, (double)(o[1] - tmin) * 1e-3, (double)(o[2] - tmin) * 1e-3, o[3]);
            }
            std::fclose(tf);
        }
        h->d_trace.release();      // tracing is per profile call
    }
    if (want_fronts) {
        // completion time of the last front of every elimination-tree level, relative to the kernel start: the critical path
        std::vector<unsigned long long> fn((size_t)std::max(S.ns, 1), 0);
        MIPM_CUDA(h, cudaMemcpy(fn.data(), h->d_front_ns.p, (size_t)S.ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE *logf = std::fopen(std::getenv("MIPM_PHASE_LOG"), "w")) {
            std::fprintf(logf, "level,n_fronts,first_done_us,last_done_us\n");
            for (int l = 0; l < S.n_levels; ++l) {
                unsigned long long lo = ~0ull, hi = 0;
                for (int64_t t = S.level_ptr[(size_t)l]; t < S.level_ptr[(size_t)l + 1]; ++t) {
                    unsigned long long v = fn[(size_t)S.level_sn[(size_t)t]];
                    if (v == 0) continue;
                    lo = std::min(lo, v);
                    hi = std::max(hi, v);
                }
                std::fprintf(logf, "%d,%lld,%.2f,%.2f\n", l, (long long)(S.level_ptr[(size_t)l + 1] - S.level_ptr[(size_t)l]),
                             hi ? (double)(lo - tmin) * 1e-3 : 0.0, hi ? (double)(hi - tmin) * 1e-3 : 0.0);
            }
            std::fprintf(logf, "# classes (ms of CTA time / grid): ea %.4f diag %.4f panel %.4f trail %.4f wait %.4f | span %.4f grid %d tasks %d\n",
                         ms[1], ms[2], ms[3], ms[4], ms[5], ms[6], h->grid_factor, h->n_tasks);
            std::fclose(logf);
        }
    }
    return MIPM_OK;
}

extern "C" int mipm_bench_syrk(mipm_handle hh, int64_t n, int64_t k, double *d_C, int64_t ldc, const double *d_X, int64_t ldx)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n <= 0 || k <= 0 || !d_C || !d_X || ldc < n || ldx < n + 64 || (ldx & 1) || ((uintptr_t)d_X & 15))
        return fail(h, MIPM_ERR_ARG, "bad argument (X needs an even leading dimension >= n + 64 and 16-byte alignment)");
    DeviceInfo prop;
    if (device_info(h->device, prop) != MIPM_OK) return fail(h, MIPM_ERR_CUDA, "cudaGetDeviceProperties failed");
    MIPM_CUDA(h, cudaFuncSetAttribute(k_bench_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int64_t nt = (n + TILE - 1) / TILE;
    int64_t tiles = nt * (nt + 1) / 2;
    k_bench_syrk<<<(unsigned)std::min<int64_t>(tiles, (int64_t)prop.sm_count * 3), 256, SMEM_BYTES, h->stream>>>((int)n, (int)k, d_C, ldc, d_X, ldx);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}
