// Numeric supernodal multifrontal Cholesky / LDL^T, level-scheduled triangular solves and
// iterative refinement. Replaces cuDSS factorization / refactorization / solve as reached
// through MadNLP.factorize!(linear_solver) and MadNLP.solve!(linear_solver, x)
// (reference call sites: src/linear_solver.jl:10, src/KKT/normalkkt.jl:210).
//
// Data layout (all FP64, column-major):
//   panel of supernode s : (k+r) x k at L + lp[s], ld = k+r   (k columns, r rows below)
//   update matrix of s   : r x r     at U + up[s], ld = r      (lower triangle used)
//   LDL^T only: W + wp[s], (k+r) x NB scratch holding the unscaled panel block L*D
// Schedule: fronts are grouped by elimination-tree level; inside a level the dense partial
// factorization is blocked right-looking with block width NB=64; every (level, block step)
// is three batched launches (diagonal block, panel TRSM, trailing DMMA update) over all
// fronts of the level that still have columns left.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"

namespace mipm {

namespace {

constexpr int NB = 64;          // block width of the dense partial factorization
constexpr int LDS = 65;         // smem leading dimension for NB x NB blocks (odd: conflict-free rows)
constexpr int TRSM_ROWS = 128;  // rows per TRSM CTA
constexpr int EA_COLS = 32;     // parent-front columns per extend-add CTA
constexpr int TILE = 64;        // update tile
constexpr int KC = 32;          // K chunk staged in shared memory
constexpr int XS = 68;          // smem row stride of the staged operands (68 % 16 == 4: conflict-free DMMA loads)

struct Front {
    int k, r, N;
    int64_t lp, up;
};

__device__ __forceinline__ Front get_front(int s, const int32_t *sn_ptr, const int64_t *row_ptr,
                                           const int64_t *lp, const int64_t *up)
{
    Front f;
    f.k = sn_ptr[s + 1] - sn_ptr[s];
    f.r = (int)(row_ptr[s + 1] - row_ptr[s]);
    f.N = f.k + f.r;
    f.lp = lp[s];
    f.up = up[s];
    return f;
}

// largest i in [0, n) with prefix[i] <= x  (prefix[0] = 0, prefix[n] = total > x)
__device__ __forceinline__ int find_segment(const int32_t *prefix, int n, int x)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ assembly of the fronts
__global__ void __launch_bounds__(256)
k_scatter_a(int64_t nnz, const int64_t *__restrict__ a2l, const double *__restrict__ Ax, double *__restrict__ L)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) L[a2l[p]] = Ax[p];
}

// Extend-add: task = (parent s, parent-front columns [q0, q1)). The CTA walks the children of
// s in a fixed order and adds the part of each child's update matrix that lands in its column
// range, so every destination entry is owned by exactly one CTA -> deterministic sums.
__global__ void __launch_bounds__(256)
k_extend_add(const int32_t *__restrict__ tasks, const int32_t *__restrict__ sn_ptr,
             const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ lp, const int64_t *__restrict__ up,
             const int64_t *__restrict__ child_ptr, const int32_t *__restrict__ child_idx,
             const int32_t *__restrict__ rel_idx, double *__restrict__ L, double *__restrict__ U)
{
    const int s = tasks[3 * (int64_t)blockIdx.x], q0 = tasks[3 * (int64_t)blockIdx.x + 1], q1 = tasks[3 * (int64_t)blockIdx.x + 2];
    const Front f = get_front(s, sn_ptr, row_ptr, lp, up);
    double *P = L + f.lp;
    double *Us = U + f.up;
    for (int64_t ci = child_ptr[s]; ci < child_ptr[s + 1]; ++ci) {
        const int c = child_idx[ci];
        const int rc = (int)(row_ptr[c + 1] - row_ptr[c]);
        const int32_t *rel = rel_idx + row_ptr[c];
        const double *Uc = U + up[c];
        // first b with rel[b] >= q0 / q1 (rel is strictly increasing)
        int b0, b1;
        {
            int lo = 0, hi = rc;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (rel[mid] < q0) lo = mid + 1; else hi = mid; }
            b0 = lo;
            hi = rc;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (rel[mid] < q1) lo = mid + 1; else hi = mid; }
            b1 = lo;
        }
        for (int b = b0; b < b1; ++b) {
            const int tb = rel[b];
            const double *src = Uc + (int64_t)b * rc;
            if (tb < f.k) {
                double *dst = P + (int64_t)tb * f.N;
                for (int a = b + threadIdx.x; a < rc; a += blockDim.x) dst[rel[a]] += src[a];
            } else {
                double *dst = Us + (int64_t)(tb - f.k) * f.r - f.k;
                for (int a = b + threadIdx.x; a < rc; a += blockDim.x) dst[rel[a]] += src[a];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ diagonal block factorization
// One CTA per active front: factor the nb x nb diagonal block at column jb in shared memory.
// Cholesky: L11 L11' (a non-positive pivot sets info[0] and is replaced by 1 so the run stays
// finite; the host then retries with more regularization like src/linear_solver.jl:6-17).
// LDL^T: unit-lower L11 with D on the diagonal; |pivot| < piv_tol is replaced by +-piv_tol.
template <bool LDL>
__global__ void __launch_bounds__(256)
k_factor_diag(const int32_t *__restrict__ act, int jb, const int32_t *__restrict__ sn_ptr,
              const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ lp, double *__restrict__ L,
              int *__restrict__ info, double piv_tol)
{
    __shared__ double S[NB * LDS];
    __shared__ double diag[NB];
    const int s = act[blockIdx.x];
    const int k = sn_ptr[s + 1] - sn_ptr[s];
    const int N = k + (int)(row_ptr[s + 1] - row_ptr[s]);
    const int nb = min(NB, k - jb);
    double *P = L + lp[s] + (int64_t)jb * N + jb;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nb * nb; idx += 256) {
        int rr = idx % nb, cc = idx / nb;
        if (rr >= cc) S[cc * LDS + rr] = P[(int64_t)cc * N + rr];
    }
    __syncthreads();
    int nbad = 0, ntiny = 0;
    for (int j = 0; j < nb; ++j) {
        double d = S[j * LDS + j];
        double scale, dmul;
        if (!LDL) {
            if (!(d > 0.0) || !(d < 1.0e300)) { nbad++; d = 1.0; }
            double ljj = sqrt(d);
            scale = 1.0 / ljj;
            dmul = 1.0;
            if (tid == 0) diag[j] = ljj;
        } else {
            if (!(fabs(d) <= 1.0e300)) { nbad++; d = 1.0; }            // NaN / Inf
            else if (fabs(d) < piv_tol) { ntiny++; d = (d < 0.0) ? -piv_tol : piv_tol; }
            scale = 1.0 / d;
            dmul = d;
            if (tid == 0) diag[j] = d;
        }
        for (int rr = j + 1 + tid; rr < nb; rr += 256) S[j * LDS + rr] *= scale;
        __syncthreads();
        const int w = nb - j - 1;
        for (int idx = tid; idx < w * w; idx += 256) {
            int rr = j + 1 + idx % w, cc = j + 1 + idx / w;
            if (rr >= cc) S[cc * LDS + rr] -= S[j * LDS + rr] * (S[j * LDS + cc] * dmul);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nb * nb; idx += 256) {
        int rr = idx % nb, cc = idx / nb;
        if (rr > cc) P[(int64_t)cc * N + rr] = S[cc * LDS + rr];
        else if (rr == cc) P[(int64_t)cc * N + rr] = diag[cc];
    }
    if (tid == 0) {
        if (nbad) atomicMax(&info[0], 1);
        if (ntiny) atomicAdd(&info[2], ntiny);
        if (LDL) {
            int neg = 0;
            for (int j = 0; j < nb; ++j) neg += (diag[j] < 0.0);
            if (neg) atomicAdd(&info[1], neg);
        }
    }
}

// ------------------------------------------------------------------ panel TRSM
// Rows below the diagonal block: X = R * L11^-T (Cholesky) or X = R * L11^-T, Y = X * D^-1 (LDL^T).
// One thread per row, the row lives in registers, L11 is broadcast from shared memory.
template <bool LDL>
__global__ void __launch_bounds__(TRSM_ROWS)
k_trsm(const int32_t *__restrict__ act, const int32_t *__restrict__ prefix, int n_active, int jb,
       const int32_t *__restrict__ sn_ptr, const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ lp,
       const int64_t *__restrict__ wp, double *__restrict__ L, double *__restrict__ W)
{
    __shared__ double S[NB * LDS];
    const int fi = find_segment(prefix, n_active, (int)blockIdx.x);
    const int lc = (int)blockIdx.x - prefix[fi];
    const int s = act[fi];
    const int k = sn_ptr[s + 1] - sn_ptr[s];
    const int N = k + (int)(row_ptr[s + 1] - row_ptr[s]);
    const int nb = min(NB, k - jb);
    const int j1 = jb + nb;
    double *P = L + lp[s];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
        int rr = idx % NB, cc = idx / NB;
        double v = (rr == cc) ? 1.0 : 0.0;
        if (rr < nb && cc < nb && rr >= cc) v = P[(int64_t)(jb + cc) * N + jb + rr];
        S[cc * LDS + rr] = v;
    }
    __syncthreads();
    const int row = j1 + lc * TRSM_ROWS + tid;
    if (row >= N) return;
    double x[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? P[(int64_t)(jb + c) * N + row] : 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        if (!LDL) x[c] = x[c] / S[c * LDS + c];
#pragma unroll
        for (int c2 = c + 1; c2 < NB; ++c2) x[c2] = fma(-x[c], S[c * LDS + c2], x[c2]);
    }
    if (!LDL) {
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < nb) P[(int64_t)(jb + c) * N + row] = x[c];
    } else {
        double *Ws = W + wp[s];
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < nb) {
                Ws[(int64_t)c * N + row] = x[c];
                P[(int64_t)(jb + c) * N + row] = x[c] / S[c * LDS + c];
            }
    }
}

// ------------------------------------------------------------------ trailing update on FP64 tensor cores
// C(64x64 tile) -= X(rows, 0:nb) * Y(cols, 0:nb)'.  4 warps, each a 32x32 sub-tile made of 4x4
// DMMA m8n8k4 fragments; operands staged k-major in shared memory ([k][row], stride 68) so
// the staging copy is a straight coalesced column copy and the fragment loads are conflict-free.
// Only entries with (global row) >= (global col) are written.
__device__ __forceinline__ void tile_update(const double *__restrict__ X, int64_t ldx, const double *__restrict__ Y,
                                            int64_t ldy, int nb, int nrow, int ncol, double *__restrict__ C,
                                            int64_t ldc, int grow0, int gcol0, double (*Xs)[XS], double (*Ys)[XS])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, l3 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int k0 = 0; k0 < nb; k0 += KC) {
#pragma unroll
        for (int i = 0; i < (KC * TILE) / 128; ++i) {
            int idx = tid + i * 128;
            int rr = idx % TILE, kk = idx / TILE;
            bool kin = (k0 + kk) < nb;
            Xs[kk][rr] = (kin && rr < nrow) ? X[(int64_t)(k0 + kk) * ldx + rr] : 0.0;
            Ys[kk][rr] = (kin && rr < ncol) ? Y[(int64_t)(k0 + kk) * ldy + rr] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = Xs[kk + l3][wm * 32 + mi * 8 + g];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Ys[kk + l3][wn * 32 + ni * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int rr = wm * 32 + mi * 8 + g;
                int cc = wn * 32 + ni * 8 + l3 * 2 + e;
                if (rr < nrow && cc < ncol && (grow0 + rr) >= (gcol0 + cc)) {
                    double *p = C + (int64_t)cc * ldc + rr;
                    *p -= acc[mi][ni][e];
                }
            }
}

template <bool LDL>
__global__ void __launch_bounds__(128)
k_update(const int32_t *__restrict__ act, const int32_t *__restrict__ prefix, int n_active, int jb,
         const int32_t *__restrict__ sn_ptr, const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ lp,
         const int64_t *__restrict__ up, const int64_t *__restrict__ wp, double *__restrict__ L,
         double *__restrict__ U, const double *__restrict__ W)
{
    __shared__ double Xs[KC][XS];
    __shared__ double Ys[KC][XS];
    const int fi = find_segment(prefix, n_active, (int)blockIdx.x);
    const int lt = (int)blockIdx.x - prefix[fi];
    const int s = act[fi];
    const Front f = get_front(s, sn_ptr, row_ptr, lp, up);
    const int nb = min(NB, f.k - jb);
    const int j1 = jb + nb;
    const int nt1 = (f.k > j1) ? (f.k - j1 + TILE - 1) / TILE : 0;
    int tr = (int)((sqrt(8.0 * (double)lt + 1.0) - 1.0) * 0.5);
    while (tr * (tr + 1) / 2 > lt) --tr;
    while ((tr + 1) * (tr + 2) / 2 <= lt) ++tr;
    const int tc = lt - tr * (tr + 1) / 2;
    int row0, rend, col0, cend;
    if (tr < nt1) { row0 = j1 + TILE * tr; rend = min(row0 + TILE, f.k); }
    else { row0 = f.k + TILE * (tr - nt1); rend = min(row0 + TILE, f.N); }
    if (tc < nt1) { col0 = j1 + TILE * tc; cend = min(col0 + TILE, f.k); }
    else { col0 = f.k + TILE * (tc - nt1); cend = min(col0 + TILE, f.N); }
    double *P = L + f.lp;
    const double *Y = P + (int64_t)jb * f.N + col0;
    const double *X = LDL ? (W + wp[s] + row0) : (P + (int64_t)jb * f.N + row0);
    double *C;
    int64_t ldc;
    if (col0 < f.k) { C = P + (int64_t)col0 * f.N + row0; ldc = f.N; }
    else { C = U + f.up + (int64_t)(col0 - f.k) * f.r + (row0 - f.k); ldc = f.r; }
    tile_update(X, f.N, Y, f.N, nb, rend - row0, cend - col0, C, ldc, row0, col0, Xs, Ys);
}

// Micro-benchmark hook for the same tile kernel: C (n x n, lower tiles) -= X X'.
__global__ void __launch_bounds__(128)
k_bench_syrk(int n, int kdim, double *__restrict__ C, int64_t ldc, const double *__restrict__ X, int64_t ldx)
{
    __shared__ double Xs[KC][XS];
    __shared__ double Ys[KC][XS];
    const int lt = blockIdx.x;
    int tr = (int)((sqrt(8.0 * (double)lt + 1.0) - 1.0) * 0.5);
    while (tr * (tr + 1) / 2 > lt) --tr;
    while ((tr + 1) * (tr + 2) / 2 <= lt) ++tr;
    const int tc = lt - tr * (tr + 1) / 2;
    const int row0 = tr * TILE, col0 = tc * TILE;
    for (int k0 = 0; k0 < kdim; k0 += NB)
        tile_update(X + (int64_t)k0 * ldx + row0, ldx, X + (int64_t)k0 * ldx + col0, ldx, min(NB, kdim - k0),
                    min(TILE, n - row0), min(TILE, n - col0), C + (int64_t)col0 * ldc + row0, ldc, row0, col0, Xs, Ys);
}

// ------------------------------------------------------------------ triangular solves
__global__ void __launch_bounds__(256)
k_gather_perm(int64_t n, const int32_t *__restrict__ perm, const double *__restrict__ b, double *__restrict__ xp)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xp[i] = b[perm[i]];
}
__global__ void __launch_bounds__(256)
k_scatter_perm(int64_t n, const int32_t *__restrict__ perm, const double *__restrict__ xp, double *__restrict__ x, int accumulate)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { if (accumulate) x[perm[i]] += xp[i]; else x[perm[i]] = xp[i]; }
}

// Forward substitution for all fronts of one level; one CTA per front.
//   1. add the children's update vectors (fixed child order),
//   2. y1 = L11^-1 x1 blocked by NB with the diagonal block in shared memory,
//   3. x1' / u -= L21 y1 for the rows below each block.
template <bool LDL>
__global__ void __launch_bounds__(256)
k_solve_fwd(const int32_t *__restrict__ fronts, const int32_t *__restrict__ sn_ptr, const int64_t *__restrict__ row_ptr,
            const int64_t *__restrict__ lp, const int64_t *__restrict__ child_ptr, const int32_t *__restrict__ child_idx,
            const int32_t *__restrict__ rel_idx, const double *__restrict__ L, double *__restrict__ xp,
            double *__restrict__ uvec)
{
    __shared__ double S[NB * LDS];
    __shared__ double xb[NB];
    const int s = fronts[blockIdx.x];
    const int c0 = sn_ptr[s];
    const int k = sn_ptr[s + 1] - c0;
    const int r = (int)(row_ptr[s + 1] - row_ptr[s]);
    const int N = k + r;
    const double *P = L + lp[s];
    double *x1 = xp + c0;
    double *u = uvec + row_ptr[s];
    const int tid = threadIdx.x;
    for (int64_t ci = child_ptr[s]; ci < child_ptr[s + 1]; ++ci) {
        const int c = child_idx[ci];
        const int rc = (int)(row_ptr[c + 1] - row_ptr[c]);
        const int32_t *rel = rel_idx + row_ptr[c];
        const double *uc = uvec + row_ptr[c];
        for (int a = tid; a < rc; a += 256) {
            int t = rel[a];
            if (t < k) x1[t] += uc[a]; else u[t - k] += uc[a];
        }
        __syncthreads();
    }
    for (int jb = 0; jb < k; jb += NB) {
        const int nb = min(NB, k - jb);
        for (int idx = tid; idx < nb * nb; idx += 256) {
            int rr = idx % nb, cc = idx / nb;
            if (rr >= cc) S[cc * LDS + rr] = P[(int64_t)(jb + cc) * N + jb + rr];
        }
        if (tid < nb) xb[tid] = x1[jb + tid];
        __syncthreads();
        if (tid < 32) {
            for (int j = 0; j < nb; ++j) {
                double xj = xb[j];
                if (!LDL) xj = xj / S[j * LDS + j];
                __syncwarp();
                if (tid == 0) xb[j] = xj;
                for (int i = j + 1 + tid; i < nb; i += 32) xb[i] -= S[j * LDS + i] * xj;
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < nb) x1[jb + tid] = xb[tid];
        for (int i = jb + nb + tid; i < N; i += 256) {
            double acc = 0.0;
            const double *col = P + (int64_t)jb * N + i;
            for (int j = 0; j < nb; ++j) acc = fma(col[(int64_t)j * N], xb[j], acc);
            if (i < k) x1[i] -= acc; else u[i - k] -= acc;
        }
        __syncthreads();
    }
}

// Backward substitution for all fronts of one level (levels processed from the root down).
//   x1 = L11^-T (y1 [/ D] - L21' x_anc), blocked from the last block to the first.
template <bool LDL>
__global__ void __launch_bounds__(256)
k_solve_bwd(const int32_t *__restrict__ fronts, const int32_t *__restrict__ sn_ptr, const int64_t *__restrict__ row_ptr,
            const int64_t *__restrict__ lp, const int32_t *__restrict__ row_idx, const double *__restrict__ L,
            double *__restrict__ xp)
{
    __shared__ double S[NB * LDS];
    __shared__ double xb[NB];
    const int s = fronts[blockIdx.x];
    const int c0 = sn_ptr[s];
    const int k = sn_ptr[s + 1] - c0;
    const int r = (int)(row_ptr[s + 1] - row_ptr[s]);
    const int N = k + r;
    const double *P = L + lp[s];
    double *x1 = xp + c0;
    const int32_t *rows = row_idx + row_ptr[s];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (k + NB - 1) / NB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int jb = b * NB;
        const int nb = min(NB, k - jb);
        for (int idx = tid; idx < nb * nb; idx += 256) {
            int rr = idx % nb, cc = idx / nb;
            if (rr >= cc) S[cc * LDS + rr] = P[(int64_t)(jb + cc) * N + jb + rr];
        }
        // w[q] = y[q] (/ D[q]) - sum_{i >= jb+nb} L[i][q] * xfull[i]; one warp per column q
        for (int q = warp; q < nb; q += 8) {
            const double *col = P + (int64_t)(jb + q) * N;
            double acc = 0.0;
            for (int i = jb + nb + lane; i < N; i += 32) {
                double xv = (i < k) ? x1[i] : xp[rows[i - k]];
                acc = fma(col[i], xv, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                double y = x1[jb + q];
                if (LDL) y = y / col[jb + q];
                xb[q] = y - acc;
            }
        }
        __syncthreads();
        if (tid < 32) {
            for (int j = nb - 1; j >= 0; --j) {
                double xj = xb[j];
                if (!LDL) xj = xj / S[j * LDS + j];
                __syncwarp();
                if (tid == 0) xb[j] = xj;
                for (int i = tid; i < j; i += 32) xb[i] -= S[i * LDS + j] * xj;
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < nb) x1[jb + tid] = xb[tid];
        __syncthreads();
    }
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
    for (int64_t p = ptr[row] + lane; p < ptr[row + 1]; p += 32) acc = fma(__ldg(val + vpos[p]), __ldg(x + col[p]), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rout[row] = b[row] - acc;
}

inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)std::max<int64_t>(1, (n + per_block - 1) / per_block); }

}  // namespace

// ---------------------------------------------------------------------------------------
int ls_device_setup(Handle *h)
{
    const LsSymbolic &S = h->sym;
    const int ns = S.ns;
    MIPM_CUDA(h, cudaSetDevice(h->device));
    // ---- build the schedule
    std::vector<int32_t> sched;
    h->steps.clear();
    h->levels.assign((size_t)S.n_levels, LevelInfo());
    std::vector<int64_t> wp((size_t)ns + 1, 0);
    for (int s = 0; s < ns; ++s) {
        int64_t k = S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s];
        int64_t r = S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s];
        wp[(size_t)s + 1] = wp[(size_t)s] + (S.kind == MIPM_LDL ? (k + r) * NB : 0);
    }
    int64_t n_launch = 3;  // two memsets + scatter
    for (int l = 0; l < S.n_levels; ++l) {
        LevelInfo &li = h->levels[(size_t)l];
        const int64_t f0 = S.level_ptr[(size_t)l], f1 = S.level_ptr[(size_t)l + 1];
        li.off_all = (int64_t)sched.size();
        li.n_all = (int32_t)(f1 - f0);
        int kmax = 0;
        for (int64_t t = f0; t < f1; ++t) {
            int s = S.level_sn[(size_t)t];
            sched.push_back(s);
            kmax = std::max(kmax, S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s]);
        }
        li.off_parents = (int64_t)sched.size();
        li.n_parents = 0;
        for (int64_t t = f0; t < f1; ++t) {
            int s = S.level_sn[(size_t)t];
            if (S.child_ptr[(size_t)s + 1] > S.child_ptr[(size_t)s]) { sched.push_back(s); li.n_parents++; }
        }
        li.off_ea_tasks = (int64_t)sched.size();
        li.n_ea_tasks = 0;
        for (int64_t t = 0; t < li.n_parents; ++t) {
            int s = sched[(size_t)(li.off_parents + t)];
            int N = (S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s]) + (int)(S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s]);
            for (int q0 = 0; q0 < N; q0 += EA_COLS) {
                sched.push_back(s);
                sched.push_back(q0);
                sched.push_back(std::min(N, q0 + EA_COLS));
                li.n_ea_tasks++;
            }
        }
        if (li.n_ea_tasks) n_launch++;
        for (int jb = 0; jb < kmax; jb += NB) {
            FactorStep st;
            st.level = l;
            st.jb = jb;
            st.off_sn = (int64_t)sched.size();
            st.n_active = 0;
            for (int64_t t = f0; t < f1; ++t) {
                int s = S.level_sn[(size_t)t];
                if (S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s] > jb) { sched.push_back(s); st.n_active++; }
            }
            st.off_trsm = (int64_t)sched.size();
            sched.resize(sched.size() + (size_t)st.n_active + 1);
            st.off_upd = (int64_t)sched.size();
            sched.resize(sched.size() + (size_t)st.n_active + 1);
            int64_t nt = 0, nu = 0;
            for (int t = 0; t < st.n_active; ++t) {
                int s = sched[(size_t)(st.off_sn + t)];
                int k = S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s];
                int r = (int)(S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s]);
                int nb = std::min(NB, k - jb), j1 = jb + nb, N = k + r;
                sched[(size_t)(st.off_trsm + t)] = (int32_t)nt;
                sched[(size_t)(st.off_upd + t)] = (int32_t)nu;
                nt += (N - j1 + TRSM_ROWS - 1) / TRSM_ROWS;
                int64_t nt1 = (k > j1) ? (k - j1 + TILE - 1) / TILE : 0, nt2 = (r + TILE - 1) / TILE;
                int64_t ntl = nt1 + nt2;
                nu += ntl * (ntl + 1) / 2;
                if (nu > INT32_MAX || nt > INT32_MAX) return fail(h, MIPM_ERR_ARG, "front too large for the 32-bit tile schedule");
            }
            sched[(size_t)(st.off_trsm + st.n_active)] = (int32_t)nt;
            sched[(size_t)(st.off_upd + st.n_active)] = (int32_t)nu;
            st.n_trsm = nt;
            st.n_upd = nu;
            n_launch += 1 + (nt > 0) + (nu > 0);
            h->steps.push_back(st);
        }
    }
    h->n_launch_factor = n_launch;
    // ---- uploads and workspaces
    cudaStream_t st = h->stream;
    MIPM_CUDA(h, h->d_sched.upload(sched, st));
    MIPM_CUDA(h, h->d_sn_ptr.upload(S.sn_ptr, st));
    MIPM_CUDA(h, h->d_sn_parent.upload(S.sn_parent, st));
    MIPM_CUDA(h, h->d_row_ptr.upload(S.row_ptr, st));
    MIPM_CUDA(h, h->d_row_idx.upload(S.row_idx, st));
    MIPM_CUDA(h, h->d_rel_idx.upload(S.rel_idx, st));
    MIPM_CUDA(h, h->d_perm.upload(S.perm, st));
    MIPM_CUDA(h, h->d_lp.upload(S.lp, st));
    MIPM_CUDA(h, h->d_up.upload(S.up, st));
    MIPM_CUDA(h, h->d_wp.upload(wp, st));
    MIPM_CUDA(h, h->d_child_ptr.upload(S.child_ptr, st));
    MIPM_CUDA(h, h->d_child_idx.upload(S.child_idx, st));
    MIPM_CUDA(h, h->d_a2l.upload(S.a2l, st));
    MIPM_CUDA(h, h->d_full_ptr.upload(S.full_ptr, st));
    MIPM_CUDA(h, h->d_full_col.upload(S.full_col, st));
    MIPM_CUDA(h, h->d_full_val.upload(S.full_val, st));
    MIPM_CUDA(h, h->d_L.alloc((size_t)std::max<int64_t>(S.nnz_l, 1)));
    MIPM_CUDA(h, h->d_U.alloc((size_t)std::max<int64_t>(S.update_doubles, 1)));
    MIPM_CUDA(h, h->d_W.alloc((size_t)std::max<int64_t>(wp[(size_t)ns], 1)));
    MIPM_CUDA(h, h->d_xp.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_uvec.alloc((size_t)std::max<int64_t>(S.row_ptr[(size_t)ns], 1)));
    MIPM_CUDA(h, h->d_b.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, h->d_r.alloc((size_t)std::max<int64_t>(S.n, 1)));
    MIPM_CUDA(h, cudaStreamSynchronize(st));
    h->factorized = false;
    return MIPM_OK;
}

// Optional per-class device timing of one factorization (events around every launch).
struct FactorProfile {
    struct Rec { cudaEvent_t a, b; int cls; };
    std::vector<Rec> recs;
};
static FactorProfile *g_prof = nullptr;   // only set inside mipm_ls_factorize_profile (single-threaded per handle)
#define PROF_BEGIN(cls_)                                                  \
    FactorProfile::Rec rec__;                                             \
    if (g_prof) {                                                         \
        cudaEventCreate(&rec__.a);                                        \
        cudaEventCreate(&rec__.b);                                        \
        rec__.cls = (cls_);                                               \
        cudaEventRecord(rec__.a, st);                                     \
    }
#define PROF_END()                                                        \
    if (g_prof) {                                                         \
        cudaEventRecord(rec__.b, st);                                     \
        g_prof->recs.push_back(rec__);                                    \
    }

template <bool LDL>
static int factorize_t(Handle *h, const double *d_nzval)
{
    const LsSymbolic &S = h->sym;
    cudaStream_t st = h->stream;
    const int32_t *sched = h->d_sched.p;
    {
        PROF_BEGIN(0);
        MIPM_CUDA(h, cudaMemsetAsync(h->d_L.p, 0, (size_t)std::max<int64_t>(S.nnz_l, 1) * sizeof(double), st));
        MIPM_CUDA(h, cudaMemsetAsync(h->d_U.p, 0, (size_t)std::max<int64_t>(S.update_doubles, 1) * sizeof(double), st));
        MIPM_CUDA(h, cudaMemsetAsync(h->d_info.p, 0, 4 * sizeof(int), st));
        if (S.nnz_a > 0) {
            k_scatter_a<<<grid_for(S.nnz_a, 256), 256, 0, st>>>(S.nnz_a, h->d_a2l.p, d_nzval, h->d_L.p);
            MIPM_CHECK_LAUNCH(h);
        }
        PROF_END();
    }
    // pivot tolerance for LDL^T: relative to nothing we can see cheaply -> absolute, tiny
    const double piv_tol = 1e-13;
    size_t si = 0;
    for (int l = 0; l < S.n_levels; ++l) {
        const LevelInfo &li = h->levels[(size_t)l];
        if (li.n_ea_tasks > 0) {
            PROF_BEGIN(1);
            k_extend_add<<<(unsigned)li.n_ea_tasks, 256, 0, st>>>(sched + li.off_ea_tasks, h->d_sn_ptr.p, h->d_row_ptr.p,
                                                                h->d_lp.p, h->d_up.p, h->d_child_ptr.p, h->d_child_idx.p,
                                                                h->d_rel_idx.p, h->d_L.p, h->d_U.p);
            MIPM_CHECK_LAUNCH(h);
            PROF_END();
        }
        for (; si < h->steps.size() && h->steps[si].level == l; ++si) {
            const FactorStep &fs = h->steps[si];
            {
            PROF_BEGIN(2);
            k_factor_diag<LDL><<<(unsigned)fs.n_active, 256, 0, st>>>(sched + fs.off_sn, fs.jb, h->d_sn_ptr.p, h->d_row_ptr.p,
                                                                     h->d_lp.p, h->d_L.p, h->d_info.p, piv_tol);
            MIPM_CHECK_LAUNCH(h);
            PROF_END();
            }
            if (fs.n_trsm > 0) {
                PROF_BEGIN(3);
                k_trsm<LDL><<<(unsigned)fs.n_trsm, TRSM_ROWS, 0, st>>>(sched + fs.off_sn, sched + fs.off_trsm, fs.n_active, fs.jb,
                                                                      h->d_sn_ptr.p, h->d_row_ptr.p, h->d_lp.p, h->d_wp.p,
                                                                      h->d_L.p, h->d_W.p);
                MIPM_CHECK_LAUNCH(h);
                PROF_END();
            }
            if (fs.n_upd > 0) {
                PROF_BEGIN(4);
                k_update<LDL><<<(unsigned)fs.n_upd, 128, 0, st>>>(sched + fs.off_sn, sched + fs.off_upd, fs.n_active, fs.jb,
                                                                 h->d_sn_ptr.p, h->d_row_ptr.p, h->d_lp.p, h->d_up.p, h->d_wp.p,
                                                                 h->d_L.p, h->d_U.p, h->d_W.p);
                MIPM_CHECK_LAUNCH(h);
                PROF_END();
            }
        }
    }
    h->d_nzval = d_nzval;
    h->factorized = true;
    return MIPM_OK;
}

int ls_factorize_impl(Handle *h, const double *d_nzval)
{
    MIPM_CUDA(h, cudaSetDevice(h->device));
    return h->sym.kind == MIPM_LDL ? factorize_t<true>(h, d_nzval) : factorize_t<false>(h, d_nzval);
}

template <bool LDL>
static int solve_permuted(Handle *h)
{
    // solves in place on h->d_xp (permuted numbering)
    const LsSymbolic &S = h->sym;
    cudaStream_t st = h->stream;
    const int32_t *sched = h->d_sched.p;
    MIPM_CUDA(h, cudaMemsetAsync(h->d_uvec.p, 0, (size_t)std::max<int64_t>(S.row_ptr[(size_t)S.ns], 1) * sizeof(double), st));
    for (int l = 0; l < S.n_levels; ++l) {
        const LevelInfo &li = h->levels[(size_t)l];
        k_solve_fwd<LDL><<<(unsigned)li.n_all, 256, 0, st>>>(sched + li.off_all, h->d_sn_ptr.p, h->d_row_ptr.p, h->d_lp.p,
                                                            h->d_child_ptr.p, h->d_child_idx.p, h->d_rel_idx.p, h->d_L.p,
                                                            h->d_xp.p, h->d_uvec.p);
        MIPM_CHECK_LAUNCH(h);
    }
    for (int l = S.n_levels - 1; l >= 0; --l) {
        const LevelInfo &li = h->levels[(size_t)l];
        k_solve_bwd<LDL><<<(unsigned)li.n_all, 256, 0, st>>>(sched + li.off_all, h->d_sn_ptr.p, h->d_row_ptr.p, h->d_lp.p,
                                                            h->d_row_idx.p, h->d_L.p, h->d_xp.p);
        MIPM_CHECK_LAUNCH(h);
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
    const bool ldl = S.kind == MIPM_LDL;
    if (ir_steps > 0) MIPM_CUDA(h, cudaMemcpyAsync(h->d_b.p, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_gather_perm<<<grid_for(n, 256), 256, 0, st>>>(n, h->d_perm.p, d_x, h->d_xp.p);
    MIPM_CHECK_LAUNCH(h);
    int rc = ldl ? solve_permuted<true>(h) : solve_permuted<false>(h);
    if (rc != MIPM_OK) return rc;
    k_scatter_perm<<<grid_for(n, 256), 256, 0, st>>>(n, h->d_perm.p, h->d_xp.p, d_x, 0);
    MIPM_CHECK_LAUNCH(h);
    for (int it = 0; it < ir_steps; ++it) {
        k_sym_residual<<<grid_for(n * 32, 256), 256, 0, st>>>(n, h->d_full_ptr.p, h->d_full_col.p, h->d_full_val.p, h->d_nzval,
                                                            d_x, h->d_b.p, h->d_r.p);
        MIPM_CHECK_LAUNCH(h);
        k_gather_perm<<<grid_for(n, 256), 256, 0, st>>>(n, h->d_perm.p, h->d_r.p, h->d_xp.p);
        MIPM_CHECK_LAUNCH(h);
        rc = ldl ? solve_permuted<true>(h) : solve_permuted<false>(h);
        if (rc != MIPM_OK) return rc;
        k_scatter_perm<<<grid_for(n, 256), 256, 0, st>>>(n, h->d_perm.p, h->d_xp.p, d_x, 1);
        MIPM_CHECK_LAUNCH(h);
    }
    return MIPM_OK;
}

}  // namespace mipm

extern "C" int mipm_ls_factorize_profile(mipm_handle hh, const double *d_nzval, double *ms, double *work, int64_t *launches)
{
    using namespace mipm;
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze has not been called");
    if (!ms || !work || !launches) return fail(h, MIPM_ERR_ARG, "null argument");
    const LsSymbolic &S = h->sym;
    FactorProfile prof;
    g_prof = &prof;
    int rc = ls_factorize_impl(h, d_nzval);
    g_prof = nullptr;
    if (rc != MIPM_OK) return rc;
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int c = 0; c < 5; ++c) { ms[c] = 0.0; work[c] = 0.0; launches[c] = 0; }
    for (auto &r : prof.recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.cls] += t;
        launches[r.cls] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    // algorithmic work per class: bytes for 0 (zero-fill + scatter) and 1 (extend-add), flops for 2..4
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
    k_bench_syrk<<<(unsigned)tiles, 128, 0, h->stream>>>((int)n, (int)k, d_C, ldc, d_X, ldx);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}
