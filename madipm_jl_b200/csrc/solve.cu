// Level-scheduled supernodal triangular solves and iterative refinement. Replaces cuDSS's solve phase as reached
// through MadNLP.solve!(linear_solver, x) (reference call site: src/KKT/normalkkt.jl:210). The factor comes from factor.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "front.cuh"

namespace mipm {

namespace {

struct SolveParams {
    const FrontInfo *fi;
    const int32_t *child_idx, *rel_idx, *row_idx, *perm;
    const int32_t *leaves;      // small leaf fronts, one warp each
    int n_leaf;
    const int4 *tasks;          // (kind, id, part, 0): 0 forward front, 1 forward leaf group, 2 backward front, 3 backward leaf
                                // group, 4 forward row tile `part` of a big front, 5 backward column block `part` of a big front
    int task_begin, task_end;
    int root_sn, root_mode;     // staged (distributed) solves: front_forward mode of the border root (0 unless staged)
    int *fprog;                 // per front: children whose forward step is complete
    int *bdone;                 // per front: backward step complete
    int *yprog, *xprog, *ftiles; // big fronts: forward blocks published, backward blocks published, forward tiles finished
    int *ticket;
    const double *L, *Dinv;
    double *xp, *uvec;
    const double *b_in;         // right-hand side (original numbering), read through perm by the front that owns the entry
    double *x_out;              // solution (original numbering), written through perm by the front that owns the entry
    int accumulate;             // x_out[perm] += x instead of =
    // fronts with many children (K2: one tiny leaf child per primal variable) fold their children's update vectors in
    // through a transposed map: per destination row of the front, the update-vector slots that land on it, in child
    // order. gat_off[2 s] = offset of the front's N + 1 pointers in gat_ptr (-1: walk the children instead),
    // gat_off[2 s + 1] = offset of its source list in gat_src.
    const int64_t *gat_off;
    const int32_t *gat_ptr, *gat_src;
    unsigned long long *trace;  // optional (MIPM_SOLVE_TRACE): 3 per task: wait start, start, end
};

// ------------------------------------------------------------------ triangular solves
// One persistent launch that executes a task list: one task per front and sweep, forward tasks in elimination-tree
// level order, then backward tasks from the root down. CTAs draw tasks in list order from a ticket counter; a forward
// task waits until all children of its front have signalled (per-front counter), a backward task until its parent has
// (per-front flag). No grid-wide barrier anywhere (the round-1 kernel had one per level and sweep), and a CTA that
// waits first prefetches its front's panel into L2. A front reads its part of the right-hand side through perm and
// writes its part of the solution through perm itself. The 64 x 64 diagonal blocks are applied through their stored
// inverses (mat-vec), so nothing in a front is sequential; children's update vectors are summed by the parent in a
// fixed order (deterministic).
constexpr int XR_MAX = 1536;    // ancestor entries of x cached in shared memory by the backward sweep

// L2 prefetch of a front's panel and inverted diagonal blocks. The solves are latency-bound near the root: a front is
// a chain of dependent global loads, and the factor (hundreds of MB) does not stay in L2. A CTA whose front still waits
// for its children (or its parent) fetches the panel first, so the dependent loads hit L2 instead of HBM.
__device__ __forceinline__ void prefetch_front(const SolveParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const char *base = reinterpret_cast<const char *>(p.L + f.lp);
    const int64_t bytes = (int64_t)front_ld(f.k, f.r) * f.k * 8;
    for (int64_t off = (int64_t)threadIdx.x * 128; off < bytes; off += 256 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    const char *dv = reinterpret_cast<const char *>(p.Dinv + f.dinv * (int64_t)(NB * XS));
    const int64_t bytes2 = (int64_t)((f.k + NB - 1) / NB) * NB * XS * 8;
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
    const int k = f.k, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    double *u = p.uvec + f.rowp;
    const int tid = threadIdx.x;
    const int64_t go = p.gat_off[2 * (int64_t)s];
    if (mode != 2) {            // this front owns x1 and u: right-hand side through perm, empty update vector
        for (int t = tid; t < N; t += 256) {
            if (t < k) x1[t] = p.b_in[p.perm[f.c0 + t]]; else u[t - k] = 0.0;
        }
        __syncthreads();
    }
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
        const double *Dv = p.Dinv + (f.dinv + (jb >> 6)) * (int64_t)(NB * XS);
        if (tid < NB) xb[tid] = (tid < nb) ? x1[jb + tid] : 0.0;
        __syncthreads();
        {   // y = inv(L11 block) * xb : row rr by the 4 threads (rr, q), columns pp = q, q+4, ...
            const int rr = tid >> 2, q = tid & 3;
            double acc = 0.0;
            double dvv[NB / 4];                  // the 16 loads of this row first (entries above the diagonal are stored zeros)
#pragma unroll
            for (int t = 0; t < NB / 4; ++t) dvv[t] = Dv[(q + 4 * t) * XS + rr];
#pragma unroll
            for (int t = 0; t < NB / 4; ++t) acc = fma(dvv[t], xb[q + 4 * t], acc);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (q == 0) { yb[rr] = acc; if (rr < nb) x1[jb + rr] = acc; }
        }
        __syncthreads();
        for (int i = jb + nb + tid; i < N; i += 256) {
            // FWD_ILP independent loads in flight per thread (the panel lives in HBM: latency-bound otherwise); the last
            // trip is masked instead of walking the remaining columns one dependent load at a time (yb is zero beyond nb)
            constexpr int FWD_ILP = 16;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const double *col = P + (int64_t)jb * ld + i;
#pragma unroll 1
            for (int j = 0; j < nb; j += FWD_ILP) {
                double v[FWD_ILP];
#pragma unroll
                for (int t = 0; t < FWD_ILP; ++t) v[t] = (j + t < nb) ? col[(int64_t)(j + t) * ld] : 0.0;
#pragma unroll
                for (int t = 0; t < FWD_ILP; t += 4) {
                    a0 = fma(v[t], yb[j + t], a0);
                    a1 = fma(v[t + 1], yb[j + t + 1], a1);
                    a2 = fma(v[t + 2], yb[j + t + 2], a2);
                    a3 = fma(v[t + 3], yb[j + t + 3], a3);
                }
            }
            const double tot = (a0 + a1) + (a2 + a3);
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
    const int k = f.k, r = f.r, N = f.k + f.r, ld = front_ld(f.k, f.r);
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
        const double *Dv = p.Dinv + (f.dinv + b) * (int64_t)(NB * XS);
        for (int idx = tid; idx < NB * NB; idx += 256) S[(idx >> 6) * LDS + (idx & 63)] = Dv[(idx >> 6) * XS + (idx & 63)];
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
                const double *col = P + (int64_t)(jb + q0) * ld;
                int i = jb + nb + lane;
                for (; i + 32 < N; i += 64) {     // two row-chunks per trip: 16 loads in flight per lane
                    const int i2 = i + 32;
                    double xv = (i < k) ? x1[i] : (cached ? xr[i - k] : p.xp[rows[i - k]]);
                    double xw = (i2 < k) ? x1[i2] : (cached ? xr[i2 - k] : p.xp[rows[i2 - k]]);
                    double v[8], w[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const bool on = q0 + c < nb;
                        v[c] = on ? col[(int64_t)c * ld + i] : 0.0;
                        w[c] = on ? col[(int64_t)c * ld + i2] : 0.0;
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] = fma(w[c], xw, fma(v[c], xv, acc[c]));
                }
                for (; i < N; i += 32) {
                    double xv = (i < k) ? x1[i] : (cached ? xr[i - k] : p.xp[rows[i - k]]);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (q0 + c < nb) acc[c] = fma(col[(int64_t)c * ld + i], xv, acc[c]);
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
                    if (LDL) y = y / col[(int64_t)lane * ld + jb + q0 + lane];
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
            if (q == 0 && cc < nb) {
                x1[jb + cc] = acc;
                double *o = p.x_out + p.perm[f.c0 + jb + cc];
                *o = p.accumulate ? *o + acc : acc;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ big fronts: one front, many CTAs
// A front with thousands of columns (the root separator of a mesh-like problem, the border root of a block-angular
// one) is a chain of 64-column blocks; handled by ONE CTA it streams its whole panel through one SM (9 ms per solve on
// the mesh variant of C2). Here the forward sweep is split by 64-row tiles and the backward sweep by 64-column blocks:
// tile t owns rows [64 t, 64 t + 64) of the front, accumulates  x_t - sum_{b < t} L[t, b] y_b  as the y_b are published
// (per-front counter yprog, acquire spin) and, if it sits on the diagonal, solves its own block and publishes y_t.
// The backward task of block b waits for the blocks above it (xprog) and publishes x_b. The critical path per block is
// one 64 x 64 product + the inverse-block product instead of a whole panel column block.
constexpr int BIG_PART = 4;     // quarters of a 64-column block handled by the 4 x 64 thread layout

__device__ __forceinline__ int ld_acquire_b(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// acc[r] -= sum_{c < nc} P[(col0 + c) * ld + row0 + r] * yb[c] for the tile's rows r0 <= r < nr; thread (r, q) takes the
// columns q, q + 4, ...; all loads of a thread are issued before the first use
__device__ __forceinline__ void tile_update(const double *P, int64_t ld, int row0, int r0, int nr, int col0, int nc, const double *yb,
                                            double *acc, double *part)
{
    const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
    double sum = 0.0;
    if (r >= r0 && r < nr) {
        const double *col = P + (int64_t)col0 * ld + row0 + r;
        double v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) { const int c = q + BIG_PART * t; v[t] = (c < nc) ? col[(int64_t)c * ld] : 0.0; }
#pragma unroll
        for (int t = 0; t < 16; ++t) sum = fma(v[t], yb[q + BIG_PART * t], sum);
    }
    part[q * 64 + r] = sum;
    __syncthreads();
    if (threadIdx.x < 64 && r >= r0 && r < nr) acc[r] -= (part[r] + part[64 + r]) + (part[128 + r] + part[192 + r]);
    __syncthreads();
}

// L2 prefetch of what one big-front task will read, issued before the task waits for its dependencies
__device__ __forceinline__ void prefetch_big(const SolveParams &p, const FrontInfo &f, bool fwd, int part)
{
    const int k = f.k, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const char *P = reinterpret_cast<const char *>(p.L + f.lp);
    const int nblk = (k + NB - 1) / NB;
    if (fwd) {                  // rows [64 part, +64) of the columns before (and of) the tile's diagonal block
        const int ncol = min(k, (part + 1) * NB);
        const int64_t row_off = (int64_t)part * 64 * 8;
        for (int i = threadIdx.x; i < ncol * 4; i += 256)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P + ((int64_t)(i >> 2) * ld) * 8 + row_off + (i & 3) * 128));
    } else {                    // columns of block `part`, rows from the block down
        const int jb = part * NB, nc = min(NB, k - jb);
        const int64_t col_bytes = (int64_t)(N - jb) * 8;
        const int lines = (int)((col_bytes + 127) >> 7);
        for (int i = threadIdx.x; i < nc * lines; i += 256) {
            const int c = i / lines, l = i - c * lines;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P + ((int64_t)(jb + c) * ld + jb) * 8 + (int64_t)l * 128));
        }
    }
    if (part < nblk) {
        const char *dv = reinterpret_cast<const char *>(p.Dinv + (f.dinv + part) * (int64_t)(NB * XS));
        for (int i = threadIdx.x; i < NB * XS * 8 / 128; i += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(dv + (int64_t)i * 128));
    }
}

__device__ __forceinline__ void cp_async16_s(void *dst_smem, const void *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_s() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int FWD_DSM = 384;    // forward tile: acc (64) | yb (64) | part (256) | inverse block (64 x XS)

// Everything a forward tile reads that does not depend on other tasks is fetched BEFORE the tile waits: the inverse of
// its diagonal block (cp.async into shared memory), the L tile of the last block it will consume (the critical hand-off:
// registers), its right-hand-side entries and the gather map of its children's update vectors. After the wait the chain
// per 64-column block is: see the counter -> 64 values of y -> FMAs from registers -> inverse block from shared memory
// -> publish. `wait_children`: spin on fprog[s] here (after the preloads) instead of in the caller.
template <bool LDL>
__device__ void big_forward_tile(const SolveParams &p, int s, int t, double *smem, int mode, bool wait_children,
                                 unsigned long long *t_start)
{
    double *acc = smem, *yb = smem + 64, *part = smem + 128, *Dsm = smem + FWD_DSM;
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    double *u = p.uvec + f.rowp;
    const int tid = threadIdx.x;
    const int R0 = t * 64, nr = min(64, N - R0);
    const int nblk = (k + NB - 1) / NB, ntiles = (N + 63) / 64;
    const int nbefore = min(nblk, t);
    const int r = tid & 63, q = tid >> 6;
    // ---- preloads
    if (t < nblk && mode != 1) {
        const double *Dv = p.Dinv + (f.dinv + t) * (int64_t)(NB * XS);
        for (int i = tid; i < NB * XS / 2; i += 256) cp_async16_s(Dsm + 2 * i, Dv + 2 * i);
    }
    double vL[16];
    if (nbefore > 0 && mode != 1) {
        const int col0 = (nbefore - 1) * NB, nc = min(NB, k - col0);
        const double *col = P + (int64_t)col0 * ld + R0 + r;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int c = q + BIG_PART * i; vL[i] = (r < nr && c < nc) ? col[(int64_t)c * ld] : 0.0; }
    }
    constexpr int GPRE = 4;
    double v0 = 0.0;
    int g0 = 0, g1 = 0, gidx[GPRE];
    const int32_t *gs = nullptr;
    if (tid < nr) {
        const int row = R0 + tid;
        if (mode != 2) {
            v0 = (row < k) ? p.b_in[p.perm[f.c0 + row]] : 0.0;
            const int64_t go = p.gat_off[2 * (int64_t)s];
            if (go >= 0) {          // children's update vectors, in child order (split fronts always have the transposed map)
                const int32_t *gp = p.gat_ptr + go;
                gs = p.gat_src + p.gat_off[2 * (int64_t)s + 1];
                g0 = gp[row]; g1 = gp[row + 1];
#pragma unroll
                for (int i = 0; i < GPRE; ++i) gidx[i] = (g0 + i < g1) ? gs[g0 + i] : 0;
            }
        }
    }
    if (wait_children) {
        if (tid == 0) {
            while (ld_acquire_b(p.fprog + s) < f.nchild) __nanosleep(40);
            if (t_start) *t_start = globaltimer_ns();
        }
        __syncthreads();
    }
    if (tid < 64) {
        double v = v0;
        if (tid < nr) {
            const int row = R0 + tid;
            if (mode == 2) v = (row < k) ? x1[row] : u[row - k];
            else if (g1 > g0) {
                double w[GPRE];
#pragma unroll
                for (int i = 0; i < GPRE; ++i) w[i] = (g0 + i < g1) ? p.uvec[gidx[i]] : 0.0;
#pragma unroll
                for (int i = 0; i < GPRE; ++i) if (g0 + i < g1) v += w[i];
                for (int qq = g0 + GPRE; qq < g1; ++qq) v += p.uvec[gs[qq]];
            }
        }
        acc[tid] = v;
        yb[tid] = 0.0;
    }
    __syncthreads();
    if (mode == 1) {            // staged solve: the assembled right-hand side is all-reduced before the front is solved
        if (tid < nr) { const int row = R0 + tid; if (row < k) x1[row] = acc[tid]; else u[row - k] = acc[tid]; }
        return;
    }
    __shared__ int s_avail;
    for (int b = 0; b < nbefore;) {
        if (tid == 0) {             // wait for the next block, then take every block published so far in one go
            int a;
            while ((a = ld_acquire_b(p.yprog + s)) <= b) __nanosleep(20);
            s_avail = a;
        }
        __syncthreads();
        const int bend = min(nbefore, s_avail);
        for (; b < bend; ++b) {
            const int nc = min(NB, k - b * NB);
            if (tid < 64) yb[tid] = (tid < nc) ? x1[b * NB + tid] : 0.0;
            __syncthreads();
            if (b == nbefore - 1) {         // the block this tile waited for last: operands are already in registers
                double sum = 0.0;
#pragma unroll
                for (int i = 0; i < 16; ++i) sum = fma(vL[i], yb[q + BIG_PART * i], sum);
                part[q * 64 + r] = sum;
                __syncthreads();
                if (tid < nr) acc[tid] -= (part[tid] + part[64 + tid]) + (part[128 + tid] + part[192 + tid]);
                __syncthreads();
            } else {
                tile_update(P, ld, R0, 0, nr, b * NB, nc, yb, acc, part);
            }
        }
    }
    if (t < nblk) {             // diagonal tile: y_t = inv(L_tt) acc, published for the tiles below
        const int jb = t * NB, nbk = min(NB, k - jb);
        cp_async_wait_all_s();
        __syncthreads();
        {
            const int rr = tid >> 2, qd = tid & 3;
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < NB / 4; ++c) a = fma(Dsm[(qd + 4 * c) * XS + rr], (qd + 4 * c < nbk) ? acc[qd + 4 * c] : 0.0, a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            __syncthreads();
            if (qd == 0) { yb[rr] = (rr < nbk) ? a : 0.0; if (rr < nbk) x1[jb + rr] = a; }
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(p.yprog + s, 1); }
        if (nr > nbk) tile_update(P, ld, R0, nbk, nr, jb, nbk, yb, acc, part);     // rows of this tile below the last diagonal block
    }
    if (tid < nr && R0 + tid >= k) u[R0 + tid - k] = acc[tid];
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(p.ftiles + s, 1) + 1 == ntiles) {
            __threadfence();
            atomicAdd(p.fprog + (f.parent >= 0 ? f.parent : s), 1);
        }
    }
}

// Like the forward tile, a backward block fetches what does not depend on other tasks before it waits (for its parent
// front, then for the blocks above it): its inverse block, the row indices of the ancestors' x it reads, and the L tile
// under the block right above it (the last hand-off of its chain), kept in registers.
template <bool LDL>
__device__ void big_backward_block(const SolveParams &p, int s, int b, double *smem, unsigned long long *t_start)
{
    double *S = smem, *wb = smem + NB * LDS, *xr = wb + NB;
    const FrontInfo f = p.fi[s];
    const int k = f.k, r = f.r, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    const int32_t *rows = p.row_idx + f.rowp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cached = r <= XR_MAX;
    const int nblk = (k + NB - 1) / NB;
    const int jb = b * NB, nb = min(NB, k - jb);
    const double *Dv = p.Dinv + (f.dinv + b) * (int64_t)(NB * XS);
    const int q0 = warp * 8;
    const double *col = P + (int64_t)(jb + q0) * ld;
    // ---- preloads
    constexpr int RPRE = XR_MAX / 256;
    int ridx[RPRE];
#pragma unroll
    for (int i = 0; i < RPRE; ++i) ridx[i] = (cached && tid + 256 * i < r) ? rows[tid + 256 * i] : 0;
    for (int idx = tid; idx < NB * NB; idx += 256) S[(idx >> 6) * LDS + (idx & 63)] = Dv[(idx >> 6) * XS + (idx & 63)];
    if (tid < NB) wb[tid] = 0.0;
    double vB[16];                  // L[64 (b + 1) + lane (+ 32), jb + q0 + c]
    const int pre0 = NB * (b + 1), pre1 = min(pre0 + NB, k);        // rows of the block right above (empty for the last block)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const bool on = q0 + c < nb;
        vB[c] = (on && pre0 + lane < pre1) ? col[(int64_t)c * ld + pre0 + lane] : 0.0;
        vB[8 + c] = (on && pre0 + lane + 32 < pre1) ? col[(int64_t)c * ld + pre0 + lane + 32] : 0.0;
    }
    if (tid == 0) {
        if (f.parent >= 0) { while (ld_acquire_b(p.bdone + f.parent) == 0) __nanosleep(40); }
        else { while (ld_acquire_b(p.fprog + s) <= f.nchild) __nanosleep(40); }
        if (t_start) *t_start = globaltimer_ns();
    }
    __syncthreads();
    if (cached) {
#pragma unroll
        for (int i = 0; i < RPRE; ++i) if (tid + 256 * i < r) xr[tid + 256 * i] = p.xp[ridx[i]];
    }
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0;
    auto xfull = [&](int i) { return (i < k) ? x1[i] : (cached ? xr[i - k] : p.xp[rows[i - k]]); };
    auto rows_range = [&](int i0, int i1) {         // acc[c] += sum_{i0 <= i < i1} L[i, jb + q0 + c] xfull[i]
        if (q0 >= nb) return;
        int i = i0 + lane;
        for (; i + 32 < i1; i += 64) {              // two row chunks per trip: 16 loads in flight per lane
            const double xv = xfull(i), xw = xfull(i + 32);
            double v[8], w[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const bool on = q0 + c < nb;
                v[c] = on ? col[(int64_t)c * ld + i] : 0.0;
                w[c] = on ? col[(int64_t)c * ld + i + 32] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fma(w[c], xw, fma(v[c], xv, acc[c]));
        }
        for (; i < i1; i += 32) {
            const double xv = xfull(i);
            double v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = (q0 + c < nb) ? col[(int64_t)c * ld + i] : 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fma(v[c], xv, acc[c]);
        }
    };
    // the rows below the supernode first (ancestors' x: known), then the blocks above this one as they are published
    // (from the far end; every block published so far is taken in one go), the block right above it last
    __shared__ int s_avail;
    __syncthreads();
    rows_range(k, N);
    for (int done_t = nblk; done_t > b + 1;) {
        if (tid == 0) {
            int a;
            while ((a = ld_acquire_b(p.xprog + s)) < nblk - (done_t - 1)) __nanosleep(20);
            s_avail = a;
        }
        __syncthreads();
        const int new_t = max(b + 1, nblk - s_avail);
        rows_range(NB * max(new_t, b + 2), min(NB * done_t, k));
        if (new_t == b + 1 && q0 < nb) {            // the block right above: operands are already in registers
            const double xv = (pre0 + lane < pre1) ? x1[pre0 + lane] : 0.0;
            const double xw = (pre0 + lane + 32 < pre1) ? x1[pre0 + lane + 32] : 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fma(vB[8 + c], xw, fma(vB[c], xv, acc[c]));
        }
        done_t = new_t;
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (q0 < nb && lane < 8 && q0 + lane < nb) {
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) if (lane == c) a = acc[c];
        double y = x1[jb + q0 + lane];
        if (LDL) y = y / col[(int64_t)lane * ld + jb + q0 + lane];
        wb[q0 + lane] = y - a;
    }
    __syncthreads();
    {
        const int cc = tid >> 2, q = tid & 3;
        double a = 0.0;
        for (int rr = cc + q; rr < NB; rr += 4) a = fma(S[cc * LDS + rr], wb[rr], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (q == 0 && cc < nb) {
            x1[jb + cc] = a;
            double *o = p.x_out + p.perm[f.c0 + jb + cc];
            *o = p.accumulate ? *o + a : a;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(p.xprog + s, 1);
        if (b == 0) atomicExch(p.bdone + s, 1);
    }
}

// Small leaf fronts in the solves: one warp per front, L11 (k <= SL_K) applied by direct substitution.
template <bool LDL>
__device__ void leaf_forward(const SolveParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const int lane = threadIdx.x & 31;
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    double y[SL_K];
#pragma unroll
    for (int j = 0; j < SL_K; ++j) {
        y[j] = 0.0;
        if (j < k) {
            double t = p.b_in[p.perm[f.c0 + j]];
#pragma unroll
            for (int q = 0; q < j; ++q) t = fma(-P[(int64_t)q * ld + j], y[q], t);
            if (!LDL) t = t / P[(int64_t)j * ld + j];
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
        for (int j = 0; j < SL_K; ++j) if (j < k) acc = fma(P[(int64_t)j * ld + lane], y[j], acc);
        p.uvec[f.rowp + lane - k] = -acc;          // a leaf has no children: its update vector starts from zero
    }
}

template <bool LDL>
__device__ void leaf_backward(const SolveParams &p, int s)
{
    const FrontInfo f = p.fi[s];
    const int k = f.k, N = f.k + f.r, ld = front_ld(f.k, f.r);
    const int lane = threadIdx.x & 31;
    const double *P = p.L + f.lp;
    double *x1 = p.xp + f.c0;
    const double xv = (lane >= k && lane < N) ? p.xp[p.row_idx[f.rowp + lane - k]] : 0.0;
    double w[SL_K];
#pragma unroll
    for (int j = 0; j < SL_K; ++j) {
        double part = (j < k && lane >= k && lane < N) ? P[(int64_t)j * ld + lane] * xv : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        double yj = (j < k) ? x1[j] : 0.0;
        if (LDL && j < k) yj = yj / P[(int64_t)j * ld + j];
        w[j] = yj - part;
    }
#pragma unroll
    for (int j = SL_K - 1; j >= 0; --j) {
        if (j < k) {
            double t = w[j];
#pragma unroll
            for (int q = j + 1; q < SL_K; ++q) if (q < k) t = fma(-P[(int64_t)j * ld + q], w[q], t);
            if (!LDL) t = t / P[(int64_t)j * ld + j];
            w[j] = t;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < SL_K; ++j)
            if (j < k) {
                x1[j] = w[j];
                double *o = p.x_out + p.perm[f.c0 + j];
                *o = p.accumulate ? *o + w[j] : w[j];
            }
    }
}

__device__ __forceinline__ int ld_acquire_s(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <bool LDL>
__global__ void __launch_bounds__(256, 2) k_solve_tasks(const __grid_constant__ SolveParams p)
{
    __shared__ __align__(16) double smem[NB * XS + NB + XR_MAX];     // >= FWD_DSM + NB * XS
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = p.task_begin + atomicAdd(p.ticket, 1);
    __syncthreads();
    for (;;) {
        const int t = s_ticket;
        if (t >= p.task_end) break;
        const int4 tk = p.tasks[t];
        int next = 0;
        unsigned long long tr0 = 0, tr1 = 0;
        if (p.trace && tid == 0) tr0 = tr1 = globaltimer_ns();
        if (tk.x == 4 || tk.x == 5) {
            const int s = tk.y;
            const FrontInfo f = p.fi[s];
            const bool fwd = (tk.x == 4);
            const int mode = (fwd && s == p.root_sn) ? p.root_mode : 0;
            if (fwd ? (f.nchild > 0 || tk.z > 0) : true) prefetch_big(p, f, fwd, tk.z);      // the task is going to wait
            if (tid == 0) {
                next = p.task_begin + atomicAdd(p.ticket, 1);
                if (p.trace) tr1 = globaltimer_ns();
            }
            __syncthreads();
            if (fwd) big_forward_tile<LDL>(p, s, tk.z, smem, mode, f.nchild > 0 && mode != 2, p.trace ? &tr1 : nullptr);
            else big_backward_block<LDL>(p, s, tk.z, smem, p.trace ? &tr1 : nullptr);
        } else if (tk.x == 0 || tk.x == 2) {
            const int s = tk.y;
            const FrontInfo f = p.fi[s];
            const bool fwd = (tk.x == 0);
            const int mode = (fwd && s == p.root_sn) ? p.root_mode : 0;
            // a root's backward step follows its own forward step (another task): that one bumps fprog[s] once more
            const bool may_wait = fwd ? (f.nchild > 0 && mode != 2) : true;
            if (may_wait) prefetch_front(p, s);
            if (tid == 0) {
                if (may_wait) {
                    if (fwd) { while (ld_acquire_s(p.fprog + s) < f.nchild) __nanosleep(40); }
                    else if (f.parent >= 0) { while (ld_acquire_s(p.bdone + f.parent) == 0) __nanosleep(40); }
                    else { while (ld_acquire_s(p.fprog + s) <= f.nchild) __nanosleep(40); }
                }
                if (p.trace) tr1 = globaltimer_ns();
                next = p.task_begin + atomicAdd(p.ticket, 1);
            }
            __syncthreads();
            if (fwd) front_forward<LDL>(p, s, smem, mode);
            else front_backward<LDL>(p, s, smem);
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                if (fwd) { if (mode != 1) atomicAdd(p.fprog + (f.parent >= 0 ? f.parent : s), 1); }
                else atomicExch(p.bdone + s, 1);
            }
        } else {
            if (tid == 0) next = p.task_begin + atomicAdd(p.ticket, 1);
            const int li = tk.y + (tid >> 5);
            if (li < p.n_leaf) {
                const int s = p.leaves[li];
                if (tk.x == 1) {
                    leaf_forward<LDL>(p, s);
                    __syncwarp();
                    if ((tid & 31) == 0) {
                        const int par = p.fi[s].parent;
                        __threadfence();
                        atomicAdd(p.fprog + (par >= 0 ? par : s), 1);
                    }
                } else {
                    const int par = p.fi[s].parent;
                    if ((tid & 31) == 0) {
                        if (par >= 0) { while (ld_acquire_s(p.bdone + par) == 0) __nanosleep(40); }
                        else { while (ld_acquire_s(p.fprog + s) == 0) __nanosleep(40); }
                    }
                    __syncwarp();
                    leaf_backward<LDL>(p, s);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_ticket = next;
            if (p.trace) { p.trace[3 * (size_t)t] = tr0; p.trace[3 * (size_t)t + 1] = tr1; p.trace[3 * (size_t)t + 2] = globaltimer_ns(); }
        }
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
    for (int64_t q = ptr[row] + lane; q < ptr[row + 1]; q += 32) acc = fma(__ldg(val + vpos[q]), __ldg(x + col[q]), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rout[row] = b[row] - acc;
}

inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)std::max<int64_t>(1, (n + per_block - 1) / per_block); }

}  // namespace

// Solve-side part of ls_device_setup (factor.cu): task list, grid size, transposed child maps.
int ls_solve_setup(Handle *h, const void *finfo_host, const char *small)
{
    const LsSymbolic &S = h->sym;
    const int ns = S.ns;
    const FrontInfo *finfo = (const FrontInfo *)finfo_host;
    cudaStream_t st = h->stream;
    DeviceInfo prop;
    if (device_info(h->device, prop) != MIPM_OK) return fail(h, MIPM_ERR_CUDA, "cudaGetDeviceProperties failed");
    int occ_s = 0;
    if (S.kind == MIPM_LDL) MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_solve_tasks<true>, 256, 0));
    else MIPM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_solve_tasks<false>, 256, 0));
    if (occ_s < 1) return fail(h, MIPM_ERR_CUDA, "solve kernel does not fit on an SM");
    h->grid_solve = prop.sm_count * std::min(occ_s, 4);
    if (h->grid_limit > 0) h->grid_solve = std::min(h->grid_solve, h->grid_limit);
    // ---- task list: forward by level (small leaves first, eight per task), then backward from the root down
    std::vector<int4> tasks;
    const int n_leaf = h->n_leaf;
    int big_k = 128;            // fronts at least this wide (measured: 128 beats 64, 256 and 512 on C2 and on its mesh variant) are solved by many CTAs (row tiles forward, column blocks backward)
    if (const char *e = std::getenv("MIPM_SOLVE_BIG_K")) big_k = std::max(NB, atoi(e));
    // ... unless the level holds at least as many such fronts as the grid has CTAs (a stacked batch of dense blocks): the
    // fronts are independent, one CTA per front already fills the machine and the tile hand-offs only add latency
    // Near the root a level has fewer fronts than the machine has SMs and the solve is a chain of hand-offs: there every
    // front of two or more row tiles is split (a hand-off between tiles costs less than a block step inside one CTA).
    int sparse_level = 96;      // a level with at most this many fronts is latency bound
    if (const char *e = std::getenv("MIPM_SOLVE_SPARSE_LEVEL")) sparse_level = atoi(e);
    std::vector<char> level_split((size_t)std::max(S.n_levels, 1), 1);
    std::vector<int> level_of((size_t)std::max(ns, 1), 0);
    for (int l = 0; l < S.n_levels; ++l) {
        int64_t nbig = 0, nfront = 0;
        for (int64_t t = S.level_ptr[(size_t)l]; t < S.level_ptr[(size_t)l + 1]; ++t) {
            const int s2 = S.level_sn[(size_t)t];
            level_of[(size_t)s2] = l;
            nbig += finfo[(size_t)s2].k >= big_k;
            nfront += !small[(size_t)s2];
        }
        if (nbig >= h->grid_solve) level_split[(size_t)l] = 0;
        else if (nfront <= sparse_level) level_split[(size_t)l] = 2;
    }
    auto is_big = [&](int s2) {
        const char m = level_split[(size_t)level_of[(size_t)s2]];
        const FrontInfo &f2 = finfo[(size_t)s2];
        return (m >= 1 && f2.k >= big_k) || (m == 2 && f2.k + f2.r > 2 * NB);
    };
    for (int i = 0; i < n_leaf; i += 8) tasks.push_back(make_int4(1, i, 0, 0));
    h->solve_root_fwd_begin = 0;
    for (int l = 0; l < S.n_levels; ++l)
        for (int64_t t = S.level_ptr[(size_t)l]; t < S.level_ptr[(size_t)l + 1]; ++t) {
            const int s2 = S.level_sn[(size_t)t];
            if (small[(size_t)s2]) continue;
            if (s2 == S.root_sn) h->solve_root_fwd_begin = (int)tasks.size();
            if (is_big(s2)) {
                const int ntiles = (finfo[(size_t)s2].k + finfo[(size_t)s2].r + 63) / 64;
                for (int q = 0; q < ntiles; ++q) tasks.push_back(make_int4(4, s2, q, 0));
            } else {
                tasks.push_back(make_int4(0, s2, 0, 0));
            }
        }
    h->n_solve_fwd = (int)tasks.size();
    for (int l = S.n_levels - 1; l >= 0; --l)
        for (int64_t t = S.level_ptr[(size_t)l]; t < S.level_ptr[(size_t)l + 1]; ++t) {
            const int s2 = S.level_sn[(size_t)t];
            if (small[(size_t)s2]) continue;
            if (is_big(s2)) {
                const int nblk = (finfo[(size_t)s2].k + NB - 1) / NB;
                for (int q = nblk - 1; q >= 0; --q) tasks.push_back(make_int4(5, s2, q, 0));
            } else {
                tasks.push_back(make_int4(2, s2, 0, 0));
            }
        }
    for (int i = 0; i < n_leaf; i += 8) tasks.push_back(make_int4(3, i, 0, 0));
    h->n_solve_tasks = (int)tasks.size();
    h->grid_solve = std::max(1, std::min(h->grid_solve, h->n_solve_tasks));
    MIPM_CUDA(h, h->d_solve_tasks.alloc(std::max<size_t>(tasks.size(), 1)));
    if (!tasks.empty()) MIPM_CUDA(h, cudaMemcpyAsync(h->d_solve_tasks.p, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    MIPM_CUDA(h, h->d_solve_prog.alloc((size_t)5 * std::max(ns, 1) + 4));
    {
        // transposed child maps for the forward solve (fronts with more than GATHER_MIN_CHILDREN children)
        constexpr int GATHER_MIN_CHILDREN = 4;
        std::vector<int64_t> gat_off((size_t)2 * std::max(ns, 1), -1);
        std::vector<int32_t> gat_ptr, gat_src, cnt;
        const bool slots_fit = S.row_ptr[(size_t)ns] < (int64_t)INT32_MAX;
        for (int s = 0; s < ns && slots_fit; ++s) {
            const FrontInfo &f = finfo[(size_t)s];
            if (f.nchild <= GATHER_MIN_CHILDREN && !(is_big(s) && f.nchild > 0)) continue;
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
    return MIPM_OK;
}

// stage -1: whole solve; 0: forward sweep below the root, the root only folds its children's contributions in;
// 1: the root's own forward step, backward sweep
static int solve_once(Handle *h, const double *b_in, double *x_out, int accumulate, int stage = -1)
{
    const LsSymbolic &S = h->sym;
    const int ns = S.ns;
    SolveParams p;
    p.fi = (const FrontInfo *)h->d_finfo.p; p.child_idx = h->d_child_idx.p; p.rel_idx = h->d_rel_idx.p; p.row_idx = h->d_row_idx.p;
    p.perm = h->d_perm.p; p.leaves = h->d_sched.p + h->leaf_off; p.n_leaf = h->n_leaf;
    p.tasks = (const int4 *)h->d_solve_tasks.p;
    p.task_begin = (stage == 1) ? h->solve_root_fwd_begin : 0;
    p.task_end = (stage == 0) ? h->n_solve_fwd : h->n_solve_tasks;
    p.root_sn = (stage >= 0) ? S.root_sn : -1;
    p.root_mode = (stage == 0) ? 1 : ((stage == 1) ? 2 : 0);
    p.fprog = h->d_solve_prog.p; p.bdone = h->d_solve_prog.p + ns; p.yprog = h->d_solve_prog.p + 2 * (size_t)ns;
    p.xprog = h->d_solve_prog.p + 3 * (size_t)ns; p.ftiles = h->d_solve_prog.p + 4 * (size_t)ns; p.ticket = h->d_solve_prog.p + 5 * (size_t)ns;
    p.L = h->L_cur; p.Dinv = h->d_Dinv.p; p.xp = h->d_xp.p; p.uvec = h->d_uvec.p;
    p.b_in = b_in; p.x_out = x_out; p.accumulate = accumulate;
    p.gat_off = h->d_gat_off.p; p.gat_ptr = h->d_gat_ptr.p; p.gat_src = h->d_gat_src.p;
    if (stage == 1) { MIPM_CUDA(h, cudaMemsetAsync(h->d_solve_prog.p + 5 * (size_t)ns, 0, sizeof(int), h->stream)); }   // ticket only
    else { MIPM_CUDA(h, cudaMemsetAsync(h->d_solve_prog.p, 0, ((size_t)5 * ns + 4) * sizeof(int), h->stream)); }
    if (p.task_end <= p.task_begin) return MIPM_OK;
    const char *trace_path = std::getenv("MIPM_SOLVE_TRACE");      // diagnostic only: synchronises
    DBuf<unsigned long long> d_trace;
    p.trace = nullptr;
    if (trace_path) {
        MIPM_CUDA(h, d_trace.alloc((size_t)3 * h->n_solve_tasks));
        MIPM_CUDA(h, cudaMemsetAsync(d_trace.p, 0, (size_t)3 * h->n_solve_tasks * sizeof(unsigned long long), h->stream));
        p.trace = d_trace.p;
    }
    void *args[] = {&p};
    const void *fn = (S.kind == MIPM_LDL) ? (const void *)k_solve_tasks<true> : (const void *)k_solve_tasks<false>;
    // cooperative launch: every CTA is resident, which the dependency spins rely on
    MIPM_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)h->grid_solve), dim3(256), args, 0, h->stream));
    h->launches++;
    if (trace_path) {
        std::vector<unsigned long long> tr((size_t)3 * h->n_solve_tasks);
        std::vector<int4> tk((size_t)h->n_solve_tasks);
        MIPM_CUDA(h, cudaMemcpyAsync(tr.data(), d_trace.p, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        MIPM_CUDA(h, cudaMemcpyAsync(tk.data(), h->d_solve_tasks.p, tk.size() * sizeof(int4), cudaMemcpyDeviceToHost, h->stream));
        MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
        unsigned long long t0 = ~0ull;
        for (size_t i = 0; i < tk.size(); ++i) if (tr[3 * i]) t0 = std::min(t0, tr[3 * i]);
        if (FILE *tf = std::fopen(trace_path, "w")) {
            std::fprintf(tf, "task,kind,id,level,k,r,wait_us,start_us,end_us\n");
            for (size_t i = 0; i < tk.size(); ++i) {
                if (!tr[3 * i]) continue;
                const bool front = (tk[i].x == 0 || tk[i].x == 2 || tk[i].x == 4 || tk[i].x == 5);
                const int s2 = tk[i].y;
                std::fprintf(tf, "%zu,%d,%d,%d,%d,%d,%.2f,%.2f,%.2f\n", i, tk[i].x, s2, front ? S.sn_level[(size_t)s2] : 0,
                             front ? S.sn_ptr[(size_t)s2 + 1] - S.sn_ptr[(size_t)s2] : 0,
                             front ? (int)(S.row_ptr[(size_t)s2 + 1] - S.row_ptr[(size_t)s2]) : 0,
                             (double)(tr[3 * i] - t0) * 1e-3, (double)(tr[3 * i + 1] - t0) * 1e-3, (double)(tr[3 * i + 2] - t0) * 1e-3);
            }
            std::fclose(tf);
        }
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
    if (ir_steps > 0 && !h->d_full_ptr.p) {      // first refinement: build the symmetric operator of the residual
        ls_build_full_csr(h->sym);
        MIPM_CUDA(h, h->d_full_ptr.upload(h->sym.full_ptr, st));
        MIPM_CUDA(h, h->d_full_col.upload(h->sym.full_col, st));
        MIPM_CUDA(h, h->d_full_val.upload(h->sym.full_val, st));
        MIPM_CUDA(h, cudaStreamSynchronize(st));
        std::vector<int64_t>().swap(h->sym.full_ptr);
        std::vector<int32_t>().swap(h->sym.full_col);
        std::vector<int64_t>().swap(h->sym.full_val);
    }
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
    if (h->root_task_begin < 0) return fail(h, MIPM_ERR_STATE, "staged solve needs mipm_ls_analyze_border");
    if (S.n == 0) return MIPM_OK;
    if (stage == 0) {
        MIPM_CUDA(h, cudaMemcpyAsync(h->d_b.p, d_x, (size_t)S.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return solve_once(h, h->d_b.p, d_x, 0, 0);
    }
    return solve_once(h, h->d_b.p, d_x, 0, 1);
}

}  // namespace mipm

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
