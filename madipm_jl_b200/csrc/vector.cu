// Fused Mehrotra predictor-corrector vector kernels (replaces the ~40 broadcast and ~20
// mapreduce launches per iteration that CUDA.jl generates for src/kernels.jl and
// src/solver.jl). All reductions are single-launch, deterministic
// ("last block folds the per-block partials in a fixed order") and return through one
// pinned-host scalar block, so each entry point that returns a scalar costs one sync.
//
// Bound blocks: the reference indexes x, xl, zl, ... through ind_lb / ind_ub views
// (src/structure.jl:146-153). At bind time we build the inverse maps (variable -> position in
// the lb / ub block, or -1) so every kernel is one coalesced pass over the n variables.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "common.h"

namespace mipm {

namespace {

constexpr int TB = 256;

enum { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2 };

__device__ __forceinline__ double comb(double a, double b, int op)
{
    if (op == OP_SUM) return a + b;
    if (op == OP_MAX) return (a > b || a != a) ? a : b;     // NaN propagates like Julia's norm(.,Inf)
    return (a < b || a != a) ? a : b;
}

// Deterministic grid-wide reduction of NV values per thread; result in out[0..NV).
template <int NV>
__device__ bool grid_reduce_vals(double (&acc)[NV], const int (&op)[NV], double *partials, unsigned int *counter,
                                 double *out)
{
    __shared__ double sm[NV][TB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] = comb(acc[q], __shfl_xor_sync(0xffffffffu, acc[q], o), op[q]);
        if (lane == 0) sm[q][warp] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        const int q = threadIdx.x;
        double v = sm[q][0];
        for (int w = 1; w < TB / 32; ++w) v = comb(v, sm[q][w], op[q]);
        partials[(size_t)q * gridDim.x + blockIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    const double ident[3] = {0.0, -DBL_MAX, DBL_MAX};
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double v = ident[op[q]];
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += TB) v = comb(v, partials[(size_t)q * gridDim.x + b], op[q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = comb(v, __shfl_xor_sync(0xffffffffu, v, o), op[q]);
        if (lane == 0) sm[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        const int q = threadIdx.x;
        double v = sm[q][0];
        for (int w = 1; w < TB / 32; ++w) v = comb(v, sm[q][w], op[q]);
        out[q] = v;
    }
    if (threadIdx.x == 0) *counter = 0;
    __syncthreads();
    return true;
}

// (value, index) arg-min pairs with the reference's tie rule: `a < b ? a : b` keeps the
// right-most of equal values (src/kernels.jl:232), i.e. the larger index; init = (1.0, 0).
struct VI { double v; long long i; };
__device__ __forceinline__ VI comb_vi(VI a, VI b)
{
    if (a.v < b.v) return a;
    if (b.v < a.v) return b;
    return (a.i > b.i) ? a : b;
}
template <int NP>
__device__ bool grid_reduce_argmin(VI (&acc)[NP], double *partials, unsigned int *counter, double *out)
{
    __shared__ VI sm[NP][TB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long *ipart = (long long *)(partials + (size_t)NP * gridDim.x);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            VI other;
            other.v = __shfl_xor_sync(0xffffffffu, acc[q].v, o);
            other.i = __shfl_xor_sync(0xffffffffu, acc[q].i, o);
            acc[q] = comb_vi(acc[q], other);
        }
        if (lane == 0) sm[q][warp] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < NP) {
        const int q = threadIdx.x;
        VI v = sm[q][0];
        for (int w = 1; w < TB / 32; ++w) v = comb_vi(v, sm[q][w]);
        partials[(size_t)q * gridDim.x + blockIdx.x] = v.v;
        ipart[(size_t)q * gridDim.x + blockIdx.x] = v.i;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        VI v;
        v.v = 1.0;
        v.i = 0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += TB) {
            VI o;
            o.v = partials[(size_t)q * gridDim.x + b];
            o.i = ipart[(size_t)q * gridDim.x + b];
            v = comb_vi(v, o);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            VI other;
            other.v = __shfl_xor_sync(0xffffffffu, v.v, o);
            other.i = __shfl_xor_sync(0xffffffffu, v.i, o);
            v = comb_vi(v, other);
        }
        if (lane == 0) sm[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NP) {
        const int q = threadIdx.x;
        VI v = sm[q][0];
        for (int w = 1; w < TB / 32; ++w) v = comb_vi(v, sm[q][w]);
        out[q] = v.v;
        out[NP + q] = (double)v.i;
    }
    if (threadIdx.x == 0) *counter = 0;
    __syncthreads();
    return true;
}

// Device-resident scalar block of the fused iteration (doubles).
enum { SC_ALPHA = 0,       // alpha_xl, alpha_xu, alpha_zl, alpha_zu + 4 arg-min indices
       SC_ALPHA_P = 8, SC_ALPHA_D = 9, SC_MU_AFF = 10, SC_MU_CURR = 11, SC_MU = 12, SC_TAU = 13,
       SC_TERM = 16,       // 7 raw termination reductions
       SC_RES = 24,        // ||w||, ||p|| of the predictor solve, then of the corrector solve
       SC_OBJ = 28,        // c'x, x'Hx
       SC_SUMS = 32,       // scratch: affine_l, affine_u, cur_l, cur_u
       SC_COUNT = 40 };

struct V {   // device view of mipm_mpc_vectors + inverse maps
    int64_t n, m, nlb, nub;
    const int32_t *inv_lb, *inv_ub;
    double *x, *xl, *xu, *zl, *zu, *f, *y, *c, *rhs, *jacl, *d, *p, *w, *corr_lb, *corr_ub;
    double *reg, *pr_diag, *du_diag, *l_diag, *u_diag, *l_lower, *u_lower;
};

#define GRID_STRIDE(i, len) for (int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x; i < (len); i += (int64_t)gridDim.x * TB)

__global__ void k_build_inv(int64_t nidx, const int64_t *__restrict__ ind, int base, int32_t *__restrict__ inv)
{
    GRID_STRIDE(j, nidx) inv[ind[j] - base] = (int32_t)j;
}

// src/kernels.jl:124-136
__global__ void __launch_bounds__(TB) k_set_aug_diag(V v, double del_w, double del_c)
{
    GRID_STRIDE(i, v.n) {
        double pr = del_w;
        v.reg[i] = del_w;
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            double ld = v.xl[i] - v.x[i], zl = v.zl[i];
            v.l_diag[jl] = ld;
            v.l_lower[jl] = zl;
            pr = pr - zl / ld;
        }
        if (ju >= 0) {
            double ud = v.x[i] - v.xu[i], zu = v.zu[i];
            v.u_diag[ju] = ud;
            v.u_lower[ju] = zu;
            pr = pr - zu / ud;
        }
        v.pr_diag[i] = pr;
    }
    GRID_STRIDE(i, v.m) v.du_diag[i] = del_c;
}

// src/kernels.jl:21-58 (corr == 0: predictive; corr == 1: correction with mu)
__global__ void __launch_bounds__(TB) k_set_rhs(V v, int corr, double mu, const double *mu_dev)
{
    if (mu_dev) mu = *mu_dev;
    double *px = v.p, *py = v.p + v.n, *pzl = v.p + v.n + v.m, *pzu = v.p + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double x = v.x[i], zl = v.zl[i], zu = v.zu[i];
        px[i] = ((-v.f[i] + zl) - zu) - v.jacl[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            double t = (v.xl[i] - x) * zl;
            pzl[jl] = corr ? (t + mu) - v.corr_lb[jl] : t;
        }
        if (ju >= 0) {
            double t = (v.xu[i] - x) * zu;
            pzu[ju] = corr ? (t - mu) - v.corr_ub[ju] : t;
        }
    }
    GRID_STRIDE(i, v.m) py[i] = -v.c[i];
}

// src/kernels.jl:60-71
__global__ void __launch_bounds__(TB) k_get_correction(V v)
{
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) v.corr_lb[jl] = dx[i] * dzl[jl];
        if (ju >= 0) v.corr_ub[ju] = dx[i] * dzu[ju];
    }
}

// src/kernels.jl:74-122
__global__ void __launch_bounds__(TB) k_set_extra_correction(V v, double ap, double ad, double tmin, double tmax)
{
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            double x_ = (v.x[i] + ap * dx[i]) - v.xl[i];
            double z_ = v.zl[i] + ad * dzl[jl];
            double val = x_ * z_;
            double dl = (val < tmin) ? tmin - val : ((val > tmax) ? tmax - val : 0.0);
            v.corr_lb[jl] = v.corr_lb[jl] - dl;
        }
        if (ju >= 0) {
            double x_ = (v.xu[i] - ap * dx[i]) - v.x[i];
            double z_ = v.zu[i] + ad * dzu[ju];
            double val = x_ * z_;
            double dl = (val < tmin) ? tmin - val : ((val > tmax) ? tmax - val : 0.0);
            v.corr_ub[ju] = v.corr_ub[ju] + dl;
        }
    }
}

// src/kernels.jl:155-208: out[0] = sum_l, out[1] = sum_u (affine == 0: current point)
__global__ void __launch_bounds__(TB) k_compl_measure(V v, int affine, double ap, double ad, double *partials,
                                                      unsigned int *counter, double *out, const double *sc_in = nullptr)
{
    if (sc_in) { ap = sc_in[SC_ALPHA_P]; ad = sc_in[SC_ALPHA_D]; }     // step lengths left on the device by k_alpha_max
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    double acc[2] = {0.0, 0.0};
    GRID_STRIDE(i, v.n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            if (affine) acc[0] += ((v.x[i] + ap * dx[i]) - v.xl[i]) * (v.zl[i] + ad * dzl[jl]);
            else acc[0] += (v.x[i] - v.xl[i]) * v.zl[i];
        }
        if (ju >= 0) {
            if (affine) acc[1] += (v.xu[i] - (v.x[i] + ap * dx[i])) * (v.zu[i] + ad * dzu[ju]);
            else acc[1] += (v.xu[i] - v.x[i]) * v.zu[i];
        }
    }
    const int op[2] = {OP_SUM, OP_SUM};
    grid_reduce_vals<2>(acc, op, partials, counter, out);
}

// prediction_step! after the affine solve (src/solver.jl:232-235) in one pass: affine and current
// complementarity sums (kernels.jl:155-208), get_correction! (kernels.jl:60-71) and the Mehrotra
// barrier update (kernels.jl:210-220) by the last block. alpha_aff and the results stay on the device.
__global__ void __launch_bounds__(TB) k_predictor_measures(V v, double mu_min, double *partials, unsigned int *counter, double *sc)
{
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    const double ap = sc[SC_ALPHA_P], ad = sc[SC_ALPHA_D];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    GRID_STRIDE(i, v.n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            acc[0] += ((v.x[i] + ap * dx[i]) - v.xl[i]) * (v.zl[i] + ad * dzl[jl]);
            acc[2] += (v.x[i] - v.xl[i]) * v.zl[i];
            v.corr_lb[jl] = dx[i] * dzl[jl];
        }
        if (ju >= 0) {
            acc[1] += (v.xu[i] - (v.x[i] + ap * dx[i])) * (v.zu[i] + ad * dzu[ju]);
            acc[3] += (v.xu[i] - v.x[i]) * v.zu[i];
            v.corr_ub[ju] = dx[i] * dzu[ju];
        }
    }
    const int op[4] = {OP_SUM, OP_SUM, OP_SUM, OP_SUM};
    const bool last = grid_reduce_vals<4>(acc, op, partials, counter, sc + SC_SUMS);
    if (last && threadIdx.x == 0) {
        const double cnt = (double)(v.nlb + v.nub);
        double mu_aff = 0.0, mu_cur = 0.0, sigma = 1.0;
        if (cnt > 0) {
            mu_aff = (sc[SC_SUMS + 0] + sc[SC_SUMS + 1]) / cnt;
            mu_cur = (sc[SC_SUMS + 2] + sc[SC_SUMS + 3]) / cnt;
            const double rr = mu_aff / mu_cur;
            sigma = fmin(fmax(rr * rr * rr, 1e-6), 10.0);
        }
        sc[SC_MU_AFF] = mu_aff;
        sc[SC_MU_CURR] = mu_cur;
        sc[SC_MU] = fmax(mu_min, sigma * mu_cur);
    }
}

// src/kernels.jl:226-272: four ratio tests in one pass.
__global__ void __launch_bounds__(TB) k_alpha_max(V v, double tau, double *partials, unsigned int *counter, double *out,
                                                  const double *sc_in, double tau_min, double *sc_out)
{
    // fused iteration: AdaptiveStep tau = max(1 - mu, tau_min) from the device-resident mu
    if (sc_in) tau = fmax(1.0 - sc_in[SC_MU], tau_min);
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    VI acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[q].v = 1.0; acc[q].i = 0; }
    GRID_STRIDE(i, v.n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        double dxi = dx[i];
        if (jl >= 0) {
            VI c;
            c.i = jl + 1;
            c.v = (dxi < 0.0) ? (-v.x[i] + v.xl[i]) * tau / dxi : INFINITY;
            acc[0] = comb_vi(acc[0], c);
            double dz = dzl[jl];
            c.v = (dz < 0.0) ? (-v.zl[i]) * tau / dz : INFINITY;
            acc[2] = comb_vi(acc[2], c);
        }
        if (ju >= 0) {
            VI c;
            c.i = ju + 1;
            c.v = (dxi > 0.0) ? (-v.x[i] + v.xu[i]) * tau / dxi : INFINITY;
            acc[1] = comb_vi(acc[1], c);
            double dz = dzu[ju], zu = v.zu[i];
            c.v = ((dz < 0.0) && (zu + dz < 0.0)) ? (-zu) * tau / dz : INFINITY;
            acc[3] = comb_vi(acc[3], c);
        }
    }
    const bool last = grid_reduce_argmin<4>(acc, partials, counter, out);
    if (last && sc_out && threadIdx.x == 0) {
        sc_out[SC_ALPHA_P] = fmin(out[0], out[1]);     // get_fraction_to_boundary_step (kernels.jl:288)
        sc_out[SC_ALPHA_D] = fmin(out[2], out[3]);
        sc_out[SC_TAU] = tau;
    }
}

// update_step!(::MehrotraAdaptiveStep), src/kernels.jl:309-358, after k_alpha_max(tau = 1) left the four ratio tests and
// their arg-min positions in sc[SC_ALPHA..] and k_compl_measure(affine, alpha_max) left the two complementarity sums in
// sc[SC_SUMS..]: the reference reads single elements of device arrays from the host (scalar indexing, "CUDA.@allowscalar"
// in its comment); here one thread does it on the device and leaves (alpha_p, alpha_d) in the scalar block.
__global__ void k_mehrotra_step(V v, const int64_t *ind_lb, const int64_t *ind_ub, int base, double gamma_f, double *sc)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    const double gamma_a = 1.0 / (1.0 - gamma_f);
    const double axl = sc[SC_ALPHA + 0], axu = sc[SC_ALPHA + 1], azl = sc[SC_ALPHA + 2], azu = sc[SC_ALPHA + 3];
    const int64_t ixl = (int64_t)sc[SC_ALPHA + 4], ixu = (int64_t)sc[SC_ALPHA + 5];
    const int64_t izl = (int64_t)sc[SC_ALPHA + 6], izu = (int64_t)sc[SC_ALPHA + 7];
    const double max_ap = fmin(axl, axu), max_ad = fmin(azl, azu);
    const double cnt = (double)(v.nlb + v.nub);
    double mu_full = (cnt > 0) ? (sc[SC_SUMS + 0] + sc[SC_SUMS + 1]) / cnt : 0.0;
    mu_full /= gamma_a;
    double ap = 1.0, ad = 1.0;
    if (max_ap < 1.0) {
        if (axl <= axu) {
            const int64_t j = ixl - 1, i = ind_lb[j] - base;
            const double tmp = mu_full / (v.zl[i] + max_ad * dzl[j]);
            ap = (v.x[i] - v.xl[i] - tmp) / (-dx[i]);
        } else {
            const int64_t j = ixu - 1, i = ind_ub[j] - base;
            const double tmp = mu_full / (v.zu[i] + max_ad * dzu[j]);
            ap = (v.xu[i] - v.x[i] - tmp) / (dx[i]);
        }
    }
    if (max_ad < 1.0) {
        if (azl <= azu) {
            const int64_t j = izl - 1, i = ind_lb[j] - base;
            const double tmp = mu_full / (v.x[i] + max_ap * dx[i] - v.xl[i]);
            ad = -(v.zl[i] - tmp) / dzl[j];
        } else {
            const int64_t j = izu - 1, i = ind_ub[j] - base;
            const double tmp = mu_full / (v.xu[i] - v.x[i] - max_ap * dx[i]);
            ad = -(v.zu[i] - tmp) / dzu[j];
        }
    }
    sc[SC_ALPHA_P] = fmax(ap, gamma_f * max_ap);
    sc[SC_ALPHA_D] = fmax(ad, gamma_f * max_ad);
    sc[SC_TAU] = 1.0;
}

// The three launches of the rule; leaves alpha_p / alpha_d in sc (device).
static int mehrotra_step_launch(Handle *h, const V &v, unsigned g, double gamma_f, double *sc)
{
    k_alpha_max<<<g, TB, 0, h->stream>>>(v, 1.0, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, nullptr, 0.0, sc);
    MIPM_CHECK_LAUNCH(h);
    if (v.nlb + v.nub > 0) {
        k_compl_measure<<<g, TB, 0, h->stream>>>(v, 1, 0.0, 0.0, h->d_partials.p, h->d_counter.p, sc + SC_SUMS, sc);
        MIPM_CHECK_LAUNCH(h);
    }
    k_mehrotra_step<<<1, 32, 0, h->stream>>>(v, h->v.d_ind_lb, h->v.d_ind_ub, h->v.index_base, gamma_f, sc);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

// src/solver.jl:194-205 + src/kernels.jl:408-430 + src/structure.jl:193
// K2.5 (MadNLP.ScaledSparseKKTSystem): src/kernels.jl:139-149 + MadNLP._set_aug_diagonal! (un-vendored, restated from the
// symmetric scaling K2.5 = S K2 S with S = diag(sqrt((x - xl)(xu - x))), absent factors = 1): l_diag = x - xl and
// u_diag = xu - x are POSITIVE here (sign flipped against K2), pr_diag = zu (x - xl) + zl (xu - x) + reg S^2.
__global__ void __launch_bounds__(TB) k_set_aug_diag_scaled(V v, double del_w, double del_c, double *sf)
{
    GRID_STRIDE(i, v.n) {
        v.reg[i] = del_w;
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        double xlzu = 0.0, xuzl = 0.0, s = 1.0;
        double ld = 1.0, ud = 1.0;
        if (jl >= 0) { ld = v.x[i] - v.xl[i]; v.l_diag[jl] = ld; v.l_lower[jl] = v.zl[i]; xuzl = v.zl[i]; s *= sqrt(ld); }
        if (ju >= 0) { ud = v.xu[i] - v.x[i]; v.u_diag[ju] = ud; v.u_lower[ju] = v.zu[i]; xlzu = v.zu[i]; s *= sqrt(ud); }
        if (jl >= 0) xlzu *= ld;        // (X - Xl) zu   (zero without an upper bound)
        if (ju >= 0) xuzl *= ud;        // (Xu - X) zl   (zero without a lower bound)
        sf[i] = s;
        v.pr_diag[i] = (xlzu + xuzl) + del_w * (s * s);
    }
    GRID_STRIDE(i, v.m) v.du_diag[i] = del_c;
}

// K2.5 build_kkt!: Hessian entries times S_i S_j, Jacobian entries (slack columns included) times S_col.
__global__ void __launch_bounds__(TB) k_k25_scale(int64_t nnzh, const int32_t *hi, const int32_t *hj, const double *hraw, double *hout,
                                                  int64_t nnzj, const int32_t *jj, const double *jraw, double *jout, int base,
                                                  const double *sf)
{
    GRID_STRIDE(q, nnzh) hout[q] = hraw[q] * sf[hi[q] - base] * sf[hj[q] - base];
    GRID_STRIDE(q, nnzj) jout[q] = jraw[q] * sf[jj[q] - base];
}

// K2.5 reduce_rhs! (positive diagonals) followed by the scaling of the primal block; finish: unscale, then the bound duals.
__global__ void __launch_bounds__(TB) k_reduce_rhs_scaled(V v, double *w, const double *sf)
{
    double *wx = w, *wzl = w + v.n + v.m, *wzu = w + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double t = wx[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) t = t + wzl[jl] / v.l_diag[jl];
        if (ju >= 0) t = t + wzu[ju] / v.u_diag[ju];
        wx[i] = t * sf[i];
    }
}
__global__ void __launch_bounds__(TB) k_finish_aug_scaled(V v, double *w, const double *sf)
{
    double *wx = w, *wzl = w + v.n + v.m, *wzu = w + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        const double t = wx[i] * sf[i];
        wx[i] = t;
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) wzl[jl] = (wzl[jl] - v.l_lower[jl] * t) / v.l_diag[jl];
        if (ju >= 0) wzu[ju] = (-wzu[ju] + v.u_lower[ju] * t) / v.u_diag[ju];
    }
}
// _kktmul! on the UNSCALED unreduced system with the positive K2.5 diagonals (l_diag = -(xl - x), u_diag = -(x - xu)).
__global__ void __launch_bounds__(TB) k_kktmul_scaled(V v, double *w, const double *vv, double alpha, double beta)
{
    double *wx = w, *wy = w + v.n, *wzl = w + v.n + v.m, *wzu = w + v.n + v.m + v.nlb;
    const double *vx = vv, *vy = vv + v.n, *vzl = vv + v.n + v.m, *vzu = vv + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double t = wx[i] + alpha * v.reg[i] * vx[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            t = t - alpha * vzl[jl];
            wzl[jl] = beta * wzl[jl] + alpha * (vx[i] * v.l_lower[jl] + vzl[jl] * v.l_diag[jl]);
        }
        if (ju >= 0) {
            t = t + alpha * vzu[ju];
            wzu[ju] = beta * wzu[ju] + alpha * (vx[i] * v.u_lower[ju] - vzu[ju] * v.u_diag[ju]);
        }
        wx[i] = t;
    }
    GRID_STRIDE(i, v.m) wy[i] = wy[i] + alpha * v.du_diag[i] * vy[i];
}

// out = (-y.rhs, zl_r.xl_r, zu_r.xu_r, ||c||inf, ||f-zl+zu+jacl||inf, max compl, ||dx||inf)
__global__ void __launch_bounds__(TB) k_termination(V v, double *partials, unsigned int *counter, double *out)
{
    double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    GRID_STRIDE(i, v.n) {
        double x = v.x[i], zl = v.zl[i], zu = v.zu[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        double rd = fabs(((v.f[i] - zl) + zu) + v.jacl[i]);
        acc[4] = comb(acc[4], rd, OP_MAX);
        acc[6] = comb(acc[6], fabs(v.d[i]), OP_MAX);
        if (jl >= 0) {
            double xl = v.xl[i];
            acc[1] += zl * xl;
            acc[5] = comb(acc[5], fabs((x - xl) * zl), OP_MAX);
        }
        if (ju >= 0) {
            double xu = v.xu[i];
            acc[2] += zu * xu;
            acc[5] = comb(acc[5], fabs((xu - x) * zu), OP_MAX);
        }
    }
    GRID_STRIDE(i, v.m) {
        acc[0] -= v.y[i] * v.rhs[i];
        acc[3] = comb(acc[3], fabs(v.c[i]), OP_MAX);
    }
    const int op[7] = {OP_SUM, OP_SUM, OP_SUM, OP_MAX, OP_MAX, OP_MAX, OP_MAX};
    grid_reduce_vals<7>(acc, op, partials, counter, out);
}

// src/solver.jl:308-317 + MadNLP.adjust_boundary!
__global__ void __launch_bounds__(TB) k_apply_step(V v, double ap, double ad, double c1, double c2, const double *sc)
{
    if (sc) { ap = sc[SC_ALPHA_P]; ad = sc[SC_ALPHA_D]; c1 = DBL_EPSILON * sc[SC_MU]; }
    const double *dx = v.d, *dy = v.d + v.n, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double x = v.x[i] + ap * dx[i];
        v.x[i] = x;
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            v.zl[i] = v.zl[i] + ad * dzl[jl];
            double xl = v.xl[i];
            if (x - xl < c1) v.xl[i] = xl - c2 * fmax(1.0, fabs(x));
        }
        if (ju >= 0) {
            v.zu[i] = v.zu[i] + ad * dzu[ju];
            double xu = v.xu[i];
            if (xu - x < c1) v.xu[i] = xu + c2 * fmax(1.0, fabs(x));
        }
    }
    GRID_STRIDE(i, v.m) v.y[i] = v.y[i] + ad * dy[i];
}

// MadNLP.reduce_rhs! (+ optionally r1 = wx / Sigma, r2 = wy: normalkkt.jl:197-207)
__global__ void __launch_bounds__(TB) k_reduce_rhs(V v, double *w, double *buf_n, double *buf_m)
{
    double *wx = w, *wy = w + v.n, *wzl = w + v.n + v.m, *wzu = w + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double t = wx[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) t = t - wzl[jl] / v.l_diag[jl];
        if (ju >= 0) t = t - wzu[ju] / v.u_diag[ju];
        wx[i] = t;
        if (buf_n) buf_n[i] = t / v.pr_diag[i];
    }
    if (buf_m) { GRID_STRIDE(i, v.m) buf_m[i] = wy[i]; }
}

// normalkkt.jl:212-213: wy = dy (from buffer_m), r1 = wx
__global__ void __launch_bounds__(TB) k_normal_mid(V v, double *w, double *buf_n, const double *buf_m)
{
    double *wx = w, *wy = w + v.n;
    GRID_STRIDE(i, v.m) wy[i] = buf_m[i];
    GRID_STRIDE(i, v.n) buf_n[i] = wx[i];
}

// MadNLP.finish_aug_solve! (+ optionally wx = r1 / Sigma first: normalkkt.jl:215-217)
__global__ void __launch_bounds__(TB) k_finish_aug(V v, double *w, const double *buf_n)
{
    double *wx = w, *wzl = w + v.n + v.m, *wzu = w + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double t = buf_n ? buf_n[i] / v.pr_diag[i] : wx[i];
        if (buf_n) wx[i] = t;
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) wzl[jl] = (-wzl[jl] + v.l_lower[jl] * t) / v.l_diag[jl];
        if (ju >= 0) wzu[ju] = (wzu[ju] - v.u_lower[ju] * t) / v.u_diag[ju];
    }
}

// MadNLP._kktmul! (App. B)
__global__ void __launch_bounds__(TB) k_kktmul(V v, double *w, const double *vv, double alpha, double beta)
{
    double *wx = w, *wy = w + v.n, *wzl = w + v.n + v.m, *wzu = w + v.n + v.m + v.nlb;
    const double *vx = vv, *vy = vv + v.n, *vzl = vv + v.n + v.m, *vzu = vv + v.n + v.m + v.nlb;
    GRID_STRIDE(i, v.n) {
        double t = wx[i] + alpha * v.reg[i] * vx[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            t = t - alpha * vzl[jl];
            wzl[jl] = beta * wzl[jl] + alpha * (vx[i] * v.l_lower[jl] - vzl[jl] * v.l_diag[jl]);
        }
        if (ju >= 0) {
            t = t + alpha * vzu[ju];
            wzu[ju] = beta * wzu[ju] + alpha * (vx[i] * v.u_lower[ju] + vzu[ju] * v.u_diag[ju]);
        }
        wx[i] = t;
    }
    GRID_STRIDE(i, v.m) wy[i] = wy[i] + alpha * v.du_diag[i] * vy[i];
}

__global__ void __launch_bounds__(TB) k_two_norms(int64_t len, const double *a, const double *b, double *partials,
                                                  unsigned int *counter, double *out)
{
    double acc[2] = {0.0, 0.0};
    GRID_STRIDE(i, len) {
        acc[0] = comb(acc[0], fabs(a[i]), OP_MAX);
        acc[1] = comb(acc[1], fabs(b[i]), OP_MAX);
    }
    const int op[2] = {OP_MAX, OP_MAX};
    grid_reduce_vals<2>(acc, op, partials, counter, out);
}

// init_starting_point! stages (src/solver.jl:41-118)
__global__ void __launch_bounds__(TB) k_init0(V v, double *partials, unsigned int *counter, double *out)
{
    // res = jacl (holds A'y + f); zl/zu init (solver.jl:41-66); mins for delta_x / delta_s (:68-78)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    GRID_STRIDE(i, v.n) {
        double r = v.jacl[i], l = v.xl[i], u = v.xu[i], x = v.x[i];
        bool fl = isfinite(l), fu = isfinite(u);
        double zl = v.zl[i], zu = v.zu[i];
        if (fl && fu) { zl = 0.5 * r; zu = -0.5 * r; }
        else if (fl) zl = r;
        else if (fu) zu = -r;
        v.zl[i] = zl;
        v.zu[i] = zu;
        if (v.inv_lb[i] >= 0) { acc[0] = fmin(acc[0], x - l); acc[2] = fmin(acc[2], zl); }
        if (v.inv_ub[i] >= 0) { acc[1] = fmin(acc[1], u - x); acc[3] = fmin(acc[3], zu); }
    }
    const int op[4] = {OP_MIN, OP_MIN, OP_MIN, OP_MIN};
    grid_reduce_vals<4>(acc, op, partials, counter, out);
}
__global__ void __launch_bounds__(TB) k_init1(V v, double delta_x, double delta_s, double *partials, unsigned int *counter, double *out)
{
    // shifts (:80-83) then mu pieces and sums (:85-94)
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    GRID_STRIDE(i, v.n) {
        bool lb = v.inv_lb[i] >= 0, ub = v.inv_ub[i] >= 0;
        double x = v.x[i];
        if (lb) x = x + delta_x;
        if (ub) x = x - delta_x;
        v.x[i] = x;
        if (lb) {
            double zl = v.zl[i] + (1.0 + delta_s);
            v.zl[i] = zl;
            double l = v.xl[i];
            acc[0] += x * zl - l * zl;
            acc[1] += zl;
            acc[3] += x - l;
        }
        if (ub) {
            double zu = v.zu[i] + (1.0 + delta_s);
            v.zu[i] = zu;
            double u = v.xu[i];
            acc[0] += u * zu - x * zu;
            acc[2] += zu;
            acc[4] += u - x;
        }
    }
    const int op[5] = {OP_SUM, OP_SUM, OP_SUM, OP_SUM, OP_SUM};
    grid_reduce_vals<5>(acc, op, partials, counter, out);
}
__global__ void __launch_bounds__(TB) k_init2(V v, double delta_x2, double delta_s2, double kappa, double *partials,
                                              unsigned int *counter, double *out)
{
    // second shifts (:96-99), projection (:102-118), interior checks (:120-123)
    double acc[4] = {DBL_MAX, DBL_MAX, DBL_MAX, DBL_MAX};
    GRID_STRIDE(i, v.n) {
        bool lb = v.inv_lb[i] >= 0, ub = v.inv_ub[i] >= 0;
        double x = v.x[i], l = v.xl[i], u = v.xu[i];
        if (lb) x = x + delta_x2;
        if (ub) x = x - delta_x2;
        if (x < l) x = l + fmin(kappa * fmax(1.0, l), kappa * (u - l));
        else if (u < x) x = u - fmin(kappa * fmax(1.0, u), kappa * (u - l));
        v.x[i] = x;
        if (lb) { double zl = v.zl[i] + delta_s2; v.zl[i] = zl; acc[0] = fmin(acc[0], zl); acc[2] = fmin(acc[2], x - l); }
        if (ub) { double zu = v.zu[i] + delta_s2; v.zu[i] = zu; acc[1] = fmin(acc[1], zu); acc[3] = fmin(acc[3], u - x); }
    }
    const int op[4] = {OP_MIN, OP_MIN, OP_MIN, OP_MIN};
    grid_reduce_vals<4>(acc, op, partials, counter, out);
}

__global__ void __launch_bounds__(TB) k_axpby(int64_t n, double alpha, const double *x, double beta, double *y)
{
    GRID_STRIDE(i, n) y[i] = (beta == 0.0) ? alpha * x[i] : alpha * x[i] + beta * y[i];
}
__global__ void __launch_bounds__(TB) k_fill(int64_t n, double val, double *x)
{
    GRID_STRIDE(i, n) x[i] = val;
}
__global__ void __launch_bounds__(TB) k_gather(int64_t n, const double *src, const int64_t *map, int base, double *dst)
{
    GRID_STRIDE(i, n) dst[i] = src[map[i] - base];
}
__global__ void __launch_bounds__(TB) k_scatter(int64_t n, const double *src, const int64_t *map, int base, double *dst)
{
    GRID_STRIDE(i, n) dst[map[i] - base] = src[i];
}
__global__ void __launch_bounds__(TB) k_dot(int64_t n, const double *x, const double *y, double *partials, unsigned int *counter, double *out)
{
    double acc[1] = {0.0};
    GRID_STRIDE(i, n) acc[0] = fma(x[i], y[i], acc[0]);
    const int op[1] = {OP_SUM};
    grid_reduce_vals<1>(acc, op, partials, counter, out);
}

__global__ void __launch_bounds__(TB) k_amax(int64_t n, const double *x, double *partials, unsigned int *counter, double *out)
{
    double acc[1] = {0.0};
    GRID_STRIDE(i, n) acc[0] = comb(acc[0], fabs(x[i]), OP_MAX);
    const int op[1] = {OP_MAX};
    grid_reduce_vals<1>(acc, op, partials, counter, out);
}
// MadNLP bound relaxation + initialize_variables! push (App. B). Every product / sum is rounded on its own
// (__dmul_rn / __dadd_rn are never contracted) so the values match the reference's broadcasts bit for bit.
__global__ void __launch_bounds__(TB) k_init_bounds(int64_t n, double tol, double bp, double bf, double *x, double *xl, double *xu)
{
    GRID_STRIDE(i, n) {
        const double l = __dadd_rn(xl[i], -__dmul_rn(fmax(1.0, fabs(xl[i])), tol));
        const double u = __dadd_rn(xu[i], __dmul_rn(fmax(1.0, fabs(xu[i])), tol));
        xl[i] = l; xu[i] = u;
        const bool fl = isfinite(l), fu = isfinite(u);
        double xi = x[i];
        if (fl && fu) {
            const double w = __dmul_rn(bf, __dadd_rn(u, -l));
            const double pl = fmin(__dmul_rn(bp, fmax(1.0, fabs(l))), w);
            const double pu = fmin(__dmul_rn(bp, fmax(1.0, fabs(u))), w);
            xi = fmax(__dadd_rn(l, pl), fmin(__dadd_rn(u, -pu), xi));
        } else if (fl) {
            xi = fmax(__dadd_rn(l, __dmul_rn(bp, fmax(1.0, fabs(l)))), xi);
        } else if (fu) {
            xi = fmin(__dadd_rn(u, -__dmul_rn(bp, fmax(1.0, fabs(u)))), xi);
        }
        x[i] = xi;
    }
}

}  // namespace


// ------------------------------------------------------------------ batches of independent problems (BASELINE config C5)
// B independent LPs / QPs are STACKED into one block-diagonal problem: every element-wise kernel above, the SpMVs, the
// KKT assembly and the task-graph factorization / solves (a forest: one elimination tree per unit) run unchanged on the
// stacked vectors, so all units advance through ONE set of launches per IPM phase. What is per unit -- reductions, step
// lengths, centering parameter, barrier value, termination -- is done by the kernels below: one CTA per unit (block
// reductions only, deterministic), unit u owning variables [off_n[u], off_n[u+1]) and rows [off_m[u], off_m[u+1]) and a
// scalar block of its own. A unit that has terminated stops moving (its step is skipped).
struct UB {
    const int64_t *off_n, *off_m;
    const int *active;
    double *sc;         // B x SC_COUNT
    double *uin;        // B x 4: per-unit inputs from the host (init stages)
    double *uout;       // B x 8: per-unit outputs to the host (init stages, norms)
};
#define UNIT_STRIDE(i, off) for (int64_t i = (off)[blockIdx.x] + threadIdx.x; i < (off)[blockIdx.x + 1]; i += TB)

template <int NV>
__device__ void block_reduce_vals(double (&acc)[NV], const int (&op)[NV], double *out)
{
    __shared__ double sm[NV][TB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] = comb(acc[q], __shfl_xor_sync(0xffffffffu, acc[q], o), op[q]);
        if (lane == 0) sm[q][warp] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        const int q = threadIdx.x;
        double v = sm[q][0];
        for (int w = 1; w < TB / 32; ++w) v = comb(v, sm[q][w], op[q]);
        out[q] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TB) kb_amax(const double *x, const int64_t *off, double *out, int stride)
{
    double acc[1] = {0.0};
    UNIT_STRIDE(i, off) acc[0] = comb(acc[0], fabs(x[i]), OP_MAX);
    const int op[1] = {OP_MAX};
    block_reduce_vals<1>(acc, op, out + (size_t)blockIdx.x * stride);
}
__global__ void __launch_bounds__(TB) kb_dot(const double *x, const double *y, const int64_t *off, double *out, int stride)
{
    double acc[1] = {0.0};
    UNIT_STRIDE(i, off) acc[0] += x[i] * y[i];
    const int op[1] = {OP_SUM};
    block_reduce_vals<1>(acc, op, out + (size_t)blockIdx.x * stride);
}
// k_init0 / k_init1 / k_init2 per unit (src/solver.jl:41-118)
__global__ void __launch_bounds__(TB) kb_init0(V v, UB ub)
{
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    UNIT_STRIDE(i, ub.off_n) {
        double r = v.jacl[i], l = v.xl[i], u = v.xu[i], x = v.x[i];
        bool fl = isfinite(l), fu = isfinite(u);
        double zl = v.zl[i], zu = v.zu[i];
        if (fl && fu) { zl = 0.5 * r; zu = -0.5 * r; }
        else if (fl) zl = r;
        else if (fu) zu = -r;
        v.zl[i] = zl;
        v.zu[i] = zu;
        if (v.inv_lb[i] >= 0) { acc[0] = fmin(acc[0], x - l); acc[2] = fmin(acc[2], zl); }
        if (v.inv_ub[i] >= 0) { acc[1] = fmin(acc[1], u - x); acc[3] = fmin(acc[3], zu); }
    }
    const int op[4] = {OP_MIN, OP_MIN, OP_MIN, OP_MIN};
    block_reduce_vals<4>(acc, op, ub.uout + 8 * (size_t)blockIdx.x);
}
__global__ void __launch_bounds__(TB) kb_init1(V v, UB ub)
{
    const double delta_x = ub.uin[4 * (size_t)blockIdx.x], delta_s = ub.uin[4 * (size_t)blockIdx.x + 1];
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    UNIT_STRIDE(i, ub.off_n) {
        bool lb = v.inv_lb[i] >= 0, ubd = v.inv_ub[i] >= 0;
        double x = v.x[i];
        if (lb) x = x + delta_x;
        if (ubd) x = x - delta_x;
        v.x[i] = x;
        if (lb) {
            double zl = v.zl[i] + (1.0 + delta_s);
            v.zl[i] = zl;
            double l = v.xl[i];
            acc[0] += x * zl - l * zl;
            acc[1] += zl;
            acc[3] += x - l;
        }
        if (ubd) {
            double zu = v.zu[i] + (1.0 + delta_s);
            v.zu[i] = zu;
            double u = v.xu[i];
            acc[0] += u * zu - x * zu;
            acc[2] += zu;
            acc[4] += u - x;
        }
    }
    const int op[5] = {OP_SUM, OP_SUM, OP_SUM, OP_SUM, OP_SUM};
    block_reduce_vals<5>(acc, op, ub.uout + 8 * (size_t)blockIdx.x);
}
__global__ void __launch_bounds__(TB) kb_init2(V v, UB ub, double kappa)
{
    const double delta_x2 = ub.uin[4 * (size_t)blockIdx.x], delta_s2 = ub.uin[4 * (size_t)blockIdx.x + 1];
    double acc[4] = {DBL_MAX, DBL_MAX, DBL_MAX, DBL_MAX};
    UNIT_STRIDE(i, ub.off_n) {
        bool lb = v.inv_lb[i] >= 0, ubd = v.inv_ub[i] >= 0;
        double x = v.x[i], l = v.xl[i], u = v.xu[i];
        if (lb) x = x + delta_x2;
        if (ubd) x = x - delta_x2;
        if (x < l) x = l + fmin(kappa * fmax(1.0, l), kappa * (u - l));
        else if (u < x) x = u - fmin(kappa * fmax(1.0, u), kappa * (u - l));
        v.x[i] = x;
        if (lb) { double zl = v.zl[i] + delta_s2; v.zl[i] = zl; acc[0] = fmin(acc[0], zl); acc[2] = fmin(acc[2], x - l); }
        if (ubd) { double zu = v.zu[i] + delta_s2; v.zu[i] = zu; acc[1] = fmin(acc[1], zu); acc[3] = fmin(acc[3], u - x); }
    }
    const int op[4] = {OP_MIN, OP_MIN, OP_MIN, OP_MIN};
    block_reduce_vals<4>(acc, op, ub.uout + 8 * (size_t)blockIdx.x);
}
// k_termination per unit; also counts the unit's lower / upper bounds into sc[SC_SUMS + 4..5] (dual objective terms)
__global__ void __launch_bounds__(TB) kb_termination(V v, UB ub)
{
    double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    UNIT_STRIDE(i, ub.off_n) {
        double x = v.x[i], zl = v.zl[i], zu = v.zu[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        double rd = fabs(((v.f[i] - zl) + zu) + v.jacl[i]);
        acc[4] = comb(acc[4], rd, OP_MAX);
        acc[6] = comb(acc[6], fabs(v.d[i]), OP_MAX);
        if (jl >= 0) {
            double xl = v.xl[i];
            acc[1] += zl * xl;
            acc[5] = comb(acc[5], fabs((x - xl) * zl), OP_MAX);
        }
        if (ju >= 0) {
            double xu = v.xu[i];
            acc[2] += zu * xu;
            acc[5] = comb(acc[5], fabs((xu - x) * zu), OP_MAX);
        }
    }
    UNIT_STRIDE(i, ub.off_m) {
        acc[0] -= v.y[i] * v.rhs[i];
        acc[3] = comb(acc[3], fabs(v.c[i]), OP_MAX);
    }
    const int op[7] = {OP_SUM, OP_SUM, OP_SUM, OP_MAX, OP_MAX, OP_MAX, OP_MAX};
    block_reduce_vals<7>(acc, op, ub.sc + (size_t)blockIdx.x * SC_COUNT + SC_TERM);
}
// set_correction_rhs! with the unit's own barrier value (the predictive right-hand side has no per-unit scalar)
__global__ void __launch_bounds__(TB) kb_set_rhs_corr(V v, UB ub)
{
    const double mu = ub.sc[(size_t)blockIdx.x * SC_COUNT + SC_MU];
    double *px = v.p, *py = v.p + v.n, *pzl = v.p + v.n + v.m, *pzu = v.p + v.n + v.m + v.nlb;
    UNIT_STRIDE(i, ub.off_n) {
        double x = v.x[i], zl = v.zl[i], zu = v.zu[i];
        px[i] = ((-v.f[i] + zl) - zu) - v.jacl[i];
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) pzl[jl] = ((v.xl[i] - x) * zl + mu) - v.corr_lb[jl];
        if (ju >= 0) pzu[ju] = ((v.xu[i] - x) * zu - mu) - v.corr_ub[ju];
    }
    UNIT_STRIDE(i, ub.off_m) py[i] = -v.c[i];
}
// k_alpha_max per unit. mode 0: tau given; mode 1: AdaptiveStep tau = max(1 - mu_u, tau)
__global__ void __launch_bounds__(TB) kb_alpha_max(V v, UB ub, int mode, double tau)
{
    double *sc = ub.sc + (size_t)blockIdx.x * SC_COUNT;
    if (mode == 1) tau = fmax(1.0 - sc[SC_MU], tau);
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    double acc[4] = {1.0, 1.0, 1.0, 1.0};
    UNIT_STRIDE(i, ub.off_n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        double dxi = dx[i];
        if (jl >= 0) {
            if (dxi < 0.0) acc[0] = fmin(acc[0], (-v.x[i] + v.xl[i]) * tau / dxi);
            double dz = dzl[jl];
            if (dz < 0.0) acc[2] = fmin(acc[2], (-v.zl[i]) * tau / dz);
        }
        if (ju >= 0) {
            if (dxi > 0.0) acc[1] = fmin(acc[1], (-v.x[i] + v.xu[i]) * tau / dxi);
            double dz = dzu[ju], zu = v.zu[i];
            if ((dz < 0.0) && (zu + dz < 0.0)) acc[3] = fmin(acc[3], (-zu) * tau / dz);
        }
    }
    const int op[4] = {OP_MIN, OP_MIN, OP_MIN, OP_MIN};
    block_reduce_vals<4>(acc, op, sc + SC_ALPHA);
    if (threadIdx.x == 0) {
        sc[SC_ALPHA_P] = fmin(sc[SC_ALPHA + 0], sc[SC_ALPHA + 1]);
        sc[SC_ALPHA_D] = fmin(sc[SC_ALPHA + 2], sc[SC_ALPHA + 3]);
        sc[SC_TAU] = tau;
    }
}
// k_predictor_measures per unit
__global__ void __launch_bounds__(TB) kb_predictor_measures(V v, UB ub, double mu_min)
{
    double *sc = ub.sc + (size_t)blockIdx.x * SC_COUNT;
    const double *dx = v.d, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    const double ap = sc[SC_ALPHA_P], ad = sc[SC_ALPHA_D];
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    UNIT_STRIDE(i, ub.off_n) {
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            acc[0] += ((v.x[i] + ap * dx[i]) - v.xl[i]) * (v.zl[i] + ad * dzl[jl]);
            acc[2] += (v.x[i] - v.xl[i]) * v.zl[i];
            v.corr_lb[jl] = dx[i] * dzl[jl];
            acc[4] += 1.0;
        }
        if (ju >= 0) {
            acc[1] += (v.xu[i] - (v.x[i] + ap * dx[i])) * (v.zu[i] + ad * dzu[ju]);
            acc[3] += (v.xu[i] - v.x[i]) * v.zu[i];
            v.corr_ub[ju] = dx[i] * dzu[ju];
            acc[4] += 1.0;
        }
    }
    const int op[5] = {OP_SUM, OP_SUM, OP_SUM, OP_SUM, OP_SUM};
    block_reduce_vals<5>(acc, op, sc + SC_SUMS);
    if (threadIdx.x == 0) {
        const double cnt = sc[SC_SUMS + 4];         // number of bounds of this unit (exact in floating point)
        double mu_aff = 0.0, mu_cur = 0.0, sigma = 1.0;
        if (cnt > 0) {
            mu_aff = (sc[SC_SUMS + 0] + sc[SC_SUMS + 1]) / cnt;
            mu_cur = (sc[SC_SUMS + 2] + sc[SC_SUMS + 3]) / cnt;
            const double rr = mu_aff / mu_cur;
            sigma = fmin(fmax(rr * rr * rr, 1e-6), 10.0);
        }
        sc[SC_MU_AFF] = mu_aff;
        sc[SC_MU_CURR] = mu_cur;
        sc[SC_MU] = fmax(mu_min, sigma * mu_cur);
    }
}
// k_apply_step per unit; a unit that has terminated keeps its iterate
__global__ void __launch_bounds__(TB) kb_apply_step(V v, UB ub, double c2)
{
    if (ub.active && !ub.active[blockIdx.x]) return;
    const double *sc = ub.sc + (size_t)blockIdx.x * SC_COUNT;
    const double ap = sc[SC_ALPHA_P], ad = sc[SC_ALPHA_D], c1 = DBL_EPSILON * sc[SC_MU];
    const double *dx = v.d, *dy = v.d + v.n, *dzl = v.d + v.n + v.m, *dzu = v.d + v.n + v.m + v.nlb;
    UNIT_STRIDE(i, ub.off_n) {
        double x = v.x[i] + ap * dx[i];
        v.x[i] = x;
        int jl = v.inv_lb[i], ju = v.inv_ub[i];
        if (jl >= 0) {
            v.zl[i] = v.zl[i] + ad * dzl[jl];
            double xl = v.xl[i];
            if (x - xl < c1) v.xl[i] = xl - c2 * fmax(1.0, fabs(x));
        }
        if (ju >= 0) {
            v.zu[i] = v.zu[i] + ad * dzu[ju];
            double xu = v.xu[i];
            if (xu - x < c1) v.xu[i] = xu + c2 * fmax(1.0, fabs(x));
        }
    }
    UNIT_STRIDE(i, ub.off_m) v.y[i] = v.y[i] + ad * dy[i];
}

// ------------------------------------------------------------------ Ruiz equilibration (scripts/common.jl:57-100)
// scale_qp computes Dr, Dc with HSL.mc77(A, 0) (infinity-norm Ruiz scaling; HSL is closed source and not in the image: the
// published algorithm is restated -- Ruiz 2001; Knight, Ruiz, Ucar 2014): repeat  r_i = max_j |a_ij| / (Dr_i Dc_j),
// c_j = max_i |a_ij| / (Dr_i Dc_j);  Dr_i *= sqrt(r_i);  Dc_j *= sqrt(c_j)  (simultaneously) until max |1 - r|, |1 - c| <= tol
// or max_iter sweeps. One pass over the COO entries per sweep; the maxima are order independent (atomicMax on the bit
// pattern of non-negative doubles), so the result is deterministic and bit-identical to the sequential restatement.
__global__ void __launch_bounds__(TB) k_ruiz_max(int64_t nnz, const int32_t *ai, const int32_t *aj, const double *av, int base,
                                                 const double *dr, const double *dc, unsigned long long *rmax, unsigned long long *cmax)
{
    GRID_STRIDE(q, nnz) {
        const int i = ai[q] - base, j = aj[q] - base;
        const double v = (fabs(av[q]) / dr[i]) / dc[j];
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        atomicMax(rmax + i, bits);
        atomicMax(cmax + j, bits);
    }
}
__global__ void __launch_bounds__(TB) k_ruiz_apply(int64_t m, int64_t n, double *dr, double *dc, unsigned long long *rmax,
                                                   unsigned long long *cmax, double *partials, unsigned int *counter, double *out)
{
    double acc[1] = {0.0};
    GRID_STRIDE(i, m) {
        const double r = __longlong_as_double((long long)rmax[i]);
        rmax[i] = 0ull;
        if (r > 0.0) { dr[i] = dr[i] * sqrt(r); acc[0] = fmax(acc[0], fabs(1.0 - r)); }
    }
    GRID_STRIDE(j, n) {
        const double c = __longlong_as_double((long long)cmax[j]);
        cmax[j] = 0ull;
        if (c > 0.0) { dc[j] = dc[j] * sqrt(c); acc[0] = fmax(acc[0], fabs(1.0 - c)); }
    }
    const int op[1] = {OP_MAX};
    grid_reduce_vals<1>(acc, op, partials, counter, out);
}
// A.vals[k] /= Dr[i] Dc[j] (scripts/common.jl:37-44), in place or into `out`
__global__ void __launch_bounds__(TB) k_scale_coo(int64_t nnz, const int32_t *ai, const int32_t *aj, const double *av, int base,
                                                  const double *dr, const double *dc, double *out)
{
    GRID_STRIDE(q, nnz) out[q] = av[q] / (dr[ai[q] - base] * dc[aj[q] - base]);
}

static unsigned red_grid(Handle *h, int64_t len)
{
    int64_t g = (len + TB - 1) / TB;
    g = std::max<int64_t>(1, std::min<int64_t>(g, h->red_blocks));
    return (unsigned)g;
}

static V make_view(Handle *h, const int32_t *inv_lb, const int32_t *inv_ub)
{
    const mipm_mpc_vectors &m = h->v;
    V v;
    v.n = m.n; v.m = m.m; v.nlb = m.nlb; v.nub = m.nub;
    v.inv_lb = inv_lb; v.inv_ub = inv_ub;
    v.x = m.d_x; v.xl = m.d_xl; v.xu = m.d_xu; v.zl = m.d_zl; v.zu = m.d_zu; v.f = m.d_f;
    v.y = m.d_y; v.c = m.d_c; v.rhs = m.d_rhs; v.jacl = m.d_jacl; v.d = m.d_d; v.p = m.d_p; v.w = m.d_w;
    v.corr_lb = m.d_corr_lb; v.corr_ub = m.d_corr_ub;
    v.reg = m.d_reg; v.pr_diag = m.d_pr_diag; v.du_diag = m.d_du_diag;
    v.l_diag = m.d_l_diag; v.u_diag = m.d_u_diag; v.l_lower = m.d_l_lower; v.u_lower = m.d_u_lower;
    return v;
}

static DBuf<int32_t> &inv_lb_buf(Handle *h) { return h->d_inv_lb; }
static DBuf<int32_t> &inv_ub_buf(Handle *h) { return h->d_inv_ub; }

static int fetch_scalars(Handle *h, int count, double *out)
{
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_scal, h->d_scal.p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < count; ++i) out[i] = h->h_scal[i];
    return MIPM_OK;
}

}  // namespace mipm

using namespace mipm;

#define VIEW_OR_FAIL(h)                                                                     \
    MIPM_NEED_DEVICE(h);                                                                    \
    if (!(h)->bound) return fail((h), MIPM_ERR_STATE, "mipm_mpc_bind has not been called"); \
    MIPM_CUDA(h, cudaSetDevice((h)->device));                                               \
    V v = make_view((h), inv_lb_buf(h).p, inv_ub_buf(h).p);                                 \
    const int64_t nmax = std::max<int64_t>(std::max(v.n, v.m), 1);                          \
    (void)nmax

extern "C" {

int mipm_mpc_bind(mipm_handle hh, const mipm_mpc_vectors *mv)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!mv || mv->n < 0 || mv->m < 0 || mv->nlb < 0 || mv->nub < 0 || mv->nlb > mv->n || mv->nub > mv->n)
        return fail(h, MIPM_ERR_ARG, "bad sizes");
    if ((mv->nlb > 0 && !mv->d_ind_lb) || (mv->nub > 0 && !mv->d_ind_ub)) return fail(h, MIPM_ERR_ARG, "null index vector");
    MIPM_CUDA(h, cudaSetDevice(h->device));
    h->v = *mv;
    const int64_t n1 = std::max<int64_t>(mv->n, 1);
    MIPM_CUDA(h, inv_lb_buf(h).alloc((size_t)n1));
    MIPM_CUDA(h, inv_ub_buf(h).alloc((size_t)n1));
    MIPM_CUDA(h, cudaMemsetAsync(inv_lb_buf(h).p, 0xff, (size_t)n1 * sizeof(int32_t), h->stream));
    MIPM_CUDA(h, cudaMemsetAsync(inv_ub_buf(h).p, 0xff, (size_t)n1 * sizeof(int32_t), h->stream));
    if (mv->nlb > 0) {
        k_build_inv<<<red_grid(h, mv->nlb), TB, 0, h->stream>>>(mv->nlb, mv->d_ind_lb, mv->index_base, inv_lb_buf(h).p);
        MIPM_CHECK_LAUNCH(h);
    }
    if (mv->nub > 0) {
        k_build_inv<<<red_grid(h, mv->nub), TB, 0, h->stream>>>(mv->nub, mv->d_ind_ub, mv->index_base, inv_ub_buf(h).p);
        MIPM_CHECK_LAUNCH(h);
    }
    h->bound = true;
    return MIPM_OK;
}

int mipm_set_aug_diagonal_reg(mipm_handle hh, double del_w, double del_c)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    k_set_aug_diag<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, del_w, del_c);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_set_aug_diagonal_reg_scaled(mipm_handle hh, double del_w, double del_c, double *d_scaling_factor)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_scaling_factor && v.n > 0) return fail(h, MIPM_ERR_ARG, "null argument");
    k_set_aug_diag_scaled<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, del_w, del_c, d_scaling_factor);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_k25_scale_values(mipm_handle hh, int64_t nnzh, const int32_t *d_hess_i, const int32_t *d_hess_j, const double *d_hess_raw,
                          double *d_hess_out, int64_t nnzj, const int32_t *d_jac_j, const double *d_jac_raw, double *d_jac_out,
                          int index_base, const double *d_scaling_factor)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (nnzh < 0 || nnzj < 0 || (nnzh > 0 && (!d_hess_i || !d_hess_j || !d_hess_raw || !d_hess_out)) ||
        (nnzj > 0 && (!d_jac_j || !d_jac_raw || !d_jac_out)) || !d_scaling_factor)
        return fail(h, MIPM_ERR_ARG, "bad argument");
    if (nnzh + nnzj == 0) return MIPM_OK;
    k_k25_scale<<<red_grid(h, std::max(nnzh, nnzj)), TB, 0, h->stream>>>(nnzh, d_hess_i, d_hess_j, d_hess_raw, d_hess_out, nnzj, d_jac_j,
                                                                        d_jac_raw, d_jac_out, index_base, d_scaling_factor);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_reduce_rhs_scaled(mipm_handle hh, double *d_w, const double *d_scaling_factor)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w || !d_scaling_factor) return fail(h, MIPM_ERR_ARG, "null argument");
    k_reduce_rhs_scaled<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_scaling_factor);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_finish_aug_solve_scaled(mipm_handle hh, double *d_w, const double *d_scaling_factor)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w || !d_scaling_factor) return fail(h, MIPM_ERR_ARG, "null argument");
    k_finish_aug_scaled<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_scaling_factor);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_kktmul_scaled(mipm_handle hh, double *d_w, const double *d_v, double alpha, double beta)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w || !d_v) return fail(h, MIPM_ERR_ARG, "null argument");
    k_kktmul_scaled<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_v, alpha, beta);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_set_predictive_rhs(mipm_handle hh)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    k_set_rhs<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, 0, 0.0, nullptr);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_set_correction_rhs(mipm_handle hh, double mu)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    k_set_rhs<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, 1, mu, nullptr);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_get_correction(mipm_handle hh)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    k_get_correction<<<red_grid(h, nmax), TB, 0, h->stream>>>(v);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_set_extra_correction(mipm_handle hh, double alpha_p, double alpha_d, double beta_min, double beta_max, double mu)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    k_set_extra_correction<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, alpha_p, alpha_d, beta_min * mu, beta_max * mu);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

static int compl_measure(Handle *h, int affine, double ap, double ad, double *out)
{
    VIEW_OR_FAIL(h);
    if (!out) return fail(h, MIPM_ERR_ARG, "null argument");
    if (v.nlb + v.nub == 0) { *out = 0.0; return MIPM_OK; }
    k_compl_measure<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, affine, ap, ad, h->d_partials.p, h->d_counter.p, h->d_scal.p);
    MIPM_CHECK_LAUNCH(h);
    double s[2];
    int rc = fetch_scalars(h, 2, s);
    if (rc != MIPM_OK) return rc;
    *out = (s[0] + s[1]) / (double)(v.nlb + v.nub);
    return MIPM_OK;
}

int mipm_get_complementarity_measure(mipm_handle hh, double *out) { return compl_measure((Handle *)hh, 0, 0.0, 0.0, out); }

int mipm_get_affine_complementarity_measure(mipm_handle hh, double alpha_p, double alpha_d, double *out)
{
    return compl_measure((Handle *)hh, 1, alpha_p, alpha_d, out);
}

int mipm_get_alpha_max(mipm_handle hh, double tau, double *alpha, int64_t *idx)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!alpha) return fail(h, MIPM_ERR_ARG, "null argument");
    k_alpha_max<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, tau, h->d_partials.p, h->d_counter.p, h->d_scal.p, nullptr, 0.0, nullptr);
    MIPM_CHECK_LAUNCH(h);
    double s[8];
    int rc = fetch_scalars(h, 8, s);
    if (rc != MIPM_OK) return rc;
    for (int q = 0; q < 4; ++q) {
        alpha[q] = s[q];
        if (idx) idx[q] = (int64_t)s[4 + q];
    }
    return MIPM_OK;
}

int mipm_mehrotra_adaptive_step(mipm_handle hh, double gamma_f, double *alpha)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!alpha || !(gamma_f > 0.0 && gamma_f < 1.0)) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (!h->d_sc.p) {
        MIPM_CUDA(h, h->d_sc.alloc(SC_COUNT));
        MIPM_CUDA(h, cudaMemsetAsync(h->d_sc.p, 0, SC_COUNT * sizeof(double), h->stream));
    }
    int rc = mehrotra_step_launch(h, v, red_grid(h, nmax), gamma_f, h->d_sc.p);
    if (rc != MIPM_OK) return rc;
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_scal, h->d_sc.p + SC_ALPHA_P, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    alpha[0] = h->h_scal[0];
    alpha[1] = h->h_scal[1];
    return MIPM_OK;
}

int mipm_termination_measures(mipm_handle hh, double *out)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!out) return fail(h, MIPM_ERR_ARG, "null argument");
    k_termination<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, h->d_partials.p, h->d_counter.p, h->d_scal.p);
    MIPM_CHECK_LAUNCH(h);
    double s[7];
    int rc = fetch_scalars(h, 7, s);
    if (rc != MIPM_OK) return rc;
    // dual_objective (kernels.jl:408-417): dobj = -y.rhs; += zl.xl if nlb>0; -= zu.xu if nub>0
    double dobj = s[0];
    if (v.nlb > 0) dobj += s[1];
    if (v.nub > 0) dobj -= s[2];
    out[0] = dobj;
    out[1] = s[3];
    out[2] = s[4];
    out[3] = s[5];
    out[4] = s[6];
    return MIPM_OK;
}

int mipm_apply_step(mipm_handle hh, double alpha_p, double alpha_d, double mu)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    const double eps = DBL_EPSILON;
    k_apply_step<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, alpha_p, alpha_d, eps * mu, pow(eps, 0.75), nullptr);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_reduce_rhs(mipm_handle hh, double *d_w)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w) return fail(h, MIPM_ERR_ARG, "null argument");
    k_reduce_rhs<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, nullptr, nullptr);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_finish_aug_solve(mipm_handle hh, double *d_w)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w) return fail(h, MIPM_ERR_ARG, "null argument");
    k_finish_aug<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, nullptr);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_normal_solve_stage(mipm_handle hh, int stage, double *d_w, double *d_buffer_n, double *d_buffer_m)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w || !d_buffer_n || !d_buffer_m) return fail(h, MIPM_ERR_ARG, "null argument");
    if (stage == 0) k_reduce_rhs<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_buffer_n, d_buffer_m);
    else if (stage == 1) k_normal_mid<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_buffer_n, d_buffer_m);
    else if (stage == 2) k_finish_aug<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_buffer_n);
    else return fail(h, MIPM_ERR_ARG, "stage must be 0, 1 or 2");
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_kktmul(mipm_handle hh, double *d_w, const double *d_v, double alpha, double beta)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w || !d_v) return fail(h, MIPM_ERR_ARG, "null argument");
    k_kktmul<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, d_w, d_v, alpha, beta);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_residual_norms(mipm_handle hh, const double *d_w, const double *d_p, double *out)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!d_w || !d_p || !out) return fail(h, MIPM_ERR_ARG, "null argument");
    const int64_t len = v.n + v.m + v.nlb + v.nub;
    k_two_norms<<<red_grid(h, std::max<int64_t>(len, 1)), TB, 0, h->stream>>>(len, d_w, d_p, h->d_partials.p, h->d_counter.p, h->d_scal.p);
    MIPM_CHECK_LAUNCH(h);
    return fetch_scalars(h, 2, out);
}

int mipm_init_point_stage(mipm_handle hh, int stage, double a, double b, double kappa, double *out)
{
    Handle *h = (Handle *)hh;
    VIEW_OR_FAIL(h);
    if (!out) return fail(h, MIPM_ERR_ARG, "null argument");
    int cnt;
    if (stage == 0) { k_init0<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, h->d_partials.p, h->d_counter.p, h->d_scal.p); cnt = 4; }
    else if (stage == 1) { k_init1<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, a, b, h->d_partials.p, h->d_counter.p, h->d_scal.p); cnt = 5; }
    else if (stage == 2) { k_init2<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, a, b, kappa, h->d_partials.p, h->d_counter.p, h->d_scal.p); cnt = 4; }
    else return fail(h, MIPM_ERR_ARG, "stage must be 0, 1 or 2");
    MIPM_CHECK_LAUNCH(h);
    return fetch_scalars(h, cnt, out);
}

int mipm_axpby(mipm_handle hh, int64_t n, double alpha, const double *d_x, double beta, double *d_y)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || (n > 0 && (!d_x || !d_y))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) return MIPM_OK;
    k_axpby<<<red_grid(h, n), TB, 0, h->stream>>>(n, alpha, d_x, beta, d_y);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_fill(mipm_handle hh, int64_t n, double value, double *d_x)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || (n > 0 && !d_x)) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) return MIPM_OK;
    k_fill<<<red_grid(h, n), TB, 0, h->stream>>>(n, value, d_x);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_copy(mipm_handle hh, int64_t n, const double *d_src, double *d_dst)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || (n > 0 && (!d_src || !d_dst))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) return MIPM_OK;
    MIPM_CUDA(h, cudaMemcpyAsync(d_dst, d_src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return MIPM_OK;
}

int mipm_gather(mipm_handle hh, int64_t n, const double *d_src, const int64_t *d_map, int index_base, double *d_dst)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || (n > 0 && (!d_src || !d_map || !d_dst))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) return MIPM_OK;
    k_gather<<<red_grid(h, n), TB, 0, h->stream>>>(n, d_src, d_map, index_base, d_dst);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_scatter(mipm_handle hh, int64_t n, const double *d_src, const int64_t *d_map, int index_base, double *d_dst)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || (n > 0 && (!d_src || !d_map || !d_dst))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) return MIPM_OK;
    k_scatter<<<red_grid(h, n), TB, 0, h->stream>>>(n, d_src, d_map, index_base, d_dst);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_dot(mipm_handle hh, int64_t n, const double *d_x, const double *d_y, double *out)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || !out || (n > 0 && (!d_x || !d_y))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) { *out = 0.0; return MIPM_OK; }
    k_dot<<<red_grid(h, n), TB, 0, h->stream>>>(n, d_x, d_y, h->d_partials.p, h->d_counter.p, h->d_scal.p);
    MIPM_CHECK_LAUNCH(h);
    return fetch_scalars(h, 1, out);
}

int mipm_amax(mipm_handle hh, int64_t n, const double *d_x, double *out)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || !out || (n > 0 && !d_x)) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) { *out = 0.0; return MIPM_OK; }
    k_amax<<<red_grid(h, n), TB, 0, h->stream>>>(n, d_x, h->d_partials.p, h->d_counter.p, h->d_scal.p);
    MIPM_CHECK_LAUNCH(h);
    return fetch_scalars(h, 1, out);
}

int mipm_init_bounds(mipm_handle hh, int64_t n, double tol, double bound_push, double bound_fac,
                     double *d_x, double *d_xl, double *d_xu)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || (n > 0 && (!d_x || !d_xl || !d_xu))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (n == 0) return MIPM_OK;
    k_init_bounds<<<red_grid(h, n), TB, 0, h->stream>>>(n, tol, bound_push, bound_fac, d_x, d_xl, d_xu);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

// ---------------------------------------------------------------- fused iteration
static int fused_ready(Handle *h)
{
    if (!h->bound) return fail(h, MIPM_ERR_STATE, "mipm_mpc_bind has not been called");
    if (!h->has_model) return fail(h, MIPM_ERR_STATE, "mipm_mpc_set_model has not been called");
    if (!h->has_ls || !h->has_spmv) return fail(h, MIPM_ERR_STATE, "linear solver / SpMV not set up");
    return MIPM_OK;
}

int mipm_mpc_set_model(mipm_handle hh, const mipm_mpc_model *md)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->bound) return fail(h, MIPM_ERR_STATE, "mipm_mpc_bind has not been called");
    if (!h->has_spmv) return fail(h, MIPM_ERR_STATE, "mipm_spmv_setup has not been called");
    // zero-size pieces (no constraints, empty Jacobian) legitimately come with null pointers
    if (!md || (md->kkt_kind != 0 && md->kkt_kind != 1) || (!md->d_ATx && h->sp_nnz > 0) || (!md->d_cvec && h->v.n > 0) ||
        !md->d_aug_nz || (!md->d_buffer_n && h->v.n > 0) || (!md->d_buffer_m && h->v.m > 0) ||
        (md->kkt_kind == 1 && !md->d_aug_raw_V))
        return fail(h, MIPM_ERR_ARG, "bad model");
    if (md->d_Hx && !h->has_hess) return fail(h, MIPM_ERR_STATE, "Hessian values given but mipm_hess_setup not called");
    MIPM_CUDA(h, cudaSetDevice(h->device));
    h->model = *md;
    MIPM_CUDA(h, h->d_sc.alloc(SC_COUNT));
    MIPM_CUDA(h, cudaMemsetAsync(h->d_sc.p, 0, SC_COUNT * sizeof(double), h->stream));
    h->has_model = true;
    return MIPM_OK;
}

// build_kkt! + factorize! (MadNLP.factorize_wrapper!)
static int fused_factorize(Handle *h, double del_w, double del_c)
{
    int rc = mipm_set_aug_diagonal_reg((mipm_handle)h, del_w, del_c);
    if (rc != MIPM_OK) return rc;
    const mipm_mpc_model &md = h->model;
    if (md.kkt_kind == 0) rc = mipm_normal_assemble((mipm_handle)h, h->v.d_pr_diag, md.d_aug_nz, md.exact_order);
    else rc = mipm_k2_transfer((mipm_handle)h, md.d_aug_raw_V, md.d_aug_nz);
    if (rc != MIPM_OK) return rc;
    return mipm_ls_factorize_async((mipm_handle)h, md.d_aug_nz);
}

// solve!(kkt, w): normalkkt.jl:196-219 / MadNLP K2
static int fused_kkt_solve(Handle *h, double *w, int ir_steps)
{
    mipm_handle hh = (mipm_handle)h;
    const mipm_mpc_model &md = h->model;
    int rc;
    if (md.kkt_kind == 0) {
        if ((rc = mipm_normal_solve_stage(hh, 0, w, md.d_buffer_n, md.d_buffer_m)) != MIPM_OK) return rc;
        if ((rc = mipm_spmv(hh, 0, 1.0, md.d_ATx, md.d_buffer_n, -1.0, md.d_buffer_m)) != MIPM_OK) return rc;
        if ((rc = mipm_ls_solve(hh, md.d_buffer_m, ir_steps)) != MIPM_OK) return rc;
        if ((rc = mipm_normal_solve_stage(hh, 1, w, md.d_buffer_n, md.d_buffer_m)) != MIPM_OK) return rc;
        if ((rc = mipm_spmv(hh, 1, -1.0, md.d_ATx, md.d_buffer_m, 1.0, md.d_buffer_n)) != MIPM_OK) return rc;
        return mipm_normal_solve_stage(hh, 2, w, md.d_buffer_n, md.d_buffer_m);
    }
    if ((rc = mipm_reduce_rhs(hh, w)) != MIPM_OK) return rc;
    if ((rc = mipm_ls_solve(hh, w, ir_steps)) != MIPM_OK) return rc;
    return mipm_finish_aug_solve(hh, w);
}

// solve_system! (src/linear_solver.jl:19-44): d = K^-1 p, then w = p - K d and its norms into sc[slot..]
static int fused_solve_system(Handle *h, int ir_steps, int slot)
{
    mipm_handle hh = (mipm_handle)h;
    const mipm_mpc_vectors &m = h->v;
    const mipm_mpc_model &md = h->model;
    const int64_t N = m.n + m.m + m.nlb + m.nub;
    int rc;
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_d, m.d_p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = fused_kkt_solve(h, m.d_d, ir_steps)) != MIPM_OK) return rc;
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_w, m.d_p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    // mul!(w, kkt, d, -1, 1)
    if ((rc = mipm_spmv_pair(hh, md.d_ATx, -1.0, m.d_d, 1.0, m.d_w + m.n, -1.0, m.d_d + m.n, 1.0, m.d_w)) != MIPM_OK) return rc;
    if (md.d_Hx && md.kkt_kind == 1)
        if ((rc = mipm_hess_spmv(hh, -1.0, md.d_Hx, m.d_d, 1.0, m.d_w)) != MIPM_OK) return rc;
    if ((rc = mipm_kktmul(hh, m.d_w, m.d_d, -1.0, 1.0)) != MIPM_OK) return rc;
    k_two_norms<<<red_grid(h, std::max<int64_t>(N, 1)), TB, 0, h->stream>>>(N, m.d_w, m.d_p, h->d_partials.p, h->d_counter.p,
                                                                         h->d_sc.p + SC_RES + 2 * slot);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

static int fused_fetch(Handle *h, double *out, int *status)
{
    double sc[SC_COUNT];
    int info[4] = {0, 0, 0, 0};
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_scal, h->d_sc.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (status && h->d_info.p) MIPM_CUDA(h, cudaMemcpyAsync(info, h->d_info.p, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < SC_COUNT; ++i) sc[i] = h->h_scal[i];
    if (status) *status = (info[0] == 0) ? MIPM_OK : MIPM_ERR_NOT_FACTORIZED;
    if (out) {
        const mipm_mpc_vectors &m = h->v;
        double dobj = sc[SC_TERM + 0];
        if (m.nlb > 0) dobj += sc[SC_TERM + 1];
        if (m.nub > 0) dobj -= sc[SC_TERM + 2];
        out[0] = dobj;
        out[1] = sc[SC_TERM + 3];
        out[2] = sc[SC_TERM + 4];
        out[3] = sc[SC_TERM + 5];
        out[4] = sc[SC_TERM + 6];
        out[5] = sc[SC_OBJ];
        out[6] = sc[SC_OBJ + 1];
        out[7] = sc[SC_ALPHA_P];
        out[8] = sc[SC_ALPHA_D];
        out[9] = sc[SC_MU];
        out[10] = sc[SC_MU_CURR];
        out[11] = sc[SC_RES + 0];
        out[12] = sc[SC_RES + 1];
        out[13] = sc[SC_RES + 2];
        out[14] = sc[SC_RES + 3];
        out[15] = sc[SC_TAU];
    }
    return MIPM_OK;
}

int mipm_mpc_iter_begin(mipm_handle hh, double del_w, double del_c, double *out, int *status)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    int rc = fused_ready(h);
    if (rc != MIPM_OK) return rc;
    if (!out || !status) return fail(h, MIPM_ERR_ARG, "null argument");
    MIPM_CUDA(h, cudaSetDevice(h->device));
    V v = make_view(h, inv_lb_buf(h).p, inv_ub_buf(h).p);
    const int64_t nmax = std::max<int64_t>(std::max(v.n, v.m), 1);
    k_termination<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, h->d_partials.p, h->d_counter.p, h->d_sc.p + SC_TERM);
    MIPM_CHECK_LAUNCH(h);
    if ((rc = fused_factorize(h, del_w, del_c)) != MIPM_OK) return rc;
    return fused_fetch(h, out, status);
}

int mipm_mpc_peek(mipm_handle hh, double *out)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    // (also used around an external linear solver, where this handle has no factorization of its own)
    if (!h->bound) return fail(h, MIPM_ERR_STATE, "mipm_mpc_bind has not been called");
    if (!h->has_model) return fail(h, MIPM_ERR_STATE, "mipm_mpc_set_model has not been called");
    if (!out) return fail(h, MIPM_ERR_ARG, "null argument");
    MIPM_CUDA(h, cudaSetDevice(h->device));
    V v = make_view(h, inv_lb_buf(h).p, inv_ub_buf(h).p);
    const int64_t nmax = std::max<int64_t>(std::max(v.n, v.m), 1);
    k_termination<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, h->d_partials.p, h->d_counter.p, h->d_sc.p + SC_TERM);
    MIPM_CHECK_LAUNCH(h);
    return fused_fetch(h, out, nullptr);
}

int mipm_mpc_refactor(mipm_handle hh, double del_w, double del_c, int *status)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    int rc = fused_ready(h);
    if (rc != MIPM_OK) return rc;
    if (!status) return fail(h, MIPM_ERR_ARG, "null argument");
    MIPM_CUDA(h, cudaSetDevice(h->device));
    if ((rc = fused_factorize(h, del_w, del_c)) != MIPM_OK) return rc;
    return mipm_ls_status(hh, status);
}

int mipm_mpc_iter_rest(mipm_handle hh, double mu_min, int step_rule, double tau_param, int ir_steps)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    int rc = fused_ready(h);
    if (rc != MIPM_OK) return rc;
    if (step_rule < 0 || step_rule > 2) return fail(h, MIPM_ERR_ARG, "step_rule must be 0 (adaptive), 1 (conservative) or 2 (Mehrotra adaptive)");
    MIPM_CUDA(h, cudaSetDevice(h->device));
    V v = make_view(h, inv_lb_buf(h).p, inv_ub_buf(h).p);
    const mipm_mpc_model &md = h->model;
    const int64_t nmax = std::max<int64_t>(std::max(v.n, v.m), 1);
    const unsigned g = red_grid(h, nmax);
    double *sc = h->d_sc.p;
    // prediction_step! (solver.jl:230-237)
    k_set_rhs<<<g, TB, 0, h->stream>>>(v, 0, 0.0, nullptr);
    MIPM_CHECK_LAUNCH(h);
    if ((rc = fused_solve_system(h, ir_steps, 0)) != MIPM_OK) return rc;
    k_alpha_max<<<g, TB, 0, h->stream>>>(v, 1.0, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, nullptr, 0.0, sc);
    MIPM_CHECK_LAUNCH(h);
    k_predictor_measures<<<g, TB, 0, h->stream>>>(v, mu_min, h->d_partials.p, h->d_counter.p, sc);
    MIPM_CHECK_LAUNCH(h);
    // mehrotra_correction_direction! (solver.jl:239-243)
    k_set_rhs<<<g, TB, 0, h->stream>>>(v, 1, 0.0, sc + SC_MU);
    MIPM_CHECK_LAUNCH(h);
    if ((rc = fused_solve_system(h, ir_steps, 1)) != MIPM_OK) return rc;
    // update_step_size! (kernels.jl:291-305)
    if (step_rule == 2) {
        if ((rc = mehrotra_step_launch(h, v, g, tau_param, sc)) != MIPM_OK) return rc;      // tau_param carries gamma_f
    } else {
        if (step_rule == 0) k_alpha_max<<<g, TB, 0, h->stream>>>(v, 0.0, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, sc, tau_param, sc);
        else k_alpha_max<<<g, TB, 0, h->stream>>>(v, tau_param, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, nullptr, 0.0, sc);
        MIPM_CHECK_LAUNCH(h);
    }
    // apply_step! (solver.jl:308-317)
    k_apply_step<<<g, TB, 0, h->stream>>>(v, 0.0, 0.0, 0.0, pow(DBL_EPSILON, 0.75), sc);
    MIPM_CHECK_LAUNCH(h);
    // evaluate_model! (solver.jl:319-326): obj pieces, c(x) = A x - rhs, f = H x + c, jacl = A' y
    const mipm_mpc_vectors &m = h->v;
    k_dot<<<red_grid(h, std::max<int64_t>(md.nx, 1)), TB, 0, h->stream>>>(md.nx, md.d_cvec, m.d_x, h->d_partials.p, h->d_counter.p, sc + SC_OBJ);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_f, md.d_cvec, (size_t)m.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (md.d_Hx) {
        if ((rc = mipm_hess_spmv(hh, 1.0, md.d_Hx, m.d_x, 0.0, md.d_buffer_n)) != MIPM_OK) return rc;
        k_dot<<<red_grid(h, std::max<int64_t>(md.nx, 1)), TB, 0, h->stream>>>(md.nx, md.d_buffer_n, m.d_x, h->d_partials.p, h->d_counter.p, sc + SC_OBJ + 1);
        MIPM_CHECK_LAUNCH(h);
        if ((rc = mipm_axpby(hh, md.nx, 1.0, md.d_buffer_n, 1.0, m.d_f)) != MIPM_OK) return rc;
    }
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_c, m.d_rhs, (size_t)m.m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return mipm_spmv_pair(hh, md.d_ATx, 1.0, m.d_x, -1.0, m.d_c, 1.0, m.d_y, 0.0, m.d_jacl);
}

/* ------------------------------------------------------------------ batches (stacked independent units) ---- */
static UB make_ub(Handle *h)
{
    UB ub;
    ub.off_n = h->d_uoff_n.p; ub.off_m = h->d_uoff_m.p; ub.active = h->d_uactive.p;
    ub.sc = h->d_usc.p; ub.uin = h->d_uin.p; ub.uout = h->d_uout.p;
    return ub;
}

int mipm_batch_configure(mipm_handle hh, int64_t n_units, const int64_t *off_n, const int64_t *off_m)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n_units < 1 || n_units > (1 << 20) || !off_n || !off_m) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (!h->bound) return fail(h, MIPM_ERR_STATE, "mipm_mpc_bind (stacked vectors) has not been called");
    if (off_n[0] != 0 || off_m[0] != 0 || off_n[n_units] != h->v.n || off_m[n_units] != h->v.m)
        return fail(h, MIPM_ERR_ARG, "unit offsets must start at 0 and end at the stacked sizes");
    for (int64_t u = 0; u < n_units; ++u)
        if (off_n[u + 1] < off_n[u] || off_m[u + 1] < off_m[u]) return fail(h, MIPM_ERR_ARG, "unit offsets must be monotone");
    std::vector<int64_t> a(off_n, off_n + n_units + 1), b(off_m, off_m + n_units + 1);
    std::vector<int> act((size_t)n_units, 1);
    MIPM_CUDA(h, h->d_uoff_n.upload(a, h->stream));
    MIPM_CUDA(h, h->d_uoff_m.upload(b, h->stream));
    MIPM_CUDA(h, h->d_uactive.upload(act, h->stream));
    MIPM_CUDA(h, h->d_usc.alloc((size_t)n_units * SC_COUNT));
    MIPM_CUDA(h, h->d_uin.alloc((size_t)n_units * 4));
    MIPM_CUDA(h, h->d_uout.alloc((size_t)n_units * 8));
    MIPM_CUDA(h, cudaMemsetAsync(h->d_usc.p, 0, (size_t)n_units * SC_COUNT * sizeof(double), h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->h_ubuf.assign((size_t)n_units * SC_COUNT, 0.0);
    h->nb_units = (int)n_units;
    return MIPM_OK;
}

#define BATCH_OR_FAIL(h)                                                                                    \
    VIEW_OR_FAIL(h);                                                                                        \
    if ((h)->nb_units < 1) return fail((h), MIPM_ERR_STATE, "mipm_batch_configure has not been called");    \
    const UB ub = make_ub(h);                                                                               \
    const unsigned B = (unsigned)(h)->nb_units

int mipm_batch_set_active(mipm_handle hh, const int32_t *active)
{
    Handle *h = (Handle *)hh;
    BATCH_OR_FAIL(h);
    (void)ub;
    if (!active) return fail(h, MIPM_ERR_ARG, "null argument");
    MIPM_CUDA(h, cudaMemcpyAsync(h->d_uactive.p, active, B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));      // the host array may be reused right away
    return MIPM_OK;
}

/* per-unit max |x_i| (by_rows = 0: d_x is a stacked vector over the variables; 1: over the constraint rows) */
int mipm_batch_amax(mipm_handle hh, int by_rows, const double *d_x, double *out)
{
    Handle *h = (Handle *)hh;
    BATCH_OR_FAIL(h);
    if (!d_x || !out) return fail(h, MIPM_ERR_ARG, "null argument");
    kb_amax<<<B, TB, 0, h->stream>>>(d_x, by_rows ? ub.off_m : ub.off_n, ub.uout, 8);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_ubuf.data(), ub.uout, (size_t)B * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (unsigned u = 0; u < B; ++u) out[u] = h->h_ubuf[(size_t)u * 8];
    return MIPM_OK;
}

/* per-unit dot products of two stacked variable vectors (objective terms) */
int mipm_batch_dot(mipm_handle hh, const double *d_x, const double *d_y, double *out)
{
    Handle *h = (Handle *)hh;
    BATCH_OR_FAIL(h);
    if (!d_x || !d_y || !out) return fail(h, MIPM_ERR_ARG, "null argument");
    kb_dot<<<B, TB, 0, h->stream>>>(d_x, d_y, ub.off_n, ub.uout, 8);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_ubuf.data(), ub.uout, (size_t)B * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (unsigned u = 0; u < B; ++u) out[u] = h->h_ubuf[(size_t)u * 8];
    return MIPM_OK;
}

/* mipm_init_point_stage for every unit: a[B], b[B] are the per-unit shifts, out is B x 5 */
int mipm_batch_init_point_stage(mipm_handle hh, int stage, const double *a, const double *b, double kappa, double *out)
{
    Handle *h = (Handle *)hh;
    BATCH_OR_FAIL(h);
    if (!out || (stage > 0 && (!a || !b))) return fail(h, MIPM_ERR_ARG, "null argument");
    if (stage > 0) {
        for (unsigned u = 0; u < B; ++u) { h->h_ubuf[(size_t)u * 4] = a[u]; h->h_ubuf[(size_t)u * 4 + 1] = b[u]; }
        MIPM_CUDA(h, cudaMemcpyAsync(ub.uin, h->h_ubuf.data(), (size_t)B * 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (stage == 0) kb_init0<<<B, TB, 0, h->stream>>>(v, ub);
    else if (stage == 1) kb_init1<<<B, TB, 0, h->stream>>>(v, ub);
    else if (stage == 2) kb_init2<<<B, TB, 0, h->stream>>>(v, ub, kappa);
    else return fail(h, MIPM_ERR_ARG, "stage must be 0, 1 or 2");
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_ubuf.data(), ub.uout, (size_t)B * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    for (unsigned u = 0; u < B; ++u)
        for (int q = 0; q < 5; ++q) out[(size_t)u * 5 + q] = h->h_ubuf[(size_t)u * 8 + q];
    return MIPM_OK;
}

/* mipm_mpc_iter_begin for the whole batch: out is B x 16 (same entries per unit), *status as mipm_ls_factorize for the
 * stacked matrix (one unit's breakdown makes every unit retry with more regularization). */
// factorize != 0: mipm_batch_iter_begin; 0: mipm_batch_peek (termination measures only)
static int batch_begin(mipm_handle hh, double del_w, double del_c, double *out, int *status, int factorize)
{
    Handle *h = (Handle *)hh;
    BATCH_OR_FAIL(h);
    int rc = fused_ready(h);
    if (rc != MIPM_OK) return rc;
    if (!out || (factorize && !status)) return fail(h, MIPM_ERR_ARG, "null argument");
    kb_termination<<<B, TB, 0, h->stream>>>(v, ub);
    MIPM_CHECK_LAUNCH(h);
    if (factorize && (rc = fused_factorize(h, del_w, del_c)) != MIPM_OK) return rc;
    int info[4];
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_ubuf.data(), ub.sc, (size_t)B * SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaMemcpyAsync(info, h->d_info.p, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (status) *status = (info[0] == 0) ? MIPM_OK : MIPM_ERR_NOT_FACTORIZED;
    for (unsigned u = 0; u < B; ++u) {
        const double *sc = h->h_ubuf.data() + (size_t)u * SC_COUNT;
        double *o = out + (size_t)u * 16;
        o[0] = (sc[SC_TERM + 0] + sc[SC_TERM + 1]) - sc[SC_TERM + 2];
        o[1] = sc[SC_TERM + 3]; o[2] = sc[SC_TERM + 4]; o[3] = sc[SC_TERM + 5]; o[4] = sc[SC_TERM + 6];
        o[5] = sc[SC_OBJ]; o[6] = sc[SC_OBJ + 1];
        o[7] = sc[SC_ALPHA_P]; o[8] = sc[SC_ALPHA_D]; o[9] = sc[SC_MU]; o[10] = sc[SC_MU_CURR];
        o[11] = o[12] = o[13] = o[14] = 0.0;        // (no residual check in the batched sequence)
        o[15] = sc[SC_TAU];
    }
    return MIPM_OK;
}

int mipm_batch_iter_begin(mipm_handle hh, double del_w, double del_c, double *out, int *status)
{
    return batch_begin(hh, del_w, del_c, out, status, 1);
}

int mipm_batch_peek(mipm_handle hh, double *out) { return batch_begin(hh, 0.0, 0.0, out, nullptr, 0); }

/* mipm_mpc_iter_rest for the whole batch (step_rule 0 = AdaptiveStep, 1 = ConservativeStep). */
int mipm_batch_iter_rest(mipm_handle hh, double mu_min, int step_rule, double tau_param, int ir_steps)
{
    Handle *h = (Handle *)hh;
    BATCH_OR_FAIL(h);
    int rc = fused_ready(h);
    if (rc != MIPM_OK) return rc;
    if (step_rule != 0 && step_rule != 1) return fail(h, MIPM_ERR_ARG, "step_rule must be 0 (adaptive) or 1 (conservative)");
    const mipm_mpc_model &md = h->model;
    const mipm_mpc_vectors &m = h->v;
    const unsigned g = red_grid(h, nmax);
    const int64_t N = m.n + m.m + m.nlb + m.nub;
    // prediction_step!
    k_set_rhs<<<g, TB, 0, h->stream>>>(v, 0, 0.0, nullptr);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_d, m.d_p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = fused_kkt_solve(h, m.d_d, ir_steps)) != MIPM_OK) return rc;
    kb_alpha_max<<<B, TB, 0, h->stream>>>(v, ub, 0, 1.0);
    MIPM_CHECK_LAUNCH(h);
    kb_predictor_measures<<<B, TB, 0, h->stream>>>(v, ub, mu_min);
    MIPM_CHECK_LAUNCH(h);
    // mehrotra_correction_direction!
    kb_set_rhs_corr<<<B, TB, 0, h->stream>>>(v, ub);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_d, m.d_p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = fused_kkt_solve(h, m.d_d, ir_steps)) != MIPM_OK) return rc;
    // update_step_size!, apply_step!
    kb_alpha_max<<<B, TB, 0, h->stream>>>(v, ub, step_rule == 0 ? 1 : 0, tau_param);
    MIPM_CHECK_LAUNCH(h);
    kb_apply_step<<<B, TB, 0, h->stream>>>(v, ub, pow(DBL_EPSILON, 0.75));
    MIPM_CHECK_LAUNCH(h);
    // evaluate_model!
    kb_dot<<<B, TB, 0, h->stream>>>(md.d_cvec, m.d_x, ub.off_n, ub.sc + SC_OBJ, SC_COUNT);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_f, md.d_cvec, (size_t)m.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (md.d_Hx) {
        if ((rc = mipm_hess_spmv(hh, 1.0, md.d_Hx, m.d_x, 0.0, md.d_buffer_n)) != MIPM_OK) return rc;
        kb_dot<<<B, TB, 0, h->stream>>>(md.d_buffer_n, m.d_x, ub.off_n, ub.sc + SC_OBJ + 1, SC_COUNT);
        MIPM_CHECK_LAUNCH(h);
        if ((rc = mipm_axpby(hh, md.nx, 1.0, md.d_buffer_n, 1.0, m.d_f)) != MIPM_OK) return rc;
    }
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_c, m.d_rhs, (size_t)m.m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return mipm_spmv_pair(hh, md.d_ATx, 1.0, m.d_x, -1.0, m.d_c, 1.0, m.d_y, 0.0, m.d_jacl);
}

/* ------------------------------------------------------------------ preprocessing ---- */
int mipm_ruiz_equilibrate(mipm_handle hh, int64_t m, int64_t n, int64_t nnz, const int32_t *d_rows, const int32_t *d_cols,
                          const double *d_vals, int index_base, int max_iter, double tol, double *d_Dr, double *d_Dc, int *iters)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (m < 0 || n < 0 || nnz < 0 || max_iter < 0 || (nnz > 0 && (!d_rows || !d_cols || !d_vals)) || (m > 0 && !d_Dr) || (n > 0 && !d_Dc))
        return fail(h, MIPM_ERR_ARG, "bad argument");
    DBuf<unsigned long long> d_max;
    MIPM_CUDA(h, d_max.alloc((size_t)std::max<int64_t>(m + n, 1)));
    MIPM_CUDA(h, cudaMemsetAsync(d_max.p, 0, (size_t)std::max<int64_t>(m + n, 1) * sizeof(unsigned long long), h->stream));
    if (m > 0) { k_fill<<<red_grid(h, m), TB, 0, h->stream>>>(m, 1.0, d_Dr); MIPM_CHECK_LAUNCH(h); }
    if (n > 0) { k_fill<<<red_grid(h, n), TB, 0, h->stream>>>(n, 1.0, d_Dc); MIPM_CHECK_LAUNCH(h); }
    int it = 0;
    for (; it < max_iter; ++it) {
        if (nnz > 0) {
            k_ruiz_max<<<red_grid(h, nnz), TB, 0, h->stream>>>(nnz, d_rows, d_cols, d_vals, index_base, d_Dr, d_Dc, d_max.p, d_max.p + m);
            MIPM_CHECK_LAUNCH(h);
        }
        k_ruiz_apply<<<red_grid(h, std::max<int64_t>(std::max(m, n), 1)), TB, 0, h->stream>>>(m, n, d_Dr, d_Dc, d_max.p, d_max.p + m,
                                                                                           h->d_partials.p, h->d_counter.p, h->d_scal.p);
        MIPM_CHECK_LAUNCH(h);
        if (tol > 0.0) {            // the convergence test costs a synchronisation per sweep: only when asked for
            double dev;
            int rc = fetch_scalars(h, 1, &dev);
            if (rc != MIPM_OK) return rc;
            if (dev <= tol) { ++it; break; }
        }
    }
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));      // d_max is released on return
    if (iters) *iters = it;
    return MIPM_OK;
}

int mipm_scale_coo(mipm_handle hh, int64_t nnz, const int32_t *d_rows, const int32_t *d_cols, const double *d_vals, int index_base,
                   const double *d_Dr, const double *d_Dc, double *d_out)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (nnz < 0 || (nnz > 0 && (!d_rows || !d_cols || !d_vals || !d_Dr || !d_Dc || !d_out))) return fail(h, MIPM_ERR_ARG, "bad argument");
    if (nnz == 0) return MIPM_OK;
    k_scale_coo<<<red_grid(h, nnz), TB, 0, h->stream>>>(nnz, d_rows, d_cols, d_vals, index_base, d_Dr, d_Dc, d_out);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

/* ------------------------------------------------------------------ fused iteration around an EXTERNAL linear solver ---- */
// The same device-resident-scalar iteration, cut at the points where the normal system is factorized / solved, for
// linear solvers that live outside this handle (the distributed block-angular solver: staged factorization and solves
// with NCCL exchanges in between). NormalKKTSystem only. Sequence per iteration:
//   mipm_mpc_ext_begin(del_w, del_c)    termination measures + set_aug_diagonal_reg! + build_kkt!   [caller: factorize aug_nz]
//   mipm_mpc_ext_fetch(out)             the one synchronisation: out[16] as mipm_mpc_iter_begin
//   mipm_mpc_ext_phase(0, ..)           predictive rhs, d = p, reduce + r2 = A Sigma^-1 r1 - r2      [caller: solve buffer_m]
//   mipm_mpc_ext_phase(1, ..)           finish the solve + residual norms, ratio test, centering, correction rhs,
//                                       d = p, reduce + r2                                          [caller: solve buffer_m]
//   mipm_mpc_ext_phase(2, ..)           finish the solve + residual norms, step rule, apply_step!, evaluate_model!
static int ext_ready(Handle *h)
{
    if (!h->bound) return fail(h, MIPM_ERR_STATE, "mipm_mpc_bind has not been called");
    if (!h->has_model || h->model.kkt_kind != 0) return fail(h, MIPM_ERR_STATE, "mipm_mpc_set_model (NormalKKTSystem) has not been called");
    if (!h->has_spmv || !h->has_jac) return fail(h, MIPM_ERR_STATE, "SpMV / normal-equations assembly not set up");
    return MIPM_OK;
}

static int ext_solve_pre(Handle *h)
{
    mipm_handle hh = (mipm_handle)h;
    const mipm_mpc_vectors &m = h->v;
    const mipm_mpc_model &md = h->model;
    const int64_t N = m.n + m.m + m.nlb + m.nub;
    int rc;
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_d, m.d_p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = mipm_normal_solve_stage(hh, 0, m.d_d, md.d_buffer_n, md.d_buffer_m)) != MIPM_OK) return rc;
    return mipm_spmv(hh, 0, 1.0, md.d_ATx, md.d_buffer_n, -1.0, md.d_buffer_m);
}

static int ext_solve_post(Handle *h, int slot)
{
    mipm_handle hh = (mipm_handle)h;
    const mipm_mpc_vectors &m = h->v;
    const mipm_mpc_model &md = h->model;
    const int64_t N = m.n + m.m + m.nlb + m.nub;
    int rc;
    if ((rc = mipm_normal_solve_stage(hh, 1, m.d_d, md.d_buffer_n, md.d_buffer_m)) != MIPM_OK) return rc;
    if ((rc = mipm_spmv(hh, 1, -1.0, md.d_ATx, md.d_buffer_m, 1.0, md.d_buffer_n)) != MIPM_OK) return rc;
    if ((rc = mipm_normal_solve_stage(hh, 2, m.d_d, md.d_buffer_n, md.d_buffer_m)) != MIPM_OK) return rc;
    // residual check of solve_system! (src/linear_solver.jl:29-35)
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_w, m.d_p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = mipm_spmv_pair(hh, md.d_ATx, -1.0, m.d_d, 1.0, m.d_w + m.n, -1.0, m.d_d + m.n, 1.0, m.d_w)) != MIPM_OK) return rc;
    if ((rc = mipm_kktmul(hh, m.d_w, m.d_d, -1.0, 1.0)) != MIPM_OK) return rc;
    k_two_norms<<<red_grid(h, std::max<int64_t>(N, 1)), TB, 0, h->stream>>>(N, m.d_w, m.d_p, h->d_partials.p, h->d_counter.p,
                                                                         h->d_sc.p + SC_RES + 2 * slot);
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_mpc_ext_begin(mipm_handle hh, double del_w, double del_c)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    int rc = ext_ready(h);
    if (rc != MIPM_OK) return rc;
    V v = make_view(h, inv_lb_buf(h).p, inv_ub_buf(h).p);
    const int64_t nmax = std::max<int64_t>(std::max(v.n, v.m), 1);
    k_termination<<<red_grid(h, nmax), TB, 0, h->stream>>>(v, h->d_partials.p, h->d_counter.p, h->d_sc.p + SC_TERM);
    MIPM_CHECK_LAUNCH(h);
    if ((rc = mipm_set_aug_diagonal_reg(hh, del_w, del_c)) != MIPM_OK) return rc;
    return mipm_normal_assemble(hh, h->v.d_pr_diag, h->model.d_aug_nz, h->model.exact_order);
}

int mipm_mpc_ext_fetch(mipm_handle hh, double *out)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    int rc = ext_ready(h);
    if (rc != MIPM_OK) return rc;
    if (!out) return fail(h, MIPM_ERR_ARG, "null argument");
    MIPM_CUDA(h, cudaMemcpyAsync(h->h_scal, h->d_sc.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    const double *sc = h->h_scal;
    const mipm_mpc_vectors &m = h->v;
    double dobj = sc[SC_TERM + 0];
    if (m.nlb > 0) dobj += sc[SC_TERM + 1];
    if (m.nub > 0) dobj -= sc[SC_TERM + 2];
    out[0] = dobj;
    out[1] = sc[SC_TERM + 3]; out[2] = sc[SC_TERM + 4]; out[3] = sc[SC_TERM + 5]; out[4] = sc[SC_TERM + 6];
    out[5] = sc[SC_OBJ]; out[6] = sc[SC_OBJ + 1];
    out[7] = sc[SC_ALPHA_P]; out[8] = sc[SC_ALPHA_D]; out[9] = sc[SC_MU]; out[10] = sc[SC_MU_CURR];
    out[11] = sc[SC_RES + 0]; out[12] = sc[SC_RES + 1]; out[13] = sc[SC_RES + 2]; out[14] = sc[SC_RES + 3];
    out[15] = sc[SC_TAU];
    return MIPM_OK;
}

int mipm_mpc_ext_phase(mipm_handle hh, int phase, double mu_min, int step_rule, double tau_param)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    int rc = ext_ready(h);
    if (rc != MIPM_OK) return rc;
    if (phase < 0 || phase > 2 || step_rule < 0 || step_rule > 2) return fail(h, MIPM_ERR_ARG, "bad phase / step rule");
    V v = make_view(h, inv_lb_buf(h).p, inv_ub_buf(h).p);
    const mipm_mpc_model &md = h->model;
    const mipm_mpc_vectors &m = h->v;
    const int64_t nmax = std::max<int64_t>(std::max(v.n, v.m), 1);
    const unsigned g = red_grid(h, nmax);
    double *sc = h->d_sc.p;
    if (phase == 0) {
        k_set_rhs<<<g, TB, 0, h->stream>>>(v, 0, 0.0, nullptr);
        MIPM_CHECK_LAUNCH(h);
        return ext_solve_pre(h);
    }
    if (phase == 1) {
        if ((rc = ext_solve_post(h, 0)) != MIPM_OK) return rc;
        k_alpha_max<<<g, TB, 0, h->stream>>>(v, 1.0, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, nullptr, 0.0, sc);
        MIPM_CHECK_LAUNCH(h);
        k_predictor_measures<<<g, TB, 0, h->stream>>>(v, mu_min, h->d_partials.p, h->d_counter.p, sc);
        MIPM_CHECK_LAUNCH(h);
        k_set_rhs<<<g, TB, 0, h->stream>>>(v, 1, 0.0, sc + SC_MU);
        MIPM_CHECK_LAUNCH(h);
        return ext_solve_pre(h);
    }
    if ((rc = ext_solve_post(h, 1)) != MIPM_OK) return rc;
    if (step_rule == 2) {
        if ((rc = mehrotra_step_launch(h, v, g, tau_param, sc)) != MIPM_OK) return rc;
    } else {
        if (step_rule == 0) k_alpha_max<<<g, TB, 0, h->stream>>>(v, 0.0, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, sc, tau_param, sc);
        else k_alpha_max<<<g, TB, 0, h->stream>>>(v, tau_param, h->d_partials.p, h->d_counter.p, sc + SC_ALPHA, nullptr, 0.0, sc);
        MIPM_CHECK_LAUNCH(h);
    }
    k_apply_step<<<g, TB, 0, h->stream>>>(v, 0.0, 0.0, 0.0, pow(DBL_EPSILON, 0.75), sc);
    MIPM_CHECK_LAUNCH(h);
    k_dot<<<red_grid(h, std::max<int64_t>(md.nx, 1)), TB, 0, h->stream>>>(md.nx, md.d_cvec, m.d_x, h->d_partials.p, h->d_counter.p, sc + SC_OBJ);
    MIPM_CHECK_LAUNCH(h);
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_f, md.d_cvec, (size_t)m.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    MIPM_CUDA(h, cudaMemcpyAsync(m.d_c, m.d_rhs, (size_t)m.m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return mipm_spmv_pair(hh, md.d_ATx, 1.0, m.d_x, -1.0, m.d_c, 1.0, m.d_y, 0.0, m.d_jacl);
}

}  // extern "C"
