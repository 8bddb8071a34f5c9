// Shared declarations for the madipm_b200 CUDA library (handle, error plumbing, device buffers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/madipm_b200.h"
#include "host_symbolic.h"
#include "ls_symbolic.h"

namespace mipm {

// Stream of the API call the calling thread is in (set by MIPM_NEED_DEVICE / use_handle). Device buffers come from
// CUDA's stream-ordered memory pool on that stream: cudaMalloc / cudaFree synchronise the whole device, which hurts
// exactly when many small handles are created and destroyed side by side (batches of independent LPs, config C5).
inline thread_local cudaStream_t tl_stream = nullptr;
inline thread_local bool tl_pooled = false;

// Simple owning device buffer.
template <typename T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t st = nullptr;       // stream the pooled allocation is ordered on
    bool pooled = false;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) {
            if (pooled) cudaFreeAsync(p, st);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        if (tl_pooled) {
            pooled = true;
            st = tl_stream;
            return cudaMallocAsync((void **)&p, count * sizeof(T), st);
        }
        pooled = false;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    template <typename A>
    cudaError_t upload(const std::vector<T, A> &h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

struct Handle {
    int device = -1;
    bool host_only = false;
    bool pool_ok = false;            // device supports stream-ordered allocation (cudaMallocAsync)
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // reduction scratch
    DBuf<double> d_partials;
    DBuf<double> d_scal;
    DBuf<unsigned int> d_counter;
    double *h_scal = nullptr;   // pinned
    int red_blocks = 0;

    // ---- normal equations
    NormalSymbolic nsym;
    bool has_normal = false, has_jac = false;
    DBuf<int32_t> d_term_ptr, d_term_pi, d_term_pj, d_term_k, d_asm_blk;
    int64_t asm_nblk = 0;            // runs of stored entries for the streaming assembly
    DBuf<double> d_term_w, d_D;
    const double *d_ATx = nullptr;

    // ---- K2
    K2Symbolic ksym;
    bool has_k2 = false;
    DBuf<int64_t> d_slot_ptr, d_slot_src;

    // ---- linear solver
    LsSymbolic sym;
    bool has_ls = false, factorized = false;
    bool ldl_definite = false;   // MIPM_LDL_DEFINITE: LDL^T kernels on a definite matrix (normal equations)
    int n_tasks = 0, grid_factor = 0, grid_solve = 0;
    int grid_limit = 0;              // cap on the persistent kernels' grid (0 = whole GPU)
    int root_task_begin = -1;        // first task of the border (root) front's own factorization, -1 = no border
    int64_t leaf_off = 0;            // small leaf fronts (one warp each) in d_sched
    int n_leaf = 0;
    cudaStream_t side = nullptr;     // zero-fill of the update matrices for the next factorization
    cudaEvent_t ev_factor_done = nullptr, ev_u_zero = nullptr;
    bool u_prezeroed = false;
    DBuf<int32_t> d_sched;           // schedule arrays (per-level front lists, small-leaf list, extend-add child ranges)
    DBuf<int64_t> d_lvl;
    struct FI64 { char b[64]; };
    DBuf<FI64> d_finfo;              // FrontInfo records (64 bytes each, see front.cuh)
    struct T32 { char b[32]; };
    DBuf<T32> d_tasks;               // task list of the factorization (32 bytes each, see factor.cu)
    DBuf<int> d_prog;                // per front: completed tasks; [ns] = ticket counter
    DBuf<unsigned long long> d_prof, d_front_ns, d_trace;     // per-CTA busy time per task class; per-front completion time (diagnostic)
    DBuf<double> d_Dinv;             // inverted 64 x 64 diagonal blocks
    DBuf<double> d_Dg;               // LDL^T pivots by (permuted) column
    DBuf<int32_t> d_row_idx, d_rel_idx, d_perm, d_child_idx, d_full_col;
    DBuf<int64_t> d_a2l, d_full_ptr, d_full_val;
    DBuf<double> d_L, d_U, d_xp, d_uvec, d_b, d_r;
    DBuf<double> d_L2;               // second factor buffer: the idle one is zero-filled on the side stream (not in border mode)
    double *L_cur = nullptr;         // buffer holding the current factor
    bool l_prezeroed = false;
    DBuf<int64_t> d_gat_off;         // transposed child maps of the forward solve (fronts with many children)
    DBuf<int32_t> d_gat_ptr, d_gat_src;
    struct I4 { int x, y, z, w; };
    DBuf<I4> d_solve_tasks;          // task list of the solves (kind, id, part, 0), see solve.cu
    DBuf<int> d_solve_prog;          // per front: forward children done | backward done | big-front counters x3; [5 ns] = ticket
    int n_solve_fwd = 0, n_solve_tasks = 0, solve_root_fwd_begin = 0;
    DBuf<int> d_info;                // [0]=first failed column+1 (0 = ok), [1]=#neg pivots, [2]=#zero pivots
    const double *d_nzval = nullptr;
    int64_t n_launch_factor = 0;

    // ---- spmv
    bool has_spmv = false;
    int64_t sp_m = 0, sp_n = 0, sp_nnz = 0;
    DBuf<int32_t> d_sp_rowptr, d_sp_col, d_sp_colptr, d_sp_row, d_sp_pos;
    DBuf<int32_t> d_sp_blk_rows, d_sp_blk_cols;     // row / column runs of the streaming SpMV
    int64_t sp_nblk_rows = 0, sp_nblk_cols = 0;
    DBuf<double> d_sp_valT;                         // column-ordered copy of the values (mipm_spmv_cache_values)
    const double *sp_valT_src = nullptr;            // the caller's value array the copy was taken from

    // ---- Hessian operator (full symmetric CSR), the MadIPMOperator(H; symmetric=true) analogue
    bool has_hess = false;
    int64_t hs_n = 0, hs_nnz = 0;
    DBuf<int32_t> d_hs_rowptr, d_hs_col;

    // ---- fused MPC iteration: device-resident scalar block and the model / KKT buffers it drives
    DBuf<double> d_sc;
    bool has_model = false;
    mipm_mpc_model model{};

    // ---- batch of independent problems stacked into one (BASELINE config C5): per-unit offsets and scalar blocks
    int nb_units = 0;
    DBuf<int64_t> d_uoff_n, d_uoff_m;
    DBuf<int> d_uactive;
    DBuf<double> d_usc, d_uin, d_uout;
    std::vector<double> h_ubuf;

    // ---- mpc vectors (+ inverse maps variable -> position in the lb / ub block, -1 if none)
    DBuf<int32_t> d_inv_lb, d_inv_ub;
    bool bound = false;
    mipm_mpc_vectors v{};
};

inline int fail(Handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

#define MIPM_CUDA(h, call)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return mipm::fail((h), MIPM_ERR_CUDA,                                           \
                              std::string(#call) + ": " + cudaGetErrorString(e__));        \
    } while (0)

#define MIPM_CHECK_LAUNCH(h)                                                                \
    do {                                                                                    \
        (h)->launches++;                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess)                                                             \
            return mipm::fail((h), MIPM_ERR_CUDA, std::string("kernel launch: ") +         \
                                                      cudaGetErrorString(e__));            \
    } while (0)

#define MIPM_NEED_DEVICE(h)                                                                 \
    do {                                                                                    \
        if (!(h)) return MIPM_ERR_ARG;                                                      \
        if ((h)->host_only)                                                                 \
            return mipm::fail((h), MIPM_ERR_CUDA,                                           \
                              "analysis-only handle (device < 0): no device work possible; " \
                              "there is no CPU fallback");                                  \
        mipm::use_handle(h);                                                                \
    } while (0)

// Per-process caches of things that are slow or device-synchronising to obtain per handle (they matter when hundreds of
// small handles are created: BASELINE config C5). Implemented in api.cu.
struct DeviceInfo { int sm_count = 0; int cooperative = 0; };
int device_info(int device, DeviceInfo &out);       // cudaGetDeviceProperties once per device
double *pinned_scalars_acquire();                   // 64 pinned doubles from a free list (never returned to the driver)
void pinned_scalars_release(double *p);

inline void use_handle(const Handle *h)
{
    if (!h->host_only) cudaSetDevice(h->device);     // cheap when unchanged; handles on several GPUs in one process
    tl_stream = h->stream;
    tl_pooled = h->pool_ok && !h->host_only;
}

// implemented in the .cu files
int ls_device_setup(Handle *h);
int normal_symbolic_device(Handle *h, int64_t m, int64_t n, const int32_t *Ap, const int32_t *Aj, int index_base,
                           int32_t **Cp_out, int32_t **Cj_out);
int ls_solve_setup(Handle *h, const void *finfo_host, const char *small_leaf);
int ls_factorize_impl(Handle *h, const double *d_nzval);
int ls_solve_impl(Handle *h, double *d_x, int ir_steps);
int ls_factorize_staged(Handle *h, const double *d_nzval, int stage);
int ls_solve_staged(Handle *h, double *d_x, int stage);

}  // namespace mipm
