// Host-side symbolic analysis for the supernodal multifrontal Cholesky / LDL^T.
// Replaces the `analysis` phase of cuDSS that MadNLPGPU.CUDSSSolver runs at construction
// (reference call site: src/KKT/normalkkt.jl:113-115). Pure C++ (no CUDA) so it can be
// exercised on a CPU-only box through an analysis-only handle.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace mipm {

// std::vector without value-initialization: large arrays are first touched by the host threads that fill them (a single
// thread faulting in 100 MB of fresh pages costs as much as the pass that uses them).
template <typename T>
struct DefaultInitAlloc : std::allocator<T> {
    template <typename U> struct rebind { using other = DefaultInitAlloc<U>; };
    template <typename U> void construct(U *p) noexcept { ::new ((void *)p) U; }
    template <typename U, typename... Args> void construct(U *p, Args &&...args) { ::new ((void *)p) U(std::forward<Args>(args)...); }
};
template <typename T> using uvector = std::vector<T, DefaultInitAlloc<T>>;

struct LsOptions {
    int kind = 0;            // MIPM_CHOLESKY / MIPM_LDL
    int ordering = 0;        // MIPM_ORDER_*
    // defaults from the C2 sweep (tools/sweep_symbolic.py): larger leaves and more aggressive
    // amalgamation trade ~3% more stored nonzeros for ~25% fewer supernodes
    int nd_leaf = 256;       // stop dissecting below this many vertices
    int relax_always = 16;   // amalgamation: merge if merged width <= this
    int relax_k1 = 64;  double relax_z1 = 0.50;
    int relax_k2 = 128; double relax_z2 = 0.30;
    double relax_z3 = 0.10;
    int max_sn_cols = 1 << 30;
    // LDL^T on K2: eliminate every dual vertex after ALL of its primal neighbours (true) or after the
    // first one (false). 'All' is the numerically safe choice for tiny |delta_c| (quirk A.9 v).
    bool ldl_delay_all = true;
    // Distributed (block-angular) use: the last n_border vertices are kept last, in their given order, and form
    // ONE final dense supernode (the root separator whose Schur block is all-reduced across GPUs).
    int64_t n_border = 0;
    // scatter map of the input nonzeros on the host (analysis-only handles, tools); handles with a GPU build it there
    bool host_a2l = true;
};

struct LsSymbolic {
    int64_t n = 0, nnz_a = 0;
    int kind = 0;
    std::vector<int32_t> perm, iperm;          // perm[new] = old, iperm[old] = new
    int32_t ns = 0;
    std::vector<int32_t> sn_ptr;               // ns+1, first column of each supernode
    std::vector<int32_t> sn_parent;            // ns, -1 for roots
    std::vector<int32_t> sn_level;             // ns
    std::vector<int32_t> col2sn;               // n
    std::vector<int64_t> row_ptr;              // ns+1
    std::vector<int32_t> row_idx;              // below-diagonal rows (permuted numbering)
    std::vector<int32_t> rel_idx;              // same shape as row_idx: position in parent's front
    std::vector<int64_t> lp;                   // ns+1, panel offsets in L storage (doubles)
    std::vector<int64_t> up;                   // ns+1, update-matrix offsets
    std::vector<int64_t> child_ptr;            // ns+1
    std::vector<int32_t> child_idx;            // children in ascending order
    std::vector<int64_t> level_ptr;            // nlev+1
    std::vector<int32_t> level_sn;             // supernodes grouped by level
    uvector<int64_t> a2l;                      // nnz_a: destination of each input nonzero in L storage (empty: built on the device)
    uvector<int32_t> in_colptr, in_rowval;     // copy of the analysed pattern (0-based)
    // full symmetric CSR of the input matrix in ORIGINAL numbering (refinement residual): filled by ls_build_full_csr
    std::vector<int64_t> full_ptr;             // n+1
    std::vector<int32_t> full_col;
    std::vector<int64_t> full_val;             // index into the caller's nzval
    // stats
    int64_t nnz_l = 0, nnz_l_exact = 0, update_doubles = 0;
    double flops = 0.0;
    int32_t n_levels = 0, max_front_cols = 0, max_front_rows = 0;
    int32_t root_sn = -1;                      // the border supernode when n_border > 0
};

// colptr/rowval: lower-triangular CSC in `index_base`; user_perm 0-based. Returns "" on success or an error message.
std::string ls_analyze(int64_t n, const int32_t *colptr, const int32_t *rowval, int index_base,
                       const LsOptions &opt, const int32_t *user_perm, LsSymbolic &out);
void ls_build_a2l(LsSymbolic &S, std::string &err);

void ls_build_full_csr(LsSymbolic &S);

// Nested-dissection ordering (level-structure separators, George & Liu) of the graph of a
// symmetric matrix given by its full adjacency (no self loops). perm[new] = old.
void order_nested_dissection(int64_t n, const int64_t *xadj, const int32_t *adj, int leaf_size,
                             std::vector<int32_t> &perm);

}  // namespace mipm
