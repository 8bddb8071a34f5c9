// Host-side symbolic builders for the two KKT formulations (pure C++, no CUDA).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace mipm {

// Host threads for the one-time symbolic work (env MIPM_HOST_THREADS overrides); work(t) runs for t in [0, nthreads).
int host_threads();
void run_host_threads(int nthreads, const std::function<void(int)> &work);

// Replaces MadIPM.coo_to_csr (src/utils.jl:158-207): stable counting sort by row. 0-based.
void coo_to_csr_host(int64_t n_rows, int64_t nnz, const int32_t *Ai, const int32_t *Aj,
                     int32_t *Bp, int32_t *Bj, int64_t *Bmap);

struct NormalSymbolic {
    int64_t m = 0, n = 0, nnz_a = 0, nnz_c = 0, n_terms = 0;
    std::vector<int32_t> Ap, Aj;          // CSR of A, 0-based
    std::vector<int32_t> Cp, Cj;          // lower CSC of A A' (column i: rows j >= i), 0-based
    std::vector<int32_t> term_ptr;        // nnz_c + 1
    std::vector<int32_t> term_pi, term_pj, term_k;   // per product term: CSR positions and column
};
// Replaces MadIPM.build_normal_system (src/utils.jl:209-274), bit-exact pattern.
std::string normal_symbolic_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Aj,
                                 int index_base, NormalSymbolic &out);

struct K2Symbolic {
    int64_t dim = 0, nnz_coo = 0, nnz_csc = 0;
    std::vector<int32_t> colptr, rowval;  // 0-based lower CSC
    std::vector<int64_t> map;             // coo -> csc slot, 0-based
    std::vector<int64_t> slot_ptr;        // nnz_csc + 1: gather segments
    std::vector<int64_t> slot_src;        // coo indices, ascending inside a slot
};
// Replaces MadNLP's coo_to_csc for SparseKKTSystem (SparseArrays.sparse pattern + map).
std::string k2_symbolic_host(int64_t dim, int64_t nnz_coo, const int32_t *I, const int32_t *J,
                             int index_base, K2Symbolic &out);

}  // namespace mipm
