// Host-side symbolic builders for the two KKT formulations (pure C++, no CUDA).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace mipm {

// Host threads for the one-time symbolic work (env MIPM_HOST_THREADS overrides); work(t) runs for t in [0, nthreads).
int host_threads();
void run_host_threads(int nthreads, const std::function<void(int)> &work);

// Stable bucket sort of items by key on the host threads: thread t owns the key range [nkeys t/T, nkeys (t+1)/T), streams
// over all keys twice (count, place) and handles those in its range, so items keep their order inside a bucket and no
// per-thread histogram of all keys is needed. ptr (nkeys + 1) receives the bucket offsets; emit(dst, item) stores an item.
// Returns false if a key lies outside [key_base, key_base + nkeys).
template <typename Emit>
bool stable_bucket_parallel(int64_t nkeys, int64_t nitems, const int32_t *key, int key_base, int32_t *ptr, Emit emit)
{
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(host_threads(), nitems / 262144), nkeys));
    std::vector<int64_t> tcount((size_t)T + 1, 0);
    std::vector<int> bad((size_t)T, 0);
    ptr[0] = 0;
    run_host_threads(T, [&](int t) {
        const int64_t k0 = nkeys * t / T, k1 = nkeys * (t + 1) / T;
        for (int64_t k = k0; k < k1; ++k) ptr[k + 1] = 0;
        int64_t c = 0;
        for (int64_t q = 0; q < nitems; ++q) {
            const int64_t k = (int64_t)key[q] - key_base;
            if (k < 0 || k >= nkeys) { bad[(size_t)t] = 1; return; }
            if (k >= k0 && k < k1) { ptr[k + 1]++; ++c; }
        }
        tcount[(size_t)t + 1] = c;
    });
    for (int v : bad) if (v) return false;
    for (int t = 0; t < T; ++t) tcount[(size_t)t + 1] += tcount[(size_t)t];
    run_host_threads(T, [&](int t) {
        const int64_t k0 = nkeys * t / T, k1 = nkeys * (t + 1) / T;
        std::vector<int32_t> cur((size_t)(k1 - k0));
        int64_t run = tcount[(size_t)t];
        for (int64_t k = k0; k < k1; ++k) {            // counts -> offsets (ptr[k + 1] ends bucket k)
            const int32_t c = ptr[k + 1];
            cur[(size_t)(k - k0)] = (int32_t)run;
            run += c;
            ptr[k + 1] = (int32_t)run;
        }
        for (int64_t q = 0; q < nitems; ++q) {
            const int64_t k = (int64_t)key[q] - key_base;
            if (k >= k0 && k < k1) emit((int64_t)cur[(size_t)(k - k0)]++, q);
        }
    });
    return true;
}

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
