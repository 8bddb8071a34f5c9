// Host-side symbolic builders. See host_symbolic.h. The outputs are canonical (sorted lower
// CSC), so they are bit-exact with the reference's while using different algorithms: the
// reference's build_normal_system scans every row j >= i for every i (O(m^2) row scans,
// src/utils.jl:220-235); here row i's pattern is the sorted union, over the columns k of
// row i, of the rows j >= i of column k, and the same sweep emits the product-term map.
#include "host_symbolic.h"

#include <cstdlib>
#include <functional>
#include <thread>

#include <algorithm>
#include <cstring>

namespace mipm {

// Host threads for the one-time symbolic work (MIPM_HOST_THREADS overrides; results never depend on the count).
int host_threads()
{
    if (const char *e = std::getenv("MIPM_HOST_THREADS")) return std::max(1, atoi(e));
    unsigned hc = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(hc ? hc : 1u, 16u);
}

void run_host_threads(int nthreads, const std::function<void(int)> &work)
{
    if (nthreads <= 1) { work(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads - 1);
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
}


void coo_to_csr_host(int64_t n_rows, int64_t nnz, const int32_t *Ai, const int32_t *Aj,
                     int32_t *Bp, int32_t *Bj, int64_t *Bmap)
{
    std::vector<int64_t> pos((size_t)n_rows + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) pos[(size_t)Ai[k] + 1]++;
    for (int64_t i = 0; i < n_rows; ++i) pos[(size_t)i + 1] += pos[(size_t)i];
    for (int64_t i = 0; i <= n_rows; ++i) Bp[i] = (int32_t)pos[(size_t)i];
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t d = pos[(size_t)Ai[k]]++;
        Bj[d] = Aj[k];
        Bmap[d] = k;
    }
}

std::string normal_symbolic_host(int64_t m, int64_t n, const int32_t *Ap_in, const int32_t *Aj_in,
                                 int index_base, NormalSymbolic &S)
{
    S = NormalSymbolic();
    if (m < 0 || n < 0) return "negative dimension";
    S.m = m;
    S.n = n;
    S.Ap.resize((size_t)m + 1);
    for (int64_t i = 0; i <= m; ++i) S.Ap[(size_t)i] = Ap_in[i] - index_base;
    if (m > 0 && S.Ap[0] != 0) return "row pointer does not start at index_base";
    const int64_t nnz = m > 0 ? S.Ap[(size_t)m] : 0;
    S.nnz_a = nnz;
    S.Aj.resize((size_t)nnz);
    for (int64_t p = 0; p < nnz; ++p) {
        int32_t k = Aj_in[p] - index_base;
        if (k < 0 || k >= n) return "column index out of range";
        S.Aj[(size_t)p] = k;
    }
    // CSC of A with CSR positions; rows ascend inside a column because rows are swept in order
    std::vector<int64_t> cptr((size_t)n + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) cptr[(size_t)S.Aj[(size_t)p] + 1]++;
    for (int64_t k = 0; k < n; ++k) cptr[(size_t)k + 1] += cptr[(size_t)k];
    std::vector<int32_t> crow((size_t)nnz), cpos((size_t)nnz);
    {
        std::vector<int64_t> pos(cptr.begin(), cptr.end() - 1);
        for (int64_t i = 0; i < m; ++i) {
            if (S.Ap[(size_t)i + 1] < S.Ap[(size_t)i]) return "row pointer not monotone";
            for (int32_t p = S.Ap[(size_t)i]; p < S.Ap[(size_t)i + 1]; ++p) {
                int64_t d = pos[(size_t)S.Aj[(size_t)p]]++;
                crow[(size_t)d] = (int32_t)i;
                cpos[(size_t)d] = p;
            }
        }
    }
    // duplicates inside a row break the reference's buffer[k] overwrite semantics
    for (int64_t k = 0; k < n; ++k)
        for (int64_t d = cptr[(size_t)k] + 1; d < cptr[(size_t)k + 1]; ++d)
            if (crow[(size_t)d] == crow[(size_t)d - 1]) return "duplicate column inside a row of A";

    // Row i of the pattern = sorted union over the columns k of row i of {j >= i : A[j,k] != 0}; the same sweep emits the
    // product terms (p_i, p_j, k) grouped by destination (i, j) and ordered by p_j inside a group. Rows are independent:
    // contiguous row chunks go to host threads, each fills its own vectors, and the pieces are concatenated in row
    // order, so the result does not depend on the thread count.
    struct Term { uint64_t key; int32_t pi; };
    struct Piece {
        std::vector<int32_t> Cj, term_ptr, term_pi, term_pj, term_k, row_nnz;
        std::string err;
    };
    int nthreads = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, m / 4096));
    std::vector<Piece> pieces((size_t)nthreads);
    auto work = [&](int t) {
        Piece &P = pieces[(size_t)t];
        const int64_t r0 = m * t / nthreads, r1 = m * (t + 1) / nthreads;
        std::vector<Term> buf;
        P.row_nnz.reserve((size_t)(r1 - r0));
        for (int64_t i = r0; i < r1; ++i) {
            buf.clear();
            for (int32_t p = S.Ap[(size_t)i]; p < S.Ap[(size_t)i + 1]; ++p) {
                const int32_t k = S.Aj[(size_t)p];
                const int32_t *cb = crow.data() + cptr[(size_t)k], *ce = crow.data() + cptr[(size_t)k + 1];
                int64_t d = cptr[(size_t)k] + (std::lower_bound(cb, ce, (int32_t)i) - cb);      // first entry with row >= i
                for (; d < cptr[(size_t)k + 1]; ++d)
                    buf.push_back(Term{((uint64_t)(uint32_t)crow[(size_t)d] << 32) | (uint32_t)cpos[(size_t)d], p});
            }
            std::sort(buf.begin(), buf.end(), [](const Term &a, const Term &b) { return a.key < b.key; });
            int32_t lastj = -1, cnt = 0;
            for (const Term &tm : buf) {
                const int32_t j = (int32_t)(tm.key >> 32);
                const int32_t pj = (int32_t)(tm.key & 0xffffffffu);
                if (j != lastj) {
                    P.term_ptr.push_back((int32_t)P.term_pi.size());      // start of the group, relative to the piece
                    P.Cj.push_back(j);
                    lastj = j;
                    ++cnt;
                }
                P.term_pi.push_back(tm.pi);
                P.term_pj.push_back(pj);
                P.term_k.push_back(S.Aj[(size_t)pj]);
            }
            P.row_nnz.push_back(cnt);
        }
    };
    run_host_threads(nthreads, work);
    int64_t tot_c = 0, tot_t = 0;
    for (const Piece &P : pieces) { tot_c += (int64_t)P.Cj.size(); tot_t += (int64_t)P.term_pi.size(); }
    if (tot_t >= (int64_t)INT32_MAX || tot_c >= (int64_t)INT32_MAX) return "too many product terms for 32-bit segment pointers";
    S.Cp.assign((size_t)m + 1, 0);
    S.Cj.reserve((size_t)tot_c);
    S.term_ptr.reserve((size_t)tot_c + 1);
    S.term_pi.reserve((size_t)tot_t);
    S.term_pj.reserve((size_t)tot_t);
    S.term_k.reserve((size_t)tot_t);
    int64_t row = 0;
    for (const Piece &P : pieces) {
        const int32_t toff = (int32_t)S.term_pi.size();
        for (int32_t v : P.term_ptr) S.term_ptr.push_back(v + toff);
        S.Cj.insert(S.Cj.end(), P.Cj.begin(), P.Cj.end());
        S.term_pi.insert(S.term_pi.end(), P.term_pi.begin(), P.term_pi.end());
        S.term_pj.insert(S.term_pj.end(), P.term_pj.begin(), P.term_pj.end());
        S.term_k.insert(S.term_k.end(), P.term_k.begin(), P.term_k.end());
        for (int32_t c : P.row_nnz) { S.Cp[(size_t)row + 1] = S.Cp[(size_t)row] + c; ++row; }
    }
    S.term_ptr.push_back((int32_t)S.term_pi.size());
    S.nnz_c = (int64_t)S.Cj.size();
    S.n_terms = (int64_t)S.term_pi.size();
    return "";
}

std::string k2_symbolic_host(int64_t dim, int64_t nnz_coo, const int32_t *I, const int32_t *J,
                             int index_base, K2Symbolic &S)
{
    S = K2Symbolic();
    if (dim < 0 || nnz_coo < 0) return "negative size";
    S.dim = dim;
    S.nnz_coo = nnz_coo;
    std::vector<int64_t> cptr((size_t)dim + 1, 0);
    for (int64_t k = 0; k < nnz_coo; ++k) {
        int64_t i = (int64_t)I[k] - index_base, j = (int64_t)J[k] - index_base;
        if (i < 0 || i >= dim || j < 0 || j >= dim) return "COO index out of range";
        if (i < j) return "COO entry above the diagonal (lower triangle expected)";
        cptr[(size_t)j + 1]++;
    }
    for (int64_t j = 0; j < dim; ++j) cptr[(size_t)j + 1] += cptr[(size_t)j];
    // stable bucket by column, then stable sort by row inside each column
    std::vector<int64_t> bucket((size_t)nnz_coo);
    {
        std::vector<int64_t> pos(cptr.begin(), cptr.end() - 1);
        for (int64_t k = 0; k < nnz_coo; ++k) bucket[(size_t)pos[(size_t)(J[k] - index_base)]++] = k;
    }
    S.colptr.assign((size_t)dim + 1, 0);
    S.map.resize((size_t)nnz_coo);
    S.slot_ptr.push_back(0);
    for (int64_t j = 0; j < dim; ++j) {
        auto b = bucket.begin() + cptr[(size_t)j], e = bucket.begin() + cptr[(size_t)j + 1];
        std::stable_sort(b, e, [&](int64_t a, int64_t c) { return I[a] < I[c]; });
        int32_t last = -1;
        for (auto it = b; it != e; ++it) {
            int32_t i = I[*it] - index_base;
            if (i != last) {
                if (last >= 0) S.slot_ptr.push_back((int64_t)S.slot_src.size());
                S.rowval.push_back(i);
                last = i;
            }
            S.map[(size_t)*it] = (int64_t)S.rowval.size() - 1;
            S.slot_src.push_back(*it);
        }
        if (last >= 0) S.slot_ptr.push_back((int64_t)S.slot_src.size());
        S.colptr[(size_t)j + 1] = (int32_t)S.rowval.size();
    }
    S.nnz_csc = (int64_t)S.rowval.size();
    return "";
}

}  // namespace mipm
