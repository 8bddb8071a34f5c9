// Host-side symbolic builders. See host_symbolic.h. The outputs are canonical (sorted lower
// CSC), so they are bit-exact with the reference's while using different algorithms: the
// reference's build_normal_system scans every row j >= i for every i (O(m^2) row scans,
// src/utils.jl:220-235); here row i's pattern is the sorted union, over the columns k of
// row i, of the rows j >= i of column k, and the same sweep emits the product-term map.
#include "host_symbolic.h"

#include <algorithm>
#include <cstring>

namespace mipm {

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

    struct Term { uint64_t key; int32_t pi; };
    std::vector<int64_t> start(cptr.begin(), cptr.end() - 1);   // first entry of column k with row >= i
    std::vector<Term> buf;
    S.Cp.assign((size_t)m + 1, 0);
    S.term_ptr.push_back(0);
    for (int64_t i = 0; i < m; ++i) {
        buf.clear();
        for (int32_t p = S.Ap[(size_t)i]; p < S.Ap[(size_t)i + 1]; ++p) {
            int32_t k = S.Aj[(size_t)p];
            int64_t &st = start[(size_t)k];
            while (st < cptr[(size_t)k + 1] && crow[(size_t)st] < i) ++st;
            for (int64_t d = st; d < cptr[(size_t)k + 1]; ++d)
                buf.push_back(Term{((uint64_t)(uint32_t)crow[(size_t)d] << 32) | (uint32_t)cpos[(size_t)d], p});
        }
        std::sort(buf.begin(), buf.end(), [](const Term &a, const Term &b) { return a.key < b.key; });
        int32_t lastj = -1;
        for (const Term &t : buf) {
            int32_t j = (int32_t)(t.key >> 32);
            int32_t pj = (int32_t)(t.key & 0xffffffffu);
            if (j != lastj) {
                if (lastj >= 0) S.term_ptr.push_back((int32_t)S.term_pi.size());
                S.Cj.push_back(j);
                lastj = j;
            }
            S.term_pi.push_back(t.pi);
            S.term_pj.push_back(pj);
            S.term_k.push_back(S.Aj[(size_t)pj]);
            if (S.term_pi.size() >= (size_t)INT32_MAX) return "too many product terms for 32-bit segment pointers";
        }
        if (lastj >= 0) S.term_ptr.push_back((int32_t)S.term_pi.size());
        S.Cp[(size_t)i + 1] = (int32_t)S.Cj.size();
    }
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
