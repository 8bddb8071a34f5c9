// Regression helper for the host analysis: reads a lower CSC pattern (binary: int64 n, int64 nnz, int32 kind, int32 ordering,
// int64 n_border, colptr[n+1], rowval[nnz]), runs ls_analyze and prints an FNV hash of every array of the result, so a
// refactoring of ls_symbolic.cpp can be checked to be output-identical.  Build: see tools/analyze_hash.sh
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../madipm_jl_b200/csrc/ls_symbolic.h"
using namespace mipm;
template <typename T, typename A> static unsigned long long fnv(const std::vector<T, A> &v)
{
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *p = (const unsigned char *)v.data();
    for (size_t i = 0; i < v.size() * sizeof(T); ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t n, nnz, nb; int32_t kind, ordering;
    if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1 || fread(&kind, 4, 1, f) != 1 || fread(&ordering, 4, 1, f) != 1 || fread(&nb, 8, 1, f) != 1) return 4;
    std::vector<int32_t> cp((size_t)n + 1), ri((size_t)nnz);
    if (fread(cp.data(), 4, cp.size(), f) != cp.size() || fread(ri.data(), 4, ri.size(), f) != ri.size()) return 5;
    std::fclose(f);
    LsOptions opt; opt.kind = kind; opt.ordering = ordering; opt.n_border = nb;
    LsSymbolic S;
    auto t0 = std::chrono::steady_clock::now();
    std::string e = ls_analyze(n, cp.data(), ri.data(), 0, opt, nullptr, S);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!e.empty()) { std::printf("error: %s\n", e.c_str()); return 1; }
    std::printf("time %.3f s\n", dt);
    std::printf("perm %016llx sn_ptr %016llx sn_parent %016llx sn_level %016llx row_ptr %016llx row_idx %016llx rel_idx %016llx\n",
                fnv(S.perm), fnv(S.sn_ptr), fnv(S.sn_parent), fnv(S.sn_level), fnv(S.row_ptr), fnv(S.row_idx), fnv(S.rel_idx));
    std::printf("lp %016llx up %016llx child_ptr %016llx child_idx %016llx level_sn %016llx a2l %016llx nnz_l %lld exact %lld flops %.0f ns %d levels %d\n",
                fnv(S.lp), fnv(S.up), fnv(S.child_ptr), fnv(S.child_idx), fnv(S.level_sn), fnv(S.a2l), (long long)S.nnz_l,
                (long long)S.nnz_l_exact, S.flops, S.ns, S.n_levels);
    return 0;
}
