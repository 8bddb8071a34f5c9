// Host symbolic analysis: ordering, elimination tree, supernodes, multifrontal structure.
// See ls_symbolic.h. Written from the textbook algorithms (George & Liu nested dissection,
// Liu's elimination tree / Davis's row-subtree column counts, CHOLMOD-style relaxed
// amalgamation); nothing here comes from the reference, whose analysis lives inside the
// closed cuDSS binary.
#include "ls_symbolic.h"
#include "host_symbolic.h"

#include <atomic>
#include <condition_variable>
#include <mutex>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cstring>
#include <numeric>

namespace mipm {

namespace {

// Adjacency of the permuted graph split at the diagonal: for every new index r, low[r] = {c < r} and / or high[r] = {c > r}
// (unsorted), built vertex by vertex from the full adjacency by the host threads.
void split_permuted_adjacency(int64_t n, int nth, const std::vector<int64_t> &xadj, const uvector<int32_t> &adj,
                              const std::vector<int32_t> &perm, const std::vector<int32_t> &ip,
                              uvector<int64_t> *lptr, uvector<int32_t> *lidx, uvector<int64_t> *hptr, uvector<int32_t> *hidx)
{
    uvector<int32_t> lc((size_t)n);
    run_host_threads(nth, [&](int t) {
        for (int64_t r = n * t / nth; r < n * (t + 1) / nth; ++r) {
            const int32_t v = perm[(size_t)r];
            int32_t c = 0;
            for (int64_t p = xadj[(size_t)v]; p < xadj[(size_t)v + 1]; ++p) c += ip[(size_t)adj[(size_t)p]] < r;
            lc[(size_t)r] = c;
        }
    });
    if (lptr) lptr->resize((size_t)n + 1);
    if (hptr) hptr->resize((size_t)n + 1);
    int64_t lo = 0, hi = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int32_t v = perm[(size_t)r];
        if (lptr) (*lptr)[(size_t)r] = lo;
        if (hptr) (*hptr)[(size_t)r] = hi;
        lo += lc[(size_t)r];
        hi += (xadj[(size_t)v + 1] - xadj[(size_t)v]) - lc[(size_t)r];
    }
    if (lptr) { (*lptr)[(size_t)n] = lo; lidx->resize((size_t)lo); }
    if (hptr) { (*hptr)[(size_t)n] = hi; hidx->resize((size_t)hi); }
    run_host_threads(nth, [&](int t) {
        for (int64_t r = n * t / nth; r < n * (t + 1) / nth; ++r) {
            const int32_t v = perm[(size_t)r];
            int64_t wl = lptr ? (*lptr)[(size_t)r] : 0, wh = hptr ? (*hptr)[(size_t)r] : 0;
            for (int64_t p = xadj[(size_t)v]; p < xadj[(size_t)v + 1]; ++p) {
                const int32_t c = ip[(size_t)adj[(size_t)p]];
                if (c < r) { if (lptr) (*lidx)[(size_t)wl++] = c; }
                else if (hptr) (*hidx)[(size_t)wh++] = c;
            }
        }
    });
}

// mark[v] = region id of v (high word) | tag of the last search that reached v (low word): one load per edge.
inline uint64_t mark_key(int32_t rid, int32_t tag) { return ((uint64_t)(uint32_t)rid << 32) | (uint32_t)tag; }

// BFS over the vertices of region rid, starting at root. Fills `order` (BFS order)
// and `lvl_ptr` (start of each level in `order`). Visited vertices get the tag.
void bfs_levels(const int64_t *xadj, const int32_t *adj,
                std::vector<uint64_t> &mark, int32_t rid, int32_t root,
                int32_t tag, std::vector<int32_t> &order,
                std::vector<int64_t> &lvl_ptr)
{
    const uint64_t visited = mark_key(rid, tag);
    order.clear();
    lvl_ptr.clear();
    order.push_back(root);
    mark[root] = visited;
    lvl_ptr.push_back(0);
    size_t head = 0;
    while (head < order.size()) {
        size_t end = order.size();
        lvl_ptr.push_back((int64_t)end);
        for (; head < end; ++head) {
            int32_t v = order[head];
            for (int64_t p = xadj[v]; p < xadj[v + 1]; ++p) {
                const int32_t w = adj[p];
                const uint64_t mw = mark[w];
                if ((int32_t)(mw >> 32) == rid && mw != visited) {
                    mark[w] = visited;
                    order.push_back(w);
                }
            }
        }
    }
    // lvl_ptr currently has an entry per level start plus the final end
    if (lvl_ptr.back() != (int64_t)order.size()) lvl_ptr.push_back((int64_t)order.size());
}

}  // namespace

void order_nested_dissection(int64_t n, const int64_t *xadj, const int32_t *adj, int leaf_size,
                             std::vector<int32_t> &perm)
{
    perm.assign((size_t)n, -1);
    if (n == 0) return;
    std::vector<uint64_t> mark((size_t)n, 0);
    struct Item { std::vector<int32_t> verts; int64_t lo; bool connected; };
    std::vector<Item> stack;
    {
        Item it;
        it.verts.resize((size_t)n);
        std::iota(it.verts.begin(), it.verts.end(), 0);
        it.lo = 0;
        it.connected = false;
        stack.push_back(std::move(it));
    }
    // Sub-graphs are independent once split, so they are dissected by a small pool of host threads. Tasks touch
    // disjoint entries of mark / perm; region ids and stamp tags come from atomic counters, and a task's
    // result depends only on its vertex set, so the ordering does not depend on the schedule.
    std::atomic<int32_t> next_rid{1}, tag_counter{0};
    std::mutex mu;
    std::condition_variable cv;
    int active = 0;

    auto order_leaf = [&](const std::vector<int32_t> &order_bfs, int64_t lo) {
        // reverse Cuthill-McKee: BFS order reversed
        int64_t nv = (int64_t)order_bfs.size();
        for (int64_t t = 0; t < nv; ++t) perm[(size_t)(lo + t)] = order_bfs[(size_t)(nv - 1 - t)];
    };

    auto push_items = [&](std::vector<Item> &items) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &c : items) stack.push_back(std::move(c));
        cv.notify_all();
    };
    auto process = [&](Item &it, std::vector<int32_t> &order, std::vector<int32_t> &best_order, std::vector<int64_t> &lvl,
                       std::vector<int64_t> &best_lvl) {
        int32_t tag = 0;
        const int64_t nv = (int64_t)it.verts.size();
        if (nv == 0) return;
        const int32_t rid = next_rid.fetch_add(1);
        for (int32_t v : it.verts) mark[v] = mark_key(rid, 0);

        // pseudo-peripheral root: start at a minimum-degree vertex, iterate on the last level. The first search doubles
        // as the connectivity test of a freshly split part: if it does not reach every vertex, the part is cut into its
        // connected components instead.
        int32_t root = it.verts[0];
        {
            int64_t bestdeg = INT64_MAX;
            for (int32_t v : it.verts) {
                int64_t d = xadj[v + 1] - xadj[v];
                if (d < bestdeg) { bestdeg = d; root = v; }
            }
        }
        int64_t best_h = -1;
        for (int round = 0; round < 6; ++round) {
            tag = tag_counter.fetch_add(1) + 1;
            bfs_levels(xadj, adj, mark, rid, root, tag, order, lvl);
            if (round == 0 && !it.connected && (int64_t)order.size() != nv) {
                tag = tag_counter.fetch_add(1) + 1;
                int64_t lo = it.lo;
                std::vector<Item> comps;
                for (int32_t v : it.verts) {
                    if (mark[v] == mark_key(rid, tag)) continue;
                    bfs_levels(xadj, adj, mark, rid, v, tag, order, lvl);
                    Item c;
                    c.verts = order;
                    c.lo = lo;
                    c.connected = true;
                    lo += (int64_t)order.size();
                    comps.push_back(std::move(c));
                }
                push_items(comps);
                return;
            }
            int64_t h = (int64_t)lvl.size() - 1;
            if (h <= best_h) break;
            best_h = h;
            best_order = order;
            best_lvl = lvl;
            // next root: minimum degree vertex in the last level
            int64_t bestdeg = INT64_MAX;
            for (int64_t t = lvl[(size_t)h - 1]; t < lvl[(size_t)h]; ++t) {
                int32_t v = order[(size_t)t];
                int64_t d = xadj[v + 1] - xadj[v];
                if (d < bestdeg) { bestdeg = d; root = v; }
            }
        }
        const int64_t nlev = best_h;
        if (nv <= leaf_size || nlev < 3) {
            order_leaf(best_order, it.lo);
            return;
        }
        // choose the separator level: balance the two sides, prefer small separators
        int64_t bestj = 1;
        double bestcost = 1e300;
        for (int64_t j = 1; j <= nlev - 2; ++j) {
            int64_t before = best_lvl[(size_t)j];
            int64_t sep = best_lvl[(size_t)j + 1] - best_lvl[(size_t)j];
            int64_t after = nv - before - sep;
            double imb = (double)std::max(before, after) / (double)nv;   // 0.5 = perfect
            double cost = imb + 1.0 * (double)sep / (double)nv;
            if (cost < bestcost) { bestcost = cost; bestj = j; }
        }
        const int64_t j = bestj;
        // thin the separator: keep only level-j vertices adjacent to level j+1
        tag = tag_counter.fetch_add(1) + 1;  // stamp level j+1 vertices with tag
        for (int64_t t = best_lvl[(size_t)j + 1]; t < best_lvl[(size_t)j + 2]; ++t)
            mark[best_order[(size_t)t]] = mark_key(rid, tag);
        Item A, B;
        std::vector<int32_t> S;
        A.verts.assign(best_order.begin(), best_order.begin() + best_lvl[(size_t)j]);
        for (int64_t t = best_lvl[(size_t)j]; t < best_lvl[(size_t)j + 1]; ++t) {
            int32_t v = best_order[(size_t)t];
            bool touches = false;
            for (int64_t p = xadj[v]; p < xadj[v + 1] && !touches; ++p) {
                int32_t w = adj[p];
                touches = (mark[w] == mark_key(rid, tag));
            }
            if (touches) S.push_back(v); else A.verts.push_back(v);
        }
        B.verts.assign(best_order.begin() + best_lvl[(size_t)j + 1], best_order.end());
        A.lo = it.lo;
        B.lo = it.lo + (int64_t)A.verts.size();
        A.connected = false;
        B.connected = false;
        int64_t slo = B.lo + (int64_t)B.verts.size();
        for (size_t t = 0; t < S.size(); ++t) perm[(size_t)slo + t] = S[t];
        std::vector<Item> two;
        two.push_back(std::move(A));
        two.push_back(std::move(B));
        push_items(two);
    };
    auto worker = [&](int) {
        std::vector<int32_t> order, best_order;
        std::vector<int64_t> lvl, best_lvl;
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !stack.empty() || active == 0; });
                if (stack.empty()) return;          // nothing queued and nobody working: done
                it = std::move(stack.back());
                stack.pop_back();
                ++active;
            }
            process(it, order, best_order, lvl, best_lvl);
            {
                std::lock_guard<std::mutex> lk(mu);
                --active;
                cv.notify_all();
            }
        }
    };
    run_host_threads((int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, n / 20000)), worker);
}



static inline int64_t trap(int64_t k, int64_t r) { return k * (k + 1) / 2 + k * r; }

// Full symmetric CSR of the input matrix with the position of every value in the caller's lower-CSC array: the operator
// of the refinement residual r = b - K x. Rows come out sorted (two column-major sweeps).
void ls_build_full_csr(LsSymbolic &S)
{
    const int64_t n = S.n;
    const int32_t *colptr = S.in_colptr.data(), *rowval = S.in_rowval.data();
    S.full_ptr.assign((size_t)n + 1, 0);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            S.full_ptr[(size_t)i + 1]++;
            if (i != j) S.full_ptr[(size_t)j + 1]++;
        }
    for (int64_t i = 0; i < n; ++i) S.full_ptr[(size_t)i + 1] += S.full_ptr[(size_t)i];
    S.full_col.resize((size_t)S.full_ptr[(size_t)n]);
    S.full_val.resize((size_t)S.full_ptr[(size_t)n]);
    std::vector<int64_t> pf(S.full_ptr.begin(), S.full_ptr.end() - 1);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            S.full_col[(size_t)pf[(size_t)i]] = (int32_t)j;
            S.full_val[(size_t)pf[(size_t)i]++] = p;
        }
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            if (i != j) {
                S.full_col[(size_t)pf[(size_t)j]] = i;
                S.full_val[(size_t)pf[(size_t)j]++] = p;
            }
        }
}

// Scatter map of the input nonzeros into the panels (independent per column: host threads over column chunks). Handles
// with a GPU build the same map there (k_build_a2l, factor.cu).
void ls_build_a2l(LsSymbolic &S, std::string &err)
{
    const int64_t n = S.n, nnz = S.nnz_a;
    const int32_t *colptr = S.in_colptr.data(), *rowval = S.in_rowval.data();
    S.a2l.resize((size_t)nnz);
    const int nth = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, n / 8192));
    std::vector<int> bad((size_t)nth, 0);
    auto work = [&](int t) {
        const int64_t j0 = n * t / nth, j1 = n * (t + 1) / nth;
        for (int64_t j = j0; j < j1; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                int32_t a = S.iperm[(size_t)rowval[p]], b = S.iperm[(size_t)j];
                int32_t r = std::max(a, b), c = std::min(a, b);
                int32_t s = S.col2sn[(size_t)c];
                int32_t c0 = S.sn_ptr[(size_t)s], c1 = S.sn_ptr[(size_t)s + 1];
                int64_t k = c1 - c0;
                int64_t nr = S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s];
                int64_t tt;
                if (r < c1) {
                    tt = r - c0;
                } else {
                    const int32_t *rb = S.row_idx.data() + S.row_ptr[(size_t)s];
                    const int32_t *f = std::lower_bound(rb, rb + nr, r);
                    if (f == rb + nr || *f != r) { bad[(size_t)t] = 1; return; }
                    tt = k + (f - rb);
                }
                S.a2l[(size_t)p] = S.lp[(size_t)s] + (int64_t)(c - c0) * ((k + nr + 1) & ~(int64_t)1) + tt;
            }
    };
    run_host_threads(nth, work);
    for (int v : bad) if (v) err = "input entry outside the symbolic structure (internal error)";
}

#define TLOG(name) do { if (tlog) { auto now_ = std::chrono::steady_clock::now(); std::fprintf(stderr, "analyze stage before %s: %.3f s\n", name, std::chrono::duration<double>(now_ - t_prev).count()); t_prev = now_; } } while (0)
std::string ls_analyze(int64_t n, const int32_t *colptr_in, const int32_t *rowval_in, int index_base,
                       const LsOptions &opt, const int32_t *user_perm, LsSymbolic &S)
{
    const bool tlog = std::getenv("MIPM_ANALYZE_LOG") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    S = LsSymbolic();
    S.n = n;
    S.kind = opt.kind;
    if (n < 0) return "negative dimension";
    const int64_t nnz = n > 0 ? (int64_t)colptr_in[n] - index_base : 0;
    if (nnz < 0) return "bad column pointer";
    S.nnz_a = nnz;
    // 0-based copy of the pattern, kept for the refinement operator and the device-side scatter map; everything below
    // works on the copy
    S.in_colptr.resize((size_t)n + 1);
    S.in_rowval.resize((size_t)nnz);
    {
        const int nt0 = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), nnz / 65536));
        run_host_threads(nt0, [&](int t) {
            for (int64_t j = (n + 1) * t / nt0; j < (n + 1) * (t + 1) / nt0; ++j) S.in_colptr[(size_t)j] = colptr_in[j] - index_base;
            for (int64_t p = nnz * t / nt0; p < nnz * (t + 1) / nt0; ++p) S.in_rowval[(size_t)p] = rowval_in[p] - index_base;
        });
    }
    const int32_t *colptr = S.in_colptr.data(), *rowval = S.in_rowval.data();
    // Column chunks with about equal numbers of stored entries for the host threads of the passes below
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(host_threads(), n / 4096), (int64_t)(1 << 25) / std::max<int64_t>(n, 1)));
    std::vector<int64_t> chunk((size_t)nth + 1, n);
    chunk[0] = 0;
    if (n > 0 && colptr[0] != 0) return "column pointer does not start at index_base";
    for (int64_t j = 0; j < n; ++j)
        if (colptr[j + 1] < colptr[j]) return "colptr not monotone";
    for (int t = 1; t < nth; ++t)
        chunk[(size_t)t] = std::lower_bound(colptr, colptr + n + 1, (int32_t)(nnz * t / nth)) - colptr;
    for (int t = 1; t <= nth; ++t) chunk[(size_t)t] = std::min<int64_t>(std::max(chunk[(size_t)t], chunk[(size_t)t - 1]), n);
    chunk[(size_t)nth] = n;
    // ---- full symmetric adjacency (no diagonal). The full CSR with value positions (iterative refinement only) is
    // built on first use from the kept copy of the pattern: ls_build_full_csr.
    // Transposition by column chunks: thread t counts, per row, the entries of its columns (cntT[t][i]); a prefix over
    // the threads turns the counts into write cursors, so every adjacency list comes out as {j < i ascending} followed by
    // {j > i in column order}, exactly what two sequential column-major sweeps produce.
    std::vector<int64_t> xadj((size_t)n + 1, 0);
    uvector<int32_t> adj;
    {
        std::vector<int32_t> cntT((size_t)nth * (size_t)n, 0), lowcnt((size_t)n, 0);
        std::vector<int> bad((size_t)nth, 0);
        run_host_threads(nth, [&](int t) {
            int32_t *ct = cntT.data() + (size_t)t * (size_t)n;
            for (int64_t j = chunk[(size_t)t]; j < chunk[(size_t)t + 1]; ++j) {
                int32_t lc = 0;
                for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                    const int32_t i = rowval[p];
                    if (i < j || i >= n) { bad[(size_t)t] = 1; return; }
                    if (i != j) { ct[i]++; ++lc; }
                }
                lowcnt[(size_t)j] = lc;
            }
        });
        for (int v : bad) if (v) return "rowval outside the lower triangle";
        // per row: counts -> exclusive prefix over the threads; degree
        run_host_threads(nth, [&](int t) {
            for (int64_t i = n * t / nth; i < n * (t + 1) / nth; ++i) {
                int32_t run = 0;
                for (int u = 0; u < nth; ++u) {
                    int32_t &c = cntT[(size_t)u * (size_t)n + (size_t)i];
                    const int32_t v = c;
                    c = run;
                    run += v;
                }
                xadj[(size_t)i + 1] = (int64_t)run + lowcnt[(size_t)i];
            }
        });
        for (int64_t i = 0; i < n; ++i) xadj[(size_t)i + 1] += xadj[(size_t)i];
        adj.resize((size_t)xadj[(size_t)n]);
        run_host_threads(nth, [&](int t) {
            int32_t *ct = cntT.data() + (size_t)t * (size_t)n;
            for (int64_t j = chunk[(size_t)t]; j < chunk[(size_t)t + 1]; ++j)
                for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                    const int32_t i = rowval[p];
                    if (i != j) adj[(size_t)(xadj[(size_t)i] + ct[i]++)] = (int32_t)j;
                }
        });
        run_host_threads(nth, [&](int t) {
            for (int64_t j = chunk[(size_t)t]; j < chunk[(size_t)t + 1]; ++j) {
                int64_t w = xadj[(size_t)j + 1] - lowcnt[(size_t)j];
                for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p)
                    if (rowval[p] != j) adj[(size_t)w++] = rowval[p];
            }
        });
    }
    TLOG("0");
    // ---- ordering
    std::vector<int32_t> perm0;
    if (opt.ordering == 2 /*USER*/) {
        if (!user_perm) return "user ordering requested without a permutation";
        perm0.assign(user_perm, user_perm + n);
        std::vector<char> seen((size_t)n, 0);
        for (int64_t k = 0; k < n; ++k) {
            if (perm0[(size_t)k] < 0 || perm0[(size_t)k] >= n || seen[(size_t)perm0[(size_t)k]]) return "user permutation invalid";
            seen[(size_t)perm0[(size_t)k]] = 1;
        }
    } else if (opt.ordering == 1 /*NATURAL*/) {
        perm0.resize((size_t)n);
        std::iota(perm0.begin(), perm0.end(), 0);
    } else if (opt.n_border > 0) {
        // nested dissection of the interior subgraph only; border vertices stay last in natural order
        const int64_t ni = n - opt.n_border;
        if (ni < 0) return "border larger than the matrix";
        std::vector<int64_t> xi((size_t)ni + 1, 0);
        std::vector<int32_t> ai;
        for (int64_t v = 0; v < ni; ++v) {
            for (int64_t p = xadj[(size_t)v]; p < xadj[(size_t)v + 1]; ++p)
                if (adj[(size_t)p] < ni) ai.push_back(adj[(size_t)p]);
            xi[(size_t)v + 1] = (int64_t)ai.size();
        }
        order_nested_dissection(ni, xi.data(), ai.data(), opt.nd_leaf, perm0);
        for (int64_t v = ni; v < n; ++v) perm0.push_back((int32_t)v);
    } else {
        order_nested_dissection(n, xadj.data(), adj.data(), opt.nd_leaf, perm0);
    }
    if (opt.kind == 1 /*LDL*/ && opt.ordering != 2) {
        // Quasi-definite safeguard for K2 = [Q+Sigma A'; A delta_c I]: a vertex whose diagonal
        // block is only the (tiny) dual regularization must not be eliminated before all of
        // its neighbours. Such vertices are those without a structural coupling inside their
        // own block; we cannot see values here, so the caller marks nothing and we use the
        // structural rule "delay a vertex until one neighbour is eliminated" for vertices of
        // the trailing block, detected as vertices whose neighbours all have smaller index
        // (rows of A only touch primal columns, which come first in K2's numbering).
        std::vector<int32_t> ip((size_t)n);
        for (int64_t k = 0; k < n; ++k) ip[(size_t)perm0[(size_t)k]] = (int32_t)k;
        std::vector<char> dual((size_t)n, 0);
        for (int64_t v = 0; v < n; ++v) {
            bool all_smaller = xadj[(size_t)v + 1] > xadj[(size_t)v];
            for (int64_t p = xadj[(size_t)v]; p < xadj[(size_t)v + 1] && all_smaller; ++p)
                all_smaller = adj[(size_t)p] < v;
            dual[(size_t)v] = all_smaller;
        }
        const bool delay_all = opt.ldl_delay_all;
        // primal vertices with all neighbours larger are not delayed; only "dual" ones are
        std::vector<char> done((size_t)n, 0);
        std::vector<std::vector<int32_t>> waiting((size_t)n);  // waiting[u]: duals released when u is eliminated
        std::vector<int32_t> out;
        out.reserve((size_t)n);
        for (int64_t k = 0; k < n; ++k) {
            int32_t v = perm0[(size_t)k];
            if (dual[(size_t)v]) {
                // wait for the LAST not-yet-eliminated primal neighbour (delay_all) or the first one
                int32_t wait_nb = -1;
                int32_t wait_pos = delay_all ? -1 : INT32_MAX;
                bool ready = !delay_all ? false : true;
                for (int64_t p = xadj[(size_t)v]; p < xadj[(size_t)v + 1]; ++p) {
                    int32_t u = adj[(size_t)p];
                    if (dual[(size_t)u]) continue;
                    if (done[(size_t)u]) { if (!delay_all) { ready = true; break; } else continue; }
                    if (delay_all) {
                        ready = false;
                        if (ip[(size_t)u] > wait_pos) { wait_pos = ip[(size_t)u]; wait_nb = u; }
                    } else if (ip[(size_t)u] < wait_pos) { wait_pos = ip[(size_t)u]; wait_nb = u; }
                }
                if (!ready && wait_nb >= 0) { waiting[(size_t)wait_nb].push_back(v); continue; }
            }
            out.push_back(v);
            done[(size_t)v] = 1;
            for (int32_t w : waiting[(size_t)v]) { out.push_back(w); done[(size_t)w] = 1; }
            waiting[(size_t)v].clear();
        }
        perm0.swap(out);
    }
    std::vector<int32_t> ip0((size_t)n);
    for (int64_t k = 0; k < n; ++k) ip0[(size_t)perm0[(size_t)k]] = (int32_t)k;

    TLOG("1");
    // ---- adjacency in the perm0 numbering split at the diagonal: lowadj[r] = {c < r} (rows of the lower triangle, for
    // the elimination tree), below0[c] = {r > c} (columns, for the column counts)
    uvector<int64_t> lptr, b0ptr;
    uvector<int32_t> lidx, b0idx;
    split_permuted_adjacency(n, nth, xadj, adj, perm0, ip0, &lptr, &lidx, &b0ptr, &b0idx);

    TLOG("2");
    // ---- elimination tree (Liu, path compression through `anc`) ...
    std::vector<int32_t> parent((size_t)n, -1);
    std::vector<int64_t> cnt((size_t)n, 0);
    {
        std::vector<int32_t> anc((size_t)n, -1);
        for (int64_t k = 0; k < n; ++k)
            for (int64_t p = lptr[(size_t)k]; p < lptr[(size_t)k + 1]; ++p) {
                int32_t i = lidx[(size_t)p];
                while (i != -1 && i < k) {
                    const int32_t nx = anc[(size_t)i];
                    anc[(size_t)i] = (int32_t)k;
                    if (nx == -1) parent[(size_t)i] = (int32_t)k;
                    i = nx;
                }
            }
        // ... and off-diagonal column counts in O(nnz alpha) (Gilbert, Ng & Peyton: skeleton leaves of the row subtrees
        // found through first descendants in a postorder, overlaps removed at least common ancestors)
        std::vector<int32_t> post1((size_t)n), first((size_t)n, -1), maxfirst((size_t)n, -1), prevleaf((size_t)n, -1);
        {
            // any postorder: children lists by counting sort, iterative DFS
            std::vector<int32_t> head((size_t)n + 1, -1), next((size_t)n, -1), stk;
            for (int64_t v = n - 1; v >= 0; --v) {           // children of a node in ascending order
                const int32_t pv = parent[(size_t)v];
                next[(size_t)v] = head[(size_t)(pv + 1)];
                head[(size_t)(pv + 1)] = (int32_t)v;
            }
            int64_t k = 0;
            for (int32_t r = head[0]; r != -1; r = next[(size_t)r]) {
                stk.push_back(r);
                while (!stk.empty()) {
                    const int32_t v = stk.back();
                    const int32_t c = head[(size_t)v + 1];
                    if (c != -1) { head[(size_t)v + 1] = next[(size_t)c]; stk.push_back(c); }
                    else { post1[(size_t)k++] = v; stk.pop_back(); }
                }
            }
            if (k != n) return "postorder failed";
        }
        std::vector<int64_t> &delta = cnt;
        for (int64_t k = 0; k < n; ++k) {
            int32_t j = post1[(size_t)k];
            delta[(size_t)j] = (first[(size_t)j] == -1) ? 1 : 0;        // leaves of the tree start with their diagonal
            for (; j != -1 && first[(size_t)j] == -1; j = parent[(size_t)j]) first[(size_t)j] = (int32_t)k;
        }
        for (int64_t i = 0; i < n; ++i) anc[(size_t)i] = (int32_t)i;
        for (int64_t k = 0; k < n; ++k) {
            const int32_t j = post1[(size_t)k];
            if (parent[(size_t)j] != -1) delta[(size_t)parent[(size_t)j]]--;
            for (int64_t p = b0ptr[(size_t)j]; p < b0ptr[(size_t)j + 1]; ++p) {
                const int32_t i = b0idx[(size_t)p];                     // entry (i, j), i > j
                if (first[(size_t)j] <= maxfirst[(size_t)i]) continue;  // j is not a leaf of the row subtree of i
                maxfirst[(size_t)i] = first[(size_t)j];
                const int32_t jprev = prevleaf[(size_t)i];
                prevleaf[(size_t)i] = j;
                delta[(size_t)j]++;
                if (jprev != -1) {                                      // subsequent leaf: remove the overlap at the lca
                    int32_t q = jprev;
                    while (q != anc[(size_t)q]) q = anc[(size_t)q];
                    for (int32_t sidx = jprev; sidx != q;) {
                        const int32_t sp = anc[(size_t)sidx];
                        anc[(size_t)sidx] = q;
                        sidx = sp;
                    }
                    delta[(size_t)q]--;
                }
            }
            if (parent[(size_t)j] != -1) anc[(size_t)j] = parent[(size_t)j];
        }
        for (int64_t j = 0; j < n; ++j)                                 // parents have larger indices than children
            if (parent[(size_t)j] != -1) cnt[(size_t)parent[(size_t)j]] += cnt[(size_t)j];
        for (int64_t j = 0; j < n; ++j) cnt[(size_t)j] -= 1;           // off-diagonal count
    }
    if (opt.n_border > 0) {
        // treat the border block as structurally dense: a chain in the elimination tree with the column
        // counts of a dense trailing block (a superset of the true structure, filled with explicit zeros)
        for (int64_t j = n - opt.n_border; j < n; ++j) {
            parent[(size_t)j] = (j + 1 < n) ? (int32_t)(j + 1) : -1;
            cnt[(size_t)j] = n - 1 - j;
        }
    }
    TLOG("3");
    // ---- postorder (children by ascending count so a supernode-forming child comes last)
    std::vector<int32_t> post((size_t)n);  // post[newpos] = node (in perm0 numbering)
    {
        std::vector<int64_t> cptr((size_t)n + 2, 0);
        for (int64_t v = 0; v < n; ++v) cptr[(size_t)(parent[(size_t)v] + 1) + 1]++;   // slot 0 = roots
        for (int64_t i = 0; i <= n; ++i) cptr[(size_t)i + 1] += cptr[(size_t)i];
        std::vector<int32_t> cidx((size_t)n);
        std::vector<int64_t> pos(cptr.begin(), cptr.end() - 1);
        for (int64_t v = 0; v < n; ++v) cidx[(size_t)pos[(size_t)(parent[(size_t)v] + 1)]++] = (int32_t)v;
        for (int64_t q = 0; q <= n; ++q)
            std::stable_sort(cidx.begin() + cptr[(size_t)q], cidx.begin() + cptr[(size_t)q + 1],
                             [&](int32_t a, int32_t b) { return cnt[(size_t)a] < cnt[(size_t)b]; });
        // iterative DFS
        std::vector<int32_t> stk;
        std::vector<int64_t> it((size_t)n + 1);
        for (int64_t q = 0; q <= n; ++q) it[(size_t)q] = cptr[(size_t)q];
        int64_t k = 0;
        for (int64_t r = cptr[0]; r < cptr[1]; ++r) {
            stk.push_back(cidx[(size_t)r]);
            while (!stk.empty()) {
                int32_t v = stk.back();
                int64_t &ci = it[(size_t)v + 1];
                if (ci < cptr[(size_t)v + 2]) {
                    stk.push_back(cidx[(size_t)ci++]);
                } else {
                    post[(size_t)k++] = v;
                    stk.pop_back();
                }
            }
        }
        if (k != n) return "postorder failed";
    }
    S.perm.resize((size_t)n);
    S.iperm.resize((size_t)n);
    std::vector<int32_t> ipost((size_t)n);
    for (int64_t k = 0; k < n; ++k) ipost[(size_t)post[(size_t)k]] = (int32_t)k;
    for (int64_t k = 0; k < n; ++k) S.perm[(size_t)k] = perm0[(size_t)post[(size_t)k]];
    for (int64_t k = 0; k < n; ++k) S.iperm[(size_t)S.perm[(size_t)k]] = (int32_t)k;
    std::vector<int32_t> par((size_t)n);
    std::vector<int64_t> cc((size_t)n);
    for (int64_t v = 0; v < n; ++v) {
        par[(size_t)ipost[(size_t)v]] = parent[(size_t)v] < 0 ? -1 : ipost[(size_t)parent[(size_t)v]];
        cc[(size_t)ipost[(size_t)v]] = cnt[(size_t)v];
    }
    S.nnz_l_exact = 0;
    S.flops = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        S.nnz_l_exact += cc[(size_t)j] + 1;
        S.flops += (double)(cc[(size_t)j] + 1) * (double)(cc[(size_t)j] + 1);
    }

    TLOG("4");
    // ---- fundamental supernodes
    std::vector<int32_t> sn0;  // start columns
    for (int64_t j = 0; j < n; ++j) {
        const bool at_border = opt.n_border > 0 && j == n - opt.n_border;       // the border is its own supernode
        bool cont = j > 0 && !at_border && par[(size_t)j - 1] == j && cc[(size_t)j - 1] == cc[(size_t)j] + 1 &&
                    (j - sn0.back()) < opt.max_sn_cols;
        if (!cont) sn0.push_back((int32_t)j);
    }
    int64_t ns0 = (int64_t)sn0.size();
    sn0.push_back((int32_t)n);

    // ---- relaxed amalgamation on (start, end, rows_below = cc[end-1], true nnz)
    struct Sn { int32_t c0, c1; int64_t r; int64_t nnz_true; };
    std::vector<Sn> cur;
    cur.reserve((size_t)ns0);
    for (int64_t s = 0; s < ns0; ++s) {
        Sn x;
        x.c0 = sn0[(size_t)s];
        x.c1 = sn0[(size_t)s + 1];
        x.r = cc[(size_t)x.c1 - 1];
        x.nnz_true = 0;
        for (int32_t j = x.c0; j < x.c1; ++j) x.nnz_true += cc[(size_t)j] + 1;
        while (!cur.empty()) {
            Sn &pv = cur.back();
            // pv is a child of x iff the parent column of pv's last column lies inside x
            int32_t pc = par[(size_t)pv.c1 - 1];
            if (pv.c1 != x.c0 || pc < x.c0 || pc >= x.c1) break;
            int64_t k = (int64_t)(x.c1 - pv.c0);
            if (k > opt.max_sn_cols) break;
            if (opt.n_border > 0 && x.c1 > n - opt.n_border && pv.c0 < n - opt.n_border) break;   // interior | border
            int64_t merged = trap(k, x.r);
            int64_t tru = pv.nnz_true + x.nnz_true;
            double z = (double)(merged - tru) / (double)merged;
            bool ok = (k <= opt.relax_always) || (k <= opt.relax_k1 && z <= opt.relax_z1) ||
                      (k <= opt.relax_k2 && z <= opt.relax_z2) || (z <= opt.relax_z3);
            if (!ok) break;
            x.c0 = pv.c0;
            x.nnz_true = tru;
            cur.pop_back();
        }
        cur.push_back(x);
    }
    S.ns = (int32_t)cur.size();
    const int32_t ns = S.ns;
    S.sn_ptr.resize((size_t)ns + 1);
    for (int32_t s = 0; s < ns; ++s) S.sn_ptr[(size_t)s] = cur[(size_t)s].c0;
    S.sn_ptr[(size_t)ns] = (int32_t)n;
    S.col2sn.resize((size_t)n);
    for (int32_t s = 0; s < ns; ++s)
        for (int32_t j = S.sn_ptr[(size_t)s]; j < S.sn_ptr[(size_t)s + 1]; ++j) S.col2sn[(size_t)j] = s;
    if (opt.n_border > 0 && n > 0) {
        S.root_sn = S.col2sn[(size_t)n - 1];
        if (S.sn_ptr[(size_t)S.root_sn] != n - opt.n_border || S.sn_ptr[(size_t)S.root_sn + 1] != n)
            return "border did not end up as one final supernode (internal error)";
        for (int64_t j = n - opt.n_border; j < n; ++j)
            if (S.perm[(size_t)j] != (int32_t)j) return "border vertices were reordered (internal error)";
    }

    TLOG("5");
    // ---- strictly-lower column structure of the permuted matrix: below[c] = {r > c}
    uvector<int64_t> bptr;
    uvector<int32_t> bidx;
    split_permuted_adjacency(n, nth, xadj, adj, S.perm, S.iperm, nullptr, nullptr, &bptr, &bidx);
    TLOG("6");
    // ---- supernode parents, children, row structures (merge original entries + children)
    S.sn_parent.assign((size_t)ns, -1);
    S.row_ptr.assign((size_t)ns + 1, 0);
    std::vector<std::vector<int32_t>> rows((size_t)ns);
    std::vector<std::vector<int32_t>> kids((size_t)ns);
    {
        std::vector<int32_t> mark((size_t)n, -1);
        for (int32_t s = 0; s < ns; ++s) {
            int32_t c0 = S.sn_ptr[(size_t)s], c1 = S.sn_ptr[(size_t)s + 1];
            std::vector<int32_t> &R = rows[(size_t)s];
            for (int32_t j = c0; j < c1; ++j)
                for (int64_t p = bptr[(size_t)j]; p < bptr[(size_t)j + 1]; ++p) {
                    int32_t r = bidx[(size_t)p];
                    if (r >= c1 && mark[(size_t)r] != s) { mark[(size_t)r] = s; R.push_back(r); }
                }
            for (int32_t c : kids[(size_t)s])
                for (int32_t r : rows[(size_t)c])
                    if (r >= c1 && mark[(size_t)r] != s) { mark[(size_t)r] = s; R.push_back(r); }
            std::sort(R.begin(), R.end());
            if ((int64_t)R.size() != cur[(size_t)s].r) return "supernode row count mismatch (internal error)";
            if (!R.empty()) {
                int32_t ps = S.col2sn[(size_t)R[0]];
                S.sn_parent[(size_t)s] = ps;
                kids[(size_t)ps].push_back(s);
            }
            S.row_ptr[(size_t)s + 1] = S.row_ptr[(size_t)s] + (int64_t)R.size();
        }
    }
    S.row_idx.resize((size_t)S.row_ptr[(size_t)ns]);
    S.rel_idx.resize((size_t)S.row_ptr[(size_t)ns]);
    S.child_ptr.assign((size_t)ns + 1, 0);
    for (int32_t s = 0; s < ns; ++s) {
        std::copy(rows[(size_t)s].begin(), rows[(size_t)s].end(), S.row_idx.begin() + S.row_ptr[(size_t)s]);
        S.child_ptr[(size_t)s + 1] = S.child_ptr[(size_t)s] + (int64_t)kids[(size_t)s].size();
    }
    S.child_idx.resize((size_t)S.child_ptr[(size_t)ns]);
    for (int32_t s = 0; s < ns; ++s)
        std::copy(kids[(size_t)s].begin(), kids[(size_t)s].end(), S.child_idx.begin() + S.child_ptr[(size_t)s]);
    // relative indices of each supernode's rows inside its parent's front (cols ++ rows)
    for (int32_t s = 0; s < ns; ++s) {
        int32_t ps = S.sn_parent[(size_t)s];
        if (ps < 0) continue;
        int32_t pc0 = S.sn_ptr[(size_t)ps], pc1 = S.sn_ptr[(size_t)ps + 1];
        int32_t pk = pc1 - pc0;
        const std::vector<int32_t> &PR = rows[(size_t)ps];
        size_t q = 0;
        for (int64_t t = S.row_ptr[(size_t)s]; t < S.row_ptr[(size_t)s + 1]; ++t) {
            int32_t r = S.row_idx[(size_t)t];
            if (r < pc1) {
                S.rel_idx[(size_t)t] = r - pc0;
            } else {
                while (q < PR.size() && PR[q] < r) ++q;
                if (q >= PR.size() || PR[q] != r) return "front structure not nested (internal error)";
                S.rel_idx[(size_t)t] = pk + (int32_t)q;
            }
        }
    }
    TLOG("7");
    // ---- storage offsets, levels, stats
    S.lp.assign((size_t)ns + 1, 0);
    S.up.assign((size_t)ns + 1, 0);
    S.sn_level.assign((size_t)ns, 0);
    for (int32_t s = 0; s < ns; ++s) {
        int64_t k = S.sn_ptr[(size_t)s + 1] - S.sn_ptr[(size_t)s];
        int64_t r = S.row_ptr[(size_t)s + 1] - S.row_ptr[(size_t)s];
        // leading dimension k + r rounded up to even: every panel column starts on a 16-byte boundary (TMA bulk copies)
        S.lp[(size_t)s + 1] = S.lp[(size_t)s] + ((k + r + 1) & ~(int64_t)1) * k;
        S.up[(size_t)s + 1] = S.up[(size_t)s] + r * r;
        S.max_front_cols = std::max<int32_t>(S.max_front_cols, (int32_t)k);
        S.max_front_rows = std::max<int32_t>(S.max_front_rows, (int32_t)(k + r));
        for (int32_t c : kids[(size_t)s]) S.sn_level[(size_t)s] = std::max(S.sn_level[(size_t)s], S.sn_level[(size_t)c] + 1);
        S.n_levels = std::max(S.n_levels, S.sn_level[(size_t)s] + 1);
    }
    if (S.root_sn >= 0) {       // the border front sits alone on its own top level (staged factorization / solve)
        S.sn_level[(size_t)S.root_sn] = S.n_levels;
        S.n_levels += 1;
    }
    S.nnz_l = S.lp[(size_t)ns];
    S.update_doubles = S.up[(size_t)ns];
    S.level_ptr.assign((size_t)S.n_levels + 1, 0);
    for (int32_t s = 0; s < ns; ++s) S.level_ptr[(size_t)S.sn_level[(size_t)s] + 1]++;
    for (int32_t l = 0; l < S.n_levels; ++l) S.level_ptr[(size_t)l + 1] += S.level_ptr[(size_t)l];
    S.level_sn.resize((size_t)ns);
    {
        std::vector<int64_t> pos(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (int32_t s = 0; s < ns; ++s) S.level_sn[(size_t)pos[(size_t)S.sn_level[(size_t)s]]++] = s;
    }
    TLOG("8");
    TLOG("9");
    if (opt.host_a2l) {
        std::string err;
        ls_build_a2l(S, err);
        if (!err.empty()) return err;
    }
    TLOG("10");
    return "";
}

}  // namespace mipm
