// Device build of the normal-equations pattern tril(A A') and of its product-term map (one-time setup).
//
// Replaces the host sweep of host_symbolic.cpp on handles that own a GPU; the result is the same canonical object
// (sorted lower CSC, terms of an entry ordered by their position in row j: the order of the reference loop,
// src/utils.jl:288-301), so both builders are interchangeable bit for bit (tests/test_gpu_parity.py compares them).
// The reference builds this structure with scalar host loops (src/utils.jl:209-274) and, on the GPU, with a
// per-entry merge kernel (ext/MadIPMCUDAExt/cuda_wrapper.jl:158-234); here it is
//   1. CSC index of A with CSR positions: the one mipm_spmv_setup registered on the handle (Ap = Aj = NULL), or a
//      stable bucket sort on the host threads (rows ascend inside a column);
//   2. one thread per stored entry (i, k): number of entries (j, k), j >= i, of column k       -> exclusive scan
//   3. one CTA per row i: emit the terms (j, p_j | p_i) of the row, bitonic sort by (j, p_j) in shared memory
//      (rows with more terms than fit are sorted in a global workspace by the same code), write the sorted term
//      arrays, compact the first term of every distinct j                                      -> exclusive scan
//   4. one warp per row: move the compacted heads to their final place (Cj, term_ptr);
//   5. row runs of the streaming assembly from a flag + scan pass over term_ptr.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"

namespace mipm {

namespace {

constexpr int SCAN_TB = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_TB * SCAN_ITEMS;

__device__ __forceinline__ long long block_excl_scan_ll(long long v, long long *sh /*33*/, long long &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        long long s = lane < nw ? sh[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < nw) sh[lane] = s;
        if (lane == 31) sh[32] = s;
    }
    __syncthreads();
    const long long base = w > 0 ? sh[w - 1] : 0;
    total = sh[32];
    __syncthreads();
    return base + x - v;
}

// pass 1: per-tile sums
__global__ void __launch_bounds__(SCAN_TB) k_scan_sums(int64_t n, const int32_t *__restrict__ in, long long *__restrict__ sums)
{
    __shared__ long long sh[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    long long s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q)
        if (base + q < n) s += in[base + q];
    long long total;
    block_excl_scan_ll(s, sh, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// pass 2: one CTA scans the tile sums in place (exclusive) and writes the grand total behind them
__global__ void __launch_bounds__(SCAN_TB) k_scan_tiles(int64_t nt, long long *__restrict__ sums)
{
    __shared__ long long sh[33];
    long long carry = 0;
    for (int64_t b = 0; b < nt; b += SCAN_TB) {
        const int64_t i = b + threadIdx.x;
        const long long v = i < nt ? sums[i] : 0;
        long long total;
        const long long e = block_excl_scan_ll(v, sh, total);
        if (i < nt) sums[i] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0) sums[nt] = carry;
}

// pass 3: out[i] = exclusive prefix (n + 1 entries; out[n] = total). Offsets are stored in 32 bits: the caller checks
// the 64-bit grand total before it uses them.
__global__ void __launch_bounds__(SCAN_TB) k_scan_write(int64_t n, const int32_t *__restrict__ in, const long long *__restrict__ sums,
                                                        int32_t *__restrict__ out)
{
    __shared__ long long sh[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        v[q] = base + q < n ? in[base + q] : 0;
        s += v[q];
    }
    long long total;
    long long e = block_excl_scan_ll(s, sh, total) + sums[blockIdx.x];
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        if (base + q <= n) out[base + q] = (int32_t)e;
        e += v[q];
    }
}

// row of every CSR position
__global__ void __launch_bounds__(256) k_row_of(int64_t m, int64_t nnz, const int32_t *__restrict__ Ap, int32_t *__restrict__ row_of)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int64_t lo = 0, hi = m;                 // last row with Ap[row] <= p
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (Ap[mid] <= p) lo = mid; else hi = mid;
    }
    row_of[p] = (int32_t)lo;
}

// per stored entry (i, k) at CSR position p: first CSC slot of column k with row >= i, and how many follow
__global__ void __launch_bounds__(256)
k_ns_count(int64_t nnz, const int32_t *__restrict__ row_of, const int32_t *__restrict__ Aj, const int32_t *__restrict__ cptr,
           const int32_t *__restrict__ crow, int32_t *__restrict__ dstart, int32_t *__restrict__ cnt)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int32_t i = row_of[p], k = Aj[p];
    int32_t lo = cptr[k], hi = cptr[k + 1];
    const int32_t end = hi;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (crow[mid] < i) lo = mid + 1; else hi = mid;
    }
    dstart[p] = lo;
    cnt[p] = end - lo;
}

__global__ void __launch_bounds__(256)
k_ns_row_terms(int64_t m, const int32_t *__restrict__ Ap, const int32_t *__restrict__ ent_off, int32_t *__restrict__ row_terms,
               int *__restrict__ max_terms)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t t = ent_off[Ap[i + 1]] - ent_off[Ap[i]];
    row_terms[i] = t;
    atomicMax(max_terms, t);
}

struct NsArgs {
    int64_t m;
    const int32_t *Ap, *Aj, *cptr, *crow, *cpos, *dstart, *ent_off;
    int32_t *term_pi, *term_pj, *term_k;       // sorted terms (final)
    int32_t *head_j, *head_t;                  // compacted heads of every row, at the row's term offset
    int32_t *row_nnzc;
    unsigned long long *gkey;                  // global sort workspace (rows above the shared-memory capacity), or null
    int32_t *gval;
    int lo, hi;                                // this launch handles rows with lo < terms <= hi
    int cap;                                   // shared-memory capacity in terms (0: sort in the global workspace)
};

// One CTA per row of the pattern.
__global__ void k_ns_sort(NsArgs a)
{
    extern __shared__ unsigned long long sh_key[];
    __shared__ int sh_warp[33];
    const int64_t i = blockIdx.x;
    const int32_t p0 = a.Ap[i], p1 = a.Ap[i + 1];
    const int32_t toff = a.ent_off[p0];
    const int T = a.ent_off[p1] - toff;
    if (T <= a.lo || T > a.hi) return;
    unsigned long long *key = a.cap ? sh_key : a.gkey + toff;
    int32_t *val = a.cap ? (int32_t *)(sh_key + a.cap) : a.gval + toff;
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- emit: entry p contributes the tail of its column, already ascending in j
    for (int32_t p = p0 + tid; p < p1; p += nt) {
        const int32_t k = a.Aj[p];
        const int32_t d0 = a.dstart[p], d1 = a.cptr[k + 1];
        int32_t o = a.ent_off[p] - toff;
        for (int32_t d = d0; d < d1; ++d, ++o) {
            key[o] = ((unsigned long long)(unsigned)a.crow[d] << 32) | (unsigned)a.cpos[d];
            val[o] = p;
        }
    }
    __syncthreads();
    // ---- bitonic network with ascending comparators only ("flip" + "disperse"): the virtual padding up to the next
    // power of two is +infinity and never has to move, so pairs that reach beyond T are simply skipped
    int P = 1;
    while (P < T) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        const int hk = k >> 1;
        for (int t = tid; t < (P >> 1); t += nt) {
            const int blk = t / hk, off = t - blk * hk;
            const int x = blk * k + off, y = blk * k + k - 1 - off;
            if (y < T) {
                const unsigned long long kx = key[x], ky = key[y];
                if (kx > ky) {
                    key[x] = ky; key[y] = kx;
                    const int32_t vx = val[x]; val[x] = val[y]; val[y] = vx;
                }
            }
        }
        __syncthreads();
        for (int j = hk >> 1; j >= 1; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nt) {
                const int x = 2 * j * (t / j) + (t % j), y = x + j;
                if (y < T) {
                    const unsigned long long kx = key[x], ky = key[y];
                    if (kx > ky) {
                        key[x] = ky; key[y] = kx;
                        const int32_t vx = val[x]; val[x] = val[y]; val[y] = vx;
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- write the sorted terms; compact the first term of every distinct j
    int base = 0;
    const int lane = tid & 31, w = tid >> 5, nw = nt >> 5;
    for (int t0 = 0; t0 < T; t0 += nt) {
        const int t = t0 + tid;
        bool head = false;
        int32_t j = 0;
        if (t < T) {
            const unsigned long long kt = key[t];
            j = (int32_t)(kt >> 32);
            const int32_t pj = (int32_t)(kt & 0xffffffffu);
            a.term_pi[toff + t] = val[t];
            a.term_pj[toff + t] = pj;
            a.term_k[toff + t] = a.Aj[pj];
            head = (t == 0) || ((int32_t)(key[t - 1] >> 32) != j);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) sh_warp[w] = __popc(bal);
        __syncthreads();
        if (w == 0) {
            int s = lane < nw ? sh_warp[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            if (lane < nw) sh_warp[lane] = s;          // inclusive
            if (lane == 31) sh_warp[32] = s;
        }
        __syncthreads();
        if (head) {
            const int h = base + (w > 0 ? sh_warp[w - 1] : 0) + __popc(bal & ((1u << lane) - 1u));
            a.head_j[toff + h] = j;
            a.head_t[toff + h] = toff + t;
        }
        base += sh_warp[32];
        __syncthreads();
    }
    if (tid == 0) a.row_nnzc[i] = base;
}

// one warp per row: heads -> Cj / term_ptr at the row's offset in the pattern
__global__ void __launch_bounds__(256)
k_ns_place(int64_t m, const int32_t *__restrict__ Ap, const int32_t *__restrict__ ent_off, const int32_t *__restrict__ Cp,
           const int32_t *__restrict__ head_j, const int32_t *__restrict__ head_t, int32_t *__restrict__ Cj,
           int32_t *__restrict__ term_ptr, int32_t n_terms)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= m) return;
    const int32_t toff = ent_off[Ap[i]];
    const int32_t c0 = Cp[i], c1 = Cp[i + 1];
    for (int32_t c = c0 + lane; c < c1; c += 32) {
        Cj[c] = head_j[toff + (c - c0)];
        term_ptr[c] = head_t[toff + (c - c0)];
    }
    if (i == m - 1 && lane == 0) term_ptr[c1] = n_terms;
}

struct Scratch {       // scratch released at scope exit: from the stream-ordered pool when the handle uses it (the memory
                       // then serves the large allocations that follow: factor, update matrices), else cudaMalloc / cudaFree
    std::vector<void *> ptrs;
    cudaStream_t st = nullptr;
    bool pooled = false;
    explicit Scratch(cudaStream_t s, bool p) : st(s), pooled(p) {}
    ~Scratch() { for (void *p : ptrs) { if (pooled) cudaFreeAsync(p, st); else cudaFree(p); } }
    template <typename T> cudaError_t get(T **p, size_t count) {
        *p = nullptr;
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        cudaError_t e = pooled ? cudaMallocAsync((void **)p, bytes, st) : cudaMalloc((void **)p, bytes);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

int exclusive_scan(Handle *h, Scratch &sc, int64_t n, const int32_t *d_in, int32_t *d_out, long long **d_total)
{
    const int64_t nt = std::max<int64_t>(1, (n + 1 + SCAN_TILE - 1) / SCAN_TILE);      // tiles cover n + 1 outputs
    long long *sums;
    MIPM_CUDA(h, sc.get(&sums, (size_t)nt + 1));
    k_scan_sums<<<(unsigned)nt, SCAN_TB, 0, h->stream>>>(n, d_in, sums);
    MIPM_CHECK_LAUNCH(h);
    k_scan_tiles<<<1, SCAN_TB, 0, h->stream>>>(nt, sums);
    MIPM_CHECK_LAUNCH(h);
    k_scan_write<<<(unsigned)nt, SCAN_TB, 0, h->stream>>>(n, d_in, sums, d_out);
    MIPM_CHECK_LAUNCH(h);
    *d_total = sums + nt;
    return MIPM_OK;
}

// duplicate columns inside a row show up as equal rows next to each other inside a column of the CSC index
__global__ void __launch_bounds__(256)
k_ns_dup(int64_t nnz, const int32_t *__restrict__ Aj, const int32_t *__restrict__ crow, const int32_t *__restrict__ cpos, int *__restrict__ flag)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d + 1 >= nnz) return;
    if (crow[d] == crow[d + 1] && Aj[cpos[d]] == Aj[cpos[d + 1]]) *flag = 1;
}

// Row runs of the streaming assembly (k_spmv_stream over the term map), computed in parallel: stored entry c starts a
// new block when its first term lies in another ASM_BUCKET-sized bucket of the term array than the first term of c - 1,
// or when c or c - 1 has more than ASM_LONG terms. A block of several entries then holds fewer than ASM_BUCKET + ASM_LONG
// = 2 048 terms (the shared-memory tile of the kernel) and fewer than ASM_BUCKET entries; a long entry stands alone.
constexpr int ASM_BUCKET = 1536, ASM_LONG = 512;
__global__ void __launch_bounds__(256) k_asm_block_flags(int64_t nnzc, const int32_t *__restrict__ term_ptr, int32_t *__restrict__ flag)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nnzc) return;
    int f = 1;
    if (c > 0) {
        const int32_t t0 = term_ptr[c - 1], t1 = term_ptr[c], t2 = term_ptr[c + 1];
        f = (t1 / ASM_BUCKET != t0 / ASM_BUCKET) || (t2 - t1 > ASM_LONG) || (t1 - t0 > ASM_LONG);
    }
    flag[c] = f;
}
__global__ void __launch_bounds__(256)
k_asm_block_starts(int64_t nnzc, const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, int32_t *__restrict__ blk, int32_t nblk)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nnzc && flag[c]) blk[pos[c]] = (int32_t)c;
    if (c == 0) blk[nblk] = (int32_t)nnzc;
}

}  // namespace

// Builds the pattern of tril(A A') and the product-term map on the device. Ap_in == nullptr: the matrix registered with
// mipm_spmv_setup on this handle is used (its CSR / CSC index already lives on the device). Leaves the term map and the
// assembly blocks in the handle and returns the pattern in malloc'd host arrays (index_base added).
int normal_symbolic_device(Handle *h, int64_t m, int64_t n, const int32_t *Ap_in, const int32_t *Aj_in, int index_base,
                           int32_t **Cp_out, int32_t **Cj_out)
{
    *Cp_out = nullptr;
    *Cj_out = nullptr;
    NormalSymbolic &S = h->nsym;
    S = NormalSymbolic();
    const bool tlog = std::getenv("MIPM_ANALYZE_LOG") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto stage = [&](const char *name) {
        if (!tlog) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "normal_symbolic (device): %s %.3f s\n", name, std::chrono::duration<double>(now - t_prev).count());
        t_prev = now;
    };
    if (m < 0 || n < 0) return fail(h, MIPM_ERR_ARG, "negative dimension");
    S.m = m;
    S.n = n;
    cudaStream_t st = h->stream;
    Scratch sc(st, tl_pooled);
    const int32_t *d_Ap, *d_Aj, *d_cptr, *d_crow, *d_cpos;
    int64_t nnz = 0;
    auto finish_empty = [&]() -> int {
        std::vector<int32_t> one(1, 0);
        MIPM_CUDA(h, h->d_term_ptr.upload(one, st));
        MIPM_CUDA(h, h->d_asm_blk.upload(one, st));
        h->asm_nblk = 0;
        MIPM_CUDA(h, h->d_term_pi.alloc(0));
        MIPM_CUDA(h, h->d_term_pj.alloc(0));
        MIPM_CUDA(h, h->d_term_k.alloc(0));
        MIPM_CUDA(h, cudaStreamSynchronize(st));
        *Cp_out = (int32_t *)std::malloc(((size_t)m + 1) * sizeof(int32_t));
        *Cj_out = (int32_t *)std::malloc(sizeof(int32_t));
        if (!*Cp_out || !*Cj_out) return fail(h, MIPM_ERR_ALLOC, "host allocation failed");
        for (int64_t i = 0; i <= m; ++i) (*Cp_out)[i] = index_base;
        return MIPM_OK;
    };
    if (!Ap_in) {
        if (!h->has_spmv || h->sp_m != m || h->sp_n != n)
            return fail(h, MIPM_ERR_STATE, "no row pointer given and no matrix of this shape registered with mipm_spmv_setup");
        nnz = h->sp_nnz;
        S.nnz_a = nnz;
        if (nnz == 0 || m == 0) return finish_empty();
        d_Ap = h->d_sp_rowptr.p; d_Aj = h->d_sp_col.p; d_cptr = h->d_sp_colptr.p; d_crow = h->d_sp_row.p; d_cpos = h->d_sp_pos.p;
        int *d_dup;
        MIPM_CUDA(h, sc.get(&d_dup, 1));
        MIPM_CUDA(h, cudaMemsetAsync(d_dup, 0, sizeof(int), st));
        k_ns_dup<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, d_Aj, d_crow, d_cpos, d_dup);
        MIPM_CHECK_LAUNCH(h);
        int dup = 0;
        MIPM_CUDA(h, cudaMemcpyAsync(&dup, d_dup, sizeof(int), cudaMemcpyDeviceToHost, st));
        MIPM_CUDA(h, cudaStreamSynchronize(st));
        if (dup) return fail(h, MIPM_ERR_DUPLICATE, "duplicate column inside a row of A");
        stage("registered SpMV index reused, duplicate check");
    } else {
        uvector<int32_t> Ap((size_t)m + 1);
        for (int64_t i = 0; i <= m; ++i) Ap[(size_t)i] = Ap_in[i] - index_base;
        if (m > 0 && Ap[0] != 0) return fail(h, MIPM_ERR_ARG, "row pointer does not start at index_base");
        for (int64_t i = 0; i < m; ++i)
            if (Ap[(size_t)i + 1] < Ap[(size_t)i]) return fail(h, MIPM_ERR_ARG, "row pointer not monotone");
        nnz = m > 0 ? Ap[(size_t)m] : 0;
        S.nnz_a = nnz;
        if (nnz == 0 || m == 0) return finish_empty();
        // host: 0-based columns, CSC index with CSR positions (stable bucket sort by column on the host threads: rows
        // ascend inside a column), duplicate check
        uvector<int32_t> Aj((size_t)nnz), cptr((size_t)n + 1), crow((size_t)nnz), cpos((size_t)nnz);
        if (!stable_bucket_parallel(n, nnz, Aj_in, index_base, cptr.data(), [&](int64_t d, int64_t p) { cpos[(size_t)d] = (int32_t)p; }))
            return fail(h, MIPM_ERR_ARG, "column index out of range");
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), nnz / 262144));
        run_host_threads(T, [&](int t) {
            for (int64_t p = nnz * t / T; p < nnz * (t + 1) / T; ++p) Aj[(size_t)p] = Aj_in[p] - index_base;
            for (int64_t d = nnz * t / T; d < nnz * (t + 1) / T; ++d)
                crow[(size_t)d] = (int32_t)(std::upper_bound(Ap.begin(), Ap.end(), cpos[(size_t)d]) - Ap.begin()) - 1;
        });
        std::vector<int> dup((size_t)T, 0);
        run_host_threads(T, [&](int t) {
            for (int64_t k = n * t / T; k < n * (t + 1) / T; ++k)
                for (int32_t d = cptr[(size_t)k] + 1; d < cptr[(size_t)k + 1]; ++d)
                    if (crow[(size_t)d] == crow[(size_t)d - 1]) { dup[(size_t)t] = 1; return; }
        });
        for (int v : dup) if (v) return fail(h, MIPM_ERR_DUPLICATE, "duplicate column inside a row of A");
        stage("host CSC index");
        int32_t *u_Ap, *u_Aj, *u_cptr, *u_crow, *u_cpos;
        MIPM_CUDA(h, sc.get(&u_Ap, (size_t)m + 1));
        MIPM_CUDA(h, sc.get(&u_Aj, (size_t)nnz));
        MIPM_CUDA(h, sc.get(&u_cptr, (size_t)n + 1));
        MIPM_CUDA(h, sc.get(&u_crow, (size_t)nnz));
        MIPM_CUDA(h, sc.get(&u_cpos, (size_t)nnz));
        MIPM_CUDA(h, cudaMemcpyAsync(u_Ap, Ap.data(), ((size_t)m + 1) * 4, cudaMemcpyHostToDevice, st));
        MIPM_CUDA(h, cudaMemcpyAsync(u_Aj, Aj.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        MIPM_CUDA(h, cudaMemcpyAsync(u_cptr, cptr.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, st));
        MIPM_CUDA(h, cudaMemcpyAsync(u_crow, crow.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        MIPM_CUDA(h, cudaMemcpyAsync(u_cpos, cpos.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        MIPM_CUDA(h, cudaStreamSynchronize(st));        // the host vectors go out of scope
        d_Ap = u_Ap; d_Aj = u_Aj; d_cptr = u_cptr; d_crow = u_crow; d_cpos = u_cpos;
        stage("uploads");
    }
    int32_t *d_row_of, *d_dstart, *d_cnt, *d_ent_off, *d_row_terms, *d_row_nnzc, *d_Cp;
    int *d_max;
    MIPM_CUDA(h, sc.get(&d_row_of, (size_t)nnz));
    MIPM_CUDA(h, sc.get(&d_dstart, (size_t)nnz));
    MIPM_CUDA(h, sc.get(&d_cnt, (size_t)nnz));
    MIPM_CUDA(h, sc.get(&d_ent_off, (size_t)nnz + 1));
    MIPM_CUDA(h, sc.get(&d_row_terms, (size_t)m));
    MIPM_CUDA(h, sc.get(&d_row_nnzc, (size_t)m));
    MIPM_CUDA(h, sc.get(&d_Cp, (size_t)m + 1));
    MIPM_CUDA(h, sc.get(&d_max, 1));
    MIPM_CUDA(h, cudaMemsetAsync(d_max, 0, sizeof(int), st));
    const unsigned gp = (unsigned)((nnz + 255) / 256), gm = (unsigned)((m + 255) / 256);
    k_row_of<<<gp, 256, 0, st>>>(m, nnz, d_Ap, d_row_of);
    MIPM_CHECK_LAUNCH(h);
    k_ns_count<<<gp, 256, 0, st>>>(nnz, d_row_of, d_Aj, d_cptr, d_crow, d_dstart, d_cnt);
    MIPM_CHECK_LAUNCH(h);
    long long *d_total;
    int rc = exclusive_scan(h, sc, nnz, d_cnt, d_ent_off, &d_total);
    if (rc != MIPM_OK) return rc;
    k_ns_row_terms<<<gm, 256, 0, st>>>(m, d_Ap, d_ent_off, d_row_terms, d_max);
    MIPM_CHECK_LAUNCH(h);
    long long T = 0;
    int max_terms = 0;
    MIPM_CUDA(h, cudaMemcpyAsync(&T, d_total, sizeof(T), cudaMemcpyDeviceToHost, st));
    MIPM_CUDA(h, cudaMemcpyAsync(&max_terms, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    MIPM_CUDA(h, cudaStreamSynchronize(st));
    if (T >= (long long)INT32_MAX) return fail(h, MIPM_ERR_ARG, "too many product terms for 32-bit segment pointers");
    stage("counts + scan (sync)");
    S.n_terms = T;
    MIPM_CUDA(h, h->d_term_pi.alloc((size_t)T));
    MIPM_CUDA(h, h->d_term_pj.alloc((size_t)T));
    MIPM_CUDA(h, h->d_term_k.alloc((size_t)T));
    // the compacted heads (2 x 4 bytes per term) live in the buffer of the term weights (8 bytes per term), which
    // mipm_normal_set_jacobian fills later
    MIPM_CUDA(h, h->d_term_w.alloc((size_t)std::max<long long>(T, 1)));
    int32_t *d_head_j = reinterpret_cast<int32_t *>(h->d_term_w.p), *d_head_t = d_head_j + T;
    // ---- sort classes by the number of terms of a row: 128 threads / 24 KB, 512 threads / 192 KB, global workspace
    constexpr int CAP_S = 2048, CAP_L = 16384;
    NsArgs a;
    a.m = m; a.Ap = d_Ap; a.Aj = d_Aj; a.cptr = d_cptr; a.crow = d_crow; a.cpos = d_cpos; a.dstart = d_dstart; a.ent_off = d_ent_off;
    a.term_pi = h->d_term_pi.p; a.term_pj = h->d_term_pj.p; a.term_k = h->d_term_k.p;
    a.head_j = d_head_j; a.head_t = d_head_t; a.row_nnzc = d_row_nnzc;
    a.gkey = nullptr; a.gval = nullptr;
    MIPM_CUDA(h, cudaMemsetAsync(d_row_nnzc, 0, (size_t)m * 4, st));            // rows without terms
    {
        a.lo = 0; a.hi = CAP_S; a.cap = CAP_S;
        k_ns_sort<<<(unsigned)m, 128, (size_t)CAP_S * 12, st>>>(a);
        MIPM_CHECK_LAUNCH(h);
    }
    if (max_terms > CAP_S) {
        MIPM_CUDA(h, cudaFuncSetAttribute(k_ns_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, CAP_L * 12));
        a.lo = CAP_S; a.hi = CAP_L; a.cap = CAP_L;
        k_ns_sort<<<(unsigned)m, 512, (size_t)CAP_L * 12, st>>>(a);
        MIPM_CHECK_LAUNCH(h);
    }
    if (max_terms > CAP_L) {
        MIPM_CUDA(h, sc.get(&a.gkey, (size_t)T));
        MIPM_CUDA(h, sc.get(&a.gval, (size_t)T));
        a.lo = CAP_L; a.hi = INT32_MAX; a.cap = 0;
        k_ns_sort<<<(unsigned)m, 1024, 0, st>>>(a);
        MIPM_CHECK_LAUNCH(h);
    }
    long long *d_total_c;
    rc = exclusive_scan(h, sc, m, d_row_nnzc, d_Cp, &d_total_c);
    if (rc != MIPM_OK) return rc;
    long long nnzc = 0;
    MIPM_CUDA(h, cudaMemcpyAsync(&nnzc, d_total_c, sizeof(nnzc), cudaMemcpyDeviceToHost, st));
    MIPM_CUDA(h, cudaStreamSynchronize(st));
    S.nnz_c = nnzc;
    stage("term arrays allocated, sort + compaction (sync)");
    int32_t *d_Cj;
    MIPM_CUDA(h, sc.get(&d_Cj, (size_t)nnzc));
    MIPM_CUDA(h, h->d_term_ptr.alloc((size_t)nnzc + 1));
    k_ns_place<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(m, d_Ap, d_ent_off, d_Cp, d_head_j, d_head_t, d_Cj, h->d_term_ptr.p, (int32_t)T);
    MIPM_CHECK_LAUNCH(h);
    // ---- the pattern goes straight into the arrays the caller receives
    *Cp_out = (int32_t *)std::malloc(((size_t)m + 1) * sizeof(int32_t));
    *Cj_out = (int32_t *)std::malloc(std::max<size_t>((size_t)nnzc, 1) * sizeof(int32_t));
    if (!*Cp_out || !*Cj_out) return fail(h, MIPM_ERR_ALLOC, "host allocation failed");
    MIPM_CUDA(h, cudaMemcpyAsync(*Cp_out, d_Cp, ((size_t)m + 1) * 4, cudaMemcpyDeviceToHost, st));
    MIPM_CUDA(h, cudaMemcpyAsync(*Cj_out, d_Cj, (size_t)nnzc * 4, cudaMemcpyDeviceToHost, st));
    // ---- row runs of the streaming assembly
    int64_t nblk = 0;
    if (nnzc > 0) {
        int32_t *d_flag, *d_pos;
        MIPM_CUDA(h, sc.get(&d_flag, (size_t)nnzc));
        MIPM_CUDA(h, sc.get(&d_pos, (size_t)nnzc + 1));
        const unsigned gc = (unsigned)((nnzc + 255) / 256);
        k_asm_block_flags<<<gc, 256, 0, st>>>(nnzc, h->d_term_ptr.p, d_flag);
        MIPM_CHECK_LAUNCH(h);
        long long *d_nblk;
        rc = exclusive_scan(h, sc, nnzc, d_flag, d_pos, &d_nblk);
        if (rc != MIPM_OK) return rc;
        long long nb = 0;
        MIPM_CUDA(h, cudaMemcpyAsync(&nb, d_nblk, sizeof(nb), cudaMemcpyDeviceToHost, st));
        MIPM_CUDA(h, cudaStreamSynchronize(st));
        nblk = nb;
        MIPM_CUDA(h, h->d_asm_blk.alloc((size_t)nblk + 1));
        k_asm_block_starts<<<gc, 256, 0, st>>>(nnzc, d_flag, d_pos, h->d_asm_blk.p, (int32_t)nblk);
        MIPM_CHECK_LAUNCH(h);
    } else {
        std::vector<int32_t> one(1, 0);
        MIPM_CUDA(h, h->d_asm_blk.upload(one, st));
    }
    h->asm_nblk = nblk;
    MIPM_CUDA(h, cudaStreamSynchronize(st));
    if (index_base != 0) {
        const int Tt = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), nnzc / 262144));
        int32_t *cp = *Cp_out, *cj = *Cj_out;
        run_host_threads(Tt, [&](int t) {
            for (int64_t i = (m + 1) * t / Tt; i < (m + 1) * (t + 1) / Tt; ++i) cp[i] += index_base;
            for (int64_t c = nnzc * t / Tt; c < nnzc * (t + 1) / Tt; ++c) cj[c] += index_base;
        });
    }
    stage("placement, assembly blocks, pattern to the host");
    return MIPM_OK;
}

}  // namespace mipm
