// C-ABI entry points: lifecycle, host symbolic builders, KKT assembly kernels, SpMV.
// Each entry cites in include/madipm_b200.h the reference interface it replaces.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"

#include <mutex>

using namespace mipm;

namespace {

template <typename T>
T *host_copy(const std::vector<T> &v, int add) {
    T *p = (T *)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (!p) return nullptr;
    for (size_t i = 0; i < v.size(); ++i) p[i] = (T)(v[i] + add);
    return p;
}

// ---------------------------------------------------------------- normal-equations assembly
__global__ void k_term_weights(int64_t T, const int32_t *__restrict__ pi, const int32_t *__restrict__ pj,
                               const double *__restrict__ Ax, double *__restrict__ w)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) w[t] = Ax[pi[t]] * Ax[pj[t]];
}

__global__ void k_recip(int64_t n, const double *__restrict__ x, double *__restrict__ y)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __ddiv_rn(1.0, x[i]);   // D .= 1.0 ./ pr_diag  (normalkkt.jl:191)
}

// The fast assembly Cx[c] = sum_t w[t] * D[k[t]] over the term segment of every stored entry c of tril(A D A') is a
// sparse matrix-vector product (rows = stored entries, nonzeros = product terms, x = D) and runs through
// k_spmv_stream below: coalesced term loads, per-entry sums in storage order (deterministic).

// Same, reproducing the reference CPU loop's operation order (src/utils.jl:288-301):
// buffer[k] = A[i,k]*D[k]; Cx[c] += buffer[k]*A[j,k] in row-j order, mul and add unfused.
__global__ void __launch_bounds__(256)
k_normal_assemble_exact(int64_t nnzC, const int32_t *__restrict__ term_ptr, const int32_t *__restrict__ pi,
                        const int32_t *__restrict__ pj, const int32_t *__restrict__ term_k,
                        const double *__restrict__ Ax, const double *__restrict__ D, double *__restrict__ Cx)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nnzC) return;
    int32_t t0 = term_ptr[c], t1 = term_ptr[c + 1];
    double acc = 0.0;
    for (int32_t t = t0; t < t1; ++t) {
        double b = __dmul_rn(Ax[pi[t]], D[term_k[t]]);
        acc = __dadd_rn(acc, __dmul_rn(b, Ax[pj[t]]));
    }
    Cx[c] = acc;
}

// ---------------------------------------------------------------- K2 transfer (gather form)
__global__ void __launch_bounds__(256)
k_k2_transfer(int64_t nslots, const int64_t *__restrict__ slot_ptr, const int64_t *__restrict__ slot_src,
              const double *__restrict__ V, double *__restrict__ nz)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nslots) return;
    double acc = 0.0;
    for (int64_t s = slot_ptr[c]; s < slot_ptr[c + 1]; ++s) acc = __dadd_rn(acc, V[slot_src[s]]);
    nz[c] = acc;
}

// ---------------------------------------------------------------- SpMV
// y = alpha * A x + beta * y, CSR, G lanes per row.
template <int G>
__global__ void __launch_bounds__(256)
k_spmv_csr(int64_t m, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
           const double *__restrict__ val, const double *__restrict__ x, double alpha, double beta,
           double *__restrict__ y)
{
    // G lanes per row (G = power of two chosen from the average row length at setup)
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int lane = threadIdx.x & (G - 1);
    double acc = 0.0;
    if (row < m)
        for (int32_t p = rowptr[row] + lane; p < rowptr[row + 1]; p += G) acc = fma(val[p], __ldg(x + col[p]), acc);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < m && lane == 0) y[row] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * y[row];
}

// Streaming SpMV ("CSR-stream", Greathouse & Daga): a block owns a run of consecutive rows with at most SPMV_CAP
// nonzeros. Step 1: all threads form the products val * x[idx] with coalesced loads (8 independent loads in flight
// per thread) into shared memory; step 2: one thread per row adds its products in storage order (deterministic).
// A row longer than SPMV_CAP is a block of its own and is reduced CTA-wide. The same kernel serves A x (rows of the
// CSR index) and A' y (columns of the CSC index; POS = values reached through the position map when no
// column-ordered copy of the values has been cached).
constexpr int SPMV_CAP = 2048;
struct SpmvJob {
    const int32_t *blk, *ptr, *idx, *pos;      // row runs, segment pointers, gather indices, optional value positions
    const double *val, *x;
    double alpha, beta;
    double *y;
    int lanes;                                  // threads that share a row in step 2 (power of two <= 32)
    int nblk;
};

template <bool POS>
__device__ __forceinline__ void spmv_stream_block(const SpmvJob &j, int b, double *prod, double *red)
{
    const int r0 = j.blk[b], r1 = j.blk[b + 1];
    const int p0 = j.ptr[r0], p1 = j.ptr[r1];
    const int tid = threadIdx.x;
    if (p1 - p0 > SPMV_CAP) {               // one long row
        double acc = 0.0;
        for (int p = p0 + tid; p < p1; p += 256) acc = fma(POS ? __ldg(j.val + j.pos[p]) : j.val[p], __ldg(j.x + j.idx[p]), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            j.y[r0] = (j.beta == 0.0) ? j.alpha * t : j.alpha * t + j.beta * j.y[r0];
        }
        return;
    }
    // step 1: products, eight independent loads in flight per thread
    {
        const int np = p1 - p0;
        int q = tid;
        for (; q + 768 < np; q += 1024) {
            double v[4];
            int c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = POS ? __ldg(j.val + j.pos[p0 + q + 256 * u]) : j.val[p0 + q + 256 * u];
                c[u] = j.idx[p0 + q + 256 * u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) prod[q + 256 * u] = v[u] * __ldg(j.x + c[u]);
        }
        for (; q < np; q += 256) prod[q] = (POS ? __ldg(j.val + j.pos[p0 + q]) : j.val[p0 + q]) * __ldg(j.x + j.idx[p0 + q]);
    }
    __syncthreads();
    // step 2: `lanes` threads per row add strided partial sums and combine them by shuffles (fixed order: deterministic)
    const int g = j.lanes, per = 256 / g;
    const int sub = tid / g, l = tid & (g - 1);
    for (int base = r0; base < r1; base += per) {
        const int row = base + sub;
        double acc = 0.0;
        if (row < r1) {
            const int a = j.ptr[row] - p0, e = j.ptr[row + 1] - p0;
            for (int q = a + l; q < e; q += g) acc += prod[q];
        }
        for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (row < r1 && l == 0) j.y[row] = (j.beta == 0.0) ? j.alpha * acc : j.alpha * acc + j.beta * j.y[row];
    }
}

template <bool POS>
__global__ void __launch_bounds__(256) k_spmv_stream(SpmvJob j)
{
    __shared__ double prod[SPMV_CAP];
    __shared__ double red[8];
    spmv_stream_block<POS>(j, blockIdx.x, prod, red);
}

// Two independent products in one launch (A x and A' y of the same iteration phase): blocks [0, a.nblk) run job a,
// the rest job b.
__global__ void __launch_bounds__(256) k_spmv_stream_pair(SpmvJob a, SpmvJob b)
{
    __shared__ double prod[SPMV_CAP];
    __shared__ double red[8];
    if ((int)blockIdx.x < a.nblk) spmv_stream_block<false>(a, blockIdx.x, prod, red);
    else spmv_stream_block<false>(b, blockIdx.x - a.nblk, prod, red);
}

__global__ void __launch_bounds__(256)
k_gather32(int64_t n, const double *__restrict__ src, const int32_t *__restrict__ map, double *__restrict__ dst)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[map[i]];
}

// Row runs for k_spmv_stream: consecutive rows, at most SPMV_CAP nonzeros and 256 rows per block.
constexpr int ASM_ROWS = 1024;
template <typename Vec>
static std::vector<int32_t> spmv_blocks(const Vec &ptr, int max_rows = 256)
{
    const int64_t nrows = (int64_t)ptr.size() - 1;
    std::vector<int32_t> blk;
    blk.push_back(0);
    int64_t r = 0;
    while (r < nrows) {
        int64_t e = r + 1;      // at least one row (a long row stands alone)
        while (e < nrows && e - r < max_rows && ptr[(size_t)e + 1] - ptr[(size_t)r] <= SPMV_CAP) ++e;
        blk.push_back((int32_t)e);
        r = e;
    }
    return blk;
}

// lanes per row / column: the power of two nearest above the average length, within [2, 32]
inline int spmv_group(int64_t nnz, int64_t rows)
{
    double avg = rows > 0 ? (double)nnz / (double)rows : 1.0;
    int g = 2;
    while (g < 32 && g < avg) g *= 2;
    return g;
}

// threads per row in step 2 of the streaming SpMV: rows of a few entries are summed by one thread
inline int stream_lanes(int64_t nnz, int64_t rows)
{
    const double avg = rows > 0 ? (double)nnz / (double)rows : 1.0;
    int g = 1;
    while (g < 32 && 8.0 * g <= avg) g *= 2;
    return g;
}

inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)std::max<int64_t>(1, (n + per_block - 1) / per_block); }

}  // namespace

namespace mipm {

int device_info(int device, DeviceInfo &out)
{
    static std::mutex mu;
    static std::vector<std::pair<int, DeviceInfo>> cache;
    std::lock_guard<std::mutex> lk(mu);
    for (auto &e : cache) if (e.first == device) { out = e.second; return MIPM_OK; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MIPM_ERR_CUDA;
    out.sm_count = prop.multiProcessorCount;
    out.cooperative = prop.cooperativeLaunch;
    cache.push_back({device, out});
    return MIPM_OK;
}

static std::mutex g_pin_mu;
static std::vector<double *> g_pin_free;

double *pinned_scalars_acquire()
{
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) { double *p = g_pin_free.back(); g_pin_free.pop_back(); return p; }
    }
    double *p = nullptr;
    if (cudaMallocHost((void **)&p, 64 * sizeof(double)) != cudaSuccess) return nullptr;
    return p;
}

void pinned_scalars_release(double *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back(p);
}

}  // namespace mipm

extern "C" {

int mipm_version(void) { return 100; }

int mipm_create(mipm_handle *out, int device, void *stream)
{
    if (!out) return MIPM_ERR_ARG;
    *out = nullptr;
    Handle *h = new (std::nothrow) Handle();
    if (!h) return MIPM_ERR_ALLOC;
    h->device = device;
    h->stream = (cudaStream_t)stream;
    if (device < 0) {
        // analysis-only handle: host symbolic entry points work, every device entry fails loudly
        h->host_only = true;
        *out = (mipm_handle)h;
        return MIPM_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device >= ndev) {
        delete h;
        return MIPM_ERR_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return MIPM_ERR_CUDA; }
    DeviceInfo prop;
    if (device_info(device, prop) != MIPM_OK) { delete h; return MIPM_ERR_CUDA; }
    {
        // stream-ordered allocations for everything this handle owns; keep freed blocks in the pool instead of returning
        // them to the driver at every synchronisation
        int pools = 0;
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && pools &&
            cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && !std::getenv("MIPM_NO_POOL")) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            h->pool_ok = true;
        }
        (void)cudaGetLastError();
        use_handle(h);
    }
    h->red_blocks = prop.sm_count * 8;      // every block of a reduction kernel resident at once (8 x 256 threads / SM)
    if (h->d_partials.alloc((size_t)h->red_blocks * 16) != cudaSuccess ||
        h->d_scal.alloc(64) != cudaSuccess || h->d_counter.alloc(4) != cudaSuccess ||
        (h->h_scal = pinned_scalars_acquire()) == nullptr ||
        h->d_info.alloc(4) != cudaSuccess) {
        delete h;
        return MIPM_ERR_ALLOC;
    }
    cudaMemsetAsync(h->d_counter.p, 0, 4 * sizeof(unsigned int), h->stream);
    cudaMemsetAsync(h->d_info.p, 0, 4 * sizeof(int), h->stream);
    *out = (mipm_handle)h;
    return MIPM_OK;
}

int mipm_destroy(mipm_handle hh)
{
    Handle *h = (Handle *)hh;
    if (!h) return MIPM_ERR_ARG;
    if (!h->host_only) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        if (h->side) {
            cudaStreamSynchronize(h->side);
            cudaEventDestroy(h->ev_factor_done);
            cudaEventDestroy(h->ev_u_zero);
            cudaStreamDestroy(h->side);
        }
        pinned_scalars_release(h->h_scal);
    }
    delete h;
    return MIPM_OK;
}

const char *mipm_last_error(mipm_handle hh)
{
    Handle *h = (Handle *)hh;
    return h ? h->err.c_str() : "null handle";
}

void mipm_free(void *p) { std::free(p); }

int64_t mipm_launch_count(mipm_handle hh)
{
    Handle *h = (Handle *)hh;
    return h ? h->launches : -1;
}

int mipm_coo_to_csr(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *Ai, const int32_t *Aj,
                    int index_base, int32_t *Bp, int32_t *Bj, int64_t *Bmap)
{
    if (n_rows < 0 || n_cols < 0 || nnz < 0 || (nnz > 0 && (!Ai || !Aj)) || !Bp || (nnz > 0 && (!Bj || !Bmap)))
        return MIPM_ERR_ARG;
    if (nnz >= (int64_t)INT32_MAX) return MIPM_ERR_ARG;
    for (int64_t k = 0; k < nnz; ++k)
        if (Aj[k] < index_base || Aj[k] - index_base >= n_cols) return MIPM_ERR_ARG;
    // stable counting sort by row (the reference's loop order, src/utils.jl:158-207), rows dealt to the host threads
    if (!stable_bucket_parallel(n_rows, nnz, Ai, index_base, Bp, [&](int64_t d, int64_t k) {
            Bj[d] = Aj[k];
            Bmap[d] = k + index_base;
        }))
        return MIPM_ERR_ARG;
    for (int64_t i = 0; i <= n_rows; ++i) Bp[i] += index_base;
    return MIPM_OK;
}

int mipm_normal_symbolic(mipm_handle hh, int64_t m, int64_t n, const int32_t *Ap, const int32_t *Aj,
                         int index_base, int32_t **Cp, int32_t **Cj, int64_t *nnzC)
{
    Handle *h = (Handle *)hh;
    if (h && !h->host_only) use_handle(h);
    if (!h || !Cp || !Cj || !nnzC) return fail(h, MIPM_ERR_ARG, "null argument");
    const bool device_build = !h->host_only && !std::getenv("MIPM_HOST_SYMBOLIC");
    // Ap == NULL: the matrix registered with mipm_spmv_setup on this handle (device build only)
    if (!Ap && !(device_build && h->has_spmv)) return fail(h, MIPM_ERR_ARG, "null row pointer (and no matrix registered with mipm_spmv_setup on a device handle)");
    if (Ap && !Aj && m > 0 && Ap[m] != index_base) return fail(h, MIPM_ERR_ARG, "null argument");
    // Handles that own a GPU build the pattern and the term map on the device (normal_device.cu); analysis-only handles
    // and MIPM_HOST_SYMBOLIC=1 take the host sweep. Both produce the same arrays.
    if (device_build) {
        int rc = normal_symbolic_device(h, m, n, Ap, Aj, index_base, Cp, Cj);
        if (rc != MIPM_OK) {
            std::free(*Cp); std::free(*Cj);
            *Cp = *Cj = nullptr;
            return rc;
        }
        *nnzC = h->nsym.nnz_c;
        h->has_normal = true;
        h->has_jac = false;
        MIPM_CUDA(h, h->d_D.alloc((size_t)n));          // (d_term_w was allocated by the builder, which used it as scratch)
        return MIPM_OK;
    }
    std::string e = normal_symbolic_host(m, n, Ap, Aj, index_base, h->nsym);
    if (!e.empty()) return fail(h, e.find("duplicate") != std::string::npos ? MIPM_ERR_DUPLICATE : MIPM_ERR_ARG, e);
    *Cp = host_copy(h->nsym.Cp, index_base);
    *Cj = host_copy(h->nsym.Cj, index_base);
    *nnzC = h->nsym.nnz_c;
    if (!*Cp || !*Cj) return fail(h, MIPM_ERR_ALLOC, "host allocation failed");
    h->has_normal = true;
    h->has_jac = false;
    if (!h->host_only) {
        MIPM_CUDA(h, cudaSetDevice(h->device));
        MIPM_CUDA(h, h->d_term_ptr.upload(h->nsym.term_ptr, h->stream));
        MIPM_CUDA(h, h->d_term_pi.upload(h->nsym.term_pi, h->stream));
        MIPM_CUDA(h, h->d_term_pj.upload(h->nsym.term_pj, h->stream));
        MIPM_CUDA(h, h->d_term_k.upload(h->nsym.term_k, h->stream));
        MIPM_CUDA(h, h->d_term_w.alloc((size_t)h->nsym.n_terms));
        {
            std::vector<int32_t> blk = spmv_blocks(h->nsym.term_ptr, ASM_ROWS);       // runs of stored entries, <= 2048 terms each
            h->asm_nblk = (int64_t)blk.size() - 1;
            MIPM_CUDA(h, h->d_asm_blk.upload(blk, h->stream));
        }
        MIPM_CUDA(h, h->d_D.alloc((size_t)n));
        MIPM_CUDA(h, cudaStreamSynchronize(h->stream));   // host vectors may be reused
    }
    return MIPM_OK;
}

int mipm_normal_set_jacobian(mipm_handle hh, const double *d_ATx)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_normal) return fail(h, MIPM_ERR_STATE, "mipm_normal_symbolic has not been called");
    if (!d_ATx && h->nsym.nnz_a > 0) return fail(h, MIPM_ERR_ARG, "null Jacobian values");
    h->d_ATx = d_ATx;
    int64_t T = h->nsym.n_terms;
    if (T > 0) {
        k_term_weights<<<grid_for(T, 256), 256, 0, h->stream>>>(T, h->d_term_pi.p, h->d_term_pj.p, d_ATx, h->d_term_w.p);
        MIPM_CHECK_LAUNCH(h);
    }
    h->has_jac = true;
    return MIPM_OK;
}

int mipm_normal_assemble(mipm_handle hh, const double *d_pr_diag, double *d_Cx, int exact_order)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_jac) return fail(h, MIPM_ERR_STATE, "mipm_normal_set_jacobian has not been called");
    if (!d_pr_diag || !d_Cx) return fail(h, MIPM_ERR_ARG, "null argument");
    const NormalSymbolic &S = h->nsym;
    if (S.n > 0) {
        k_recip<<<grid_for(S.n, 256), 256, 0, h->stream>>>(S.n, d_pr_diag, h->d_D.p);
        MIPM_CHECK_LAUNCH(h);
    }
    if (S.nnz_c > 0) {
        if (exact_order)
            k_normal_assemble_exact<<<grid_for(S.nnz_c, 256), 256, 0, h->stream>>>(
                S.nnz_c, h->d_term_ptr.p, h->d_term_pi.p, h->d_term_pj.p, h->d_term_k.p, h->d_ATx, h->d_D.p, d_Cx);
        else {
            SpmvJob j{h->d_asm_blk.p, h->d_term_ptr.p, h->d_term_k.p, nullptr, h->d_term_w.p, h->d_D.p, 1.0, 0.0, d_Cx, 1, (int)h->asm_nblk};
            k_spmv_stream<false><<<(unsigned)h->asm_nblk, 256, 0, h->stream>>>(j);
        }
        MIPM_CHECK_LAUNCH(h);
    }
    return MIPM_OK;
}

int mipm_k2_symbolic(mipm_handle hh, int64_t dim, int64_t nnz_coo, const int32_t *I, const int32_t *J,
                     int index_base, int32_t **colptr, int32_t **rowval, int64_t **map, int64_t *nnz_csc)
{
    Handle *h = (Handle *)hh;
    if (h && !h->host_only) use_handle(h);
    if (!h || (nnz_coo > 0 && (!I || !J)) || !colptr || !rowval || !map || !nnz_csc) return fail(h, MIPM_ERR_ARG, "null argument");
    std::string e = k2_symbolic_host(dim, nnz_coo, I, J, index_base, h->ksym);
    if (!e.empty()) return fail(h, MIPM_ERR_ARG, e);
    *colptr = host_copy(h->ksym.colptr, index_base);
    *rowval = host_copy(h->ksym.rowval, index_base);
    *map = host_copy(h->ksym.map, index_base);
    *nnz_csc = h->ksym.nnz_csc;
    if (!*colptr || !*rowval || !*map) return fail(h, MIPM_ERR_ALLOC, "host allocation failed");
    h->has_k2 = true;
    if (!h->host_only) {
        MIPM_CUDA(h, cudaSetDevice(h->device));
        MIPM_CUDA(h, h->d_slot_ptr.upload(h->ksym.slot_ptr, h->stream));
        MIPM_CUDA(h, h->d_slot_src.upload(h->ksym.slot_src, h->stream));
        MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return MIPM_OK;
}

int mipm_k2_transfer(mipm_handle hh, const double *d_V, double *d_nz)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_k2) return fail(h, MIPM_ERR_STATE, "mipm_k2_symbolic has not been called");
    if (!d_V || !d_nz) return fail(h, MIPM_ERR_ARG, "null argument");
    if (h->ksym.nnz_csc > 0) {
        k_k2_transfer<<<grid_for(h->ksym.nnz_csc, 256), 256, 0, h->stream>>>(h->ksym.nnz_csc, h->d_slot_ptr.p,
                                                                          h->d_slot_src.p, d_V, d_nz);
        MIPM_CHECK_LAUNCH(h);
    }
    return MIPM_OK;
}

int mipm_spmv_setup(mipm_handle hh, int64_t m, int64_t n, const int32_t *Ap, const int32_t *Aj, int index_base)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (m < 0 || n < 0 || !Ap) return fail(h, MIPM_ERR_ARG, "bad argument");
    int64_t nnz = Ap[m] - index_base;
    if (nnz < 0 || (nnz > 0 && !Aj)) return fail(h, MIPM_ERR_ARG, "bad row pointer");
    uvector<int32_t> rp((size_t)m + 1), cj((size_t)nnz), cp((size_t)n + 1), ri((size_t)nnz), pos((size_t)nnz);
    for (int64_t i = 0; i <= m; ++i) rp[(size_t)i] = Ap[i] - index_base;
    if (m > 0 && rp[0] != 0) return fail(h, MIPM_ERR_ARG, "row pointer does not start at index_base");
    for (int64_t i = 0; i < m; ++i)
        if (rp[(size_t)i + 1] < rp[(size_t)i]) return fail(h, MIPM_ERR_ARG, "row pointer not monotone");
    // CSC index with CSR positions: stable bucket sort of the positions by column (rows ascend inside a column)
    if (!stable_bucket_parallel(n, nnz, Aj, index_base, cp.data(), [&](int64_t d, int64_t p) { pos[(size_t)d] = (int32_t)p; }))
        return fail(h, MIPM_ERR_ARG, "column index out of range");
    {
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), nnz / 262144));
        run_host_threads(T, [&](int t) {
            for (int64_t p = nnz * t / T; p < nnz * (t + 1) / T; ++p) cj[(size_t)p] = Aj[p] - index_base;
            for (int64_t d = nnz * t / T; d < nnz * (t + 1) / T; ++d) {
                const int32_t p = pos[(size_t)d];                      // row of CSR position p
                ri[(size_t)d] = (int32_t)(std::upper_bound(rp.begin(), rp.end(), p) - rp.begin()) - 1;
            }
        });
    }
    MIPM_CUDA(h, cudaSetDevice(h->device));
    MIPM_CUDA(h, h->d_sp_rowptr.upload(rp, h->stream));
    MIPM_CUDA(h, h->d_sp_col.upload(cj, h->stream));
    MIPM_CUDA(h, h->d_sp_colptr.upload(cp, h->stream));
    MIPM_CUDA(h, h->d_sp_row.upload(ri, h->stream));
    MIPM_CUDA(h, h->d_sp_pos.upload(pos, h->stream));
    {
        std::vector<int32_t> br = spmv_blocks(rp), bc = spmv_blocks(cp);
        h->sp_nblk_rows = (int64_t)br.size() - 1;
        h->sp_nblk_cols = (int64_t)bc.size() - 1;
        MIPM_CUDA(h, h->d_sp_blk_rows.upload(br, h->stream));
        MIPM_CUDA(h, h->d_sp_blk_cols.upload(bc, h->stream));
    }
    MIPM_CUDA(h, h->d_sp_valT.alloc((size_t)nnz));
    h->sp_valT_src = nullptr;
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->sp_m = m;
    h->sp_n = n;
    h->sp_nnz = nnz;
    h->has_spmv = true;
    return MIPM_OK;
}

static SpmvJob spmv_job(Handle *h, int trans, double alpha, const double *d_Ax, const double *d_x, double beta, double *d_y)
{
    if (trans == 0)
        return SpmvJob{h->d_sp_blk_rows.p, h->d_sp_rowptr.p, h->d_sp_col.p, nullptr, d_Ax, d_x, alpha, beta, d_y,
                       stream_lanes(h->sp_nnz, h->sp_m), (int)h->sp_nblk_rows};
    const bool cached = h->sp_valT_src == d_Ax && d_Ax;
    return SpmvJob{h->d_sp_blk_cols.p, h->d_sp_colptr.p, h->d_sp_row.p, cached ? nullptr : h->d_sp_pos.p, cached ? h->d_sp_valT.p : d_Ax, d_x,
                   alpha, beta, d_y, stream_lanes(h->sp_nnz, h->sp_n), (int)h->sp_nblk_cols};
}

int mipm_spmv(mipm_handle hh, int trans, double alpha, const double *d_Ax, const double *d_x, double beta, double *d_y)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_spmv) return fail(h, MIPM_ERR_STATE, "mipm_spmv_setup has not been called");
    {
        const int64_t xlen = trans == 0 ? h->sp_n : h->sp_m, ylen = trans == 0 ? h->sp_m : h->sp_n;
        if ((!d_Ax && h->sp_nnz > 0) || (!d_x && xlen > 0) || (!d_y && ylen > 0)) return fail(h, MIPM_ERR_ARG, "null argument");
    }
    if (trans == 0) {
        if (h->sp_m > 0) {
            k_spmv_stream<false><<<(unsigned)h->sp_nblk_rows, 256, 0, h->stream>>>(spmv_job(h, 0, alpha, d_Ax, d_x, beta, d_y));
            MIPM_CHECK_LAUNCH(h);
        }
    } else {
        if (h->sp_n > 0) {
            if (h->sp_valT_src == d_Ax && d_Ax)      // column-ordered copy cached by mipm_spmv_cache_values
                k_spmv_stream<false><<<(unsigned)h->sp_nblk_cols, 256, 0, h->stream>>>(spmv_job(h, 1, alpha, d_Ax, d_x, beta, d_y));
            else
                k_spmv_stream<true><<<(unsigned)h->sp_nblk_cols, 256, 0, h->stream>>>(spmv_job(h, 1, alpha, d_Ax, d_x, beta, d_y));
            MIPM_CHECK_LAUNCH(h);
        }
    }
    return MIPM_OK;
}

// y1 = alpha1 A x1 + beta1 y1 and y2 = alpha2 A' x2 + beta2 y2 in ONE launch (the two products of evaluate_model!,
// src/solver.jl:319-326, and of the residual K d, src/linear_solver.jl:29-35, are independent of each other).
int mipm_spmv_pair(mipm_handle hh, const double *d_Ax, double alpha1, const double *d_x1, double beta1, double *d_y1,
                   double alpha2, const double *d_x2, double beta2, double *d_y2)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_spmv) return fail(h, MIPM_ERR_STATE, "mipm_spmv_setup has not been called");
    if ((!d_Ax && h->sp_nnz > 0) || (h->sp_n > 0 && (!d_x1 || !d_y2)) || (h->sp_m > 0 && (!d_y1 || !d_x2))) return fail(h, MIPM_ERR_ARG, "null argument");
    if (h->sp_m == 0 || h->sp_n == 0 || !(h->sp_valT_src == d_Ax && d_Ax)) {      // no column-ordered copy: two launches
        int rc = mipm_spmv(hh, 0, alpha1, d_Ax, d_x1, beta1, d_y1);
        if (rc != MIPM_OK) return rc;
        return mipm_spmv(hh, 1, alpha2, d_Ax, d_x2, beta2, d_y2);
    }
    k_spmv_stream_pair<<<(unsigned)(h->sp_nblk_rows + h->sp_nblk_cols), 256, 0, h->stream>>>(spmv_job(h, 0, alpha1, d_Ax, d_x1, beta1, d_y1),
                                                                                            spmv_job(h, 1, alpha2, d_Ax, d_x2, beta2, d_y2));
    MIPM_CHECK_LAUNCH(h);
    return MIPM_OK;
}

int mipm_set_grid_limit(mipm_handle hh, int max_ctas)
{
    Handle *h = (Handle *)hh;
    if (!h || max_ctas < 0) return fail(h, MIPM_ERR_ARG, "bad argument");
    h->grid_limit = max_ctas;
    return MIPM_OK;
}

int mipm_spmv_cache_values(mipm_handle hh, const double *d_Ax)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_spmv) return fail(h, MIPM_ERR_STATE, "mipm_spmv_setup has not been called");
    h->sp_valT_src = nullptr;
    if (!d_Ax || h->sp_nnz == 0) return MIPM_OK;       // null: drop the cache
    k_gather32<<<(unsigned)std::min<int64_t>(grid_for(h->sp_nnz, 256), 148 * 16), 256, 0, h->stream>>>(h->sp_nnz, d_Ax, h->d_sp_pos.p, h->d_sp_valT.p);
    MIPM_CHECK_LAUNCH(h);
    h->sp_valT_src = d_Ax;
    return MIPM_OK;
}

int mipm_hess_setup(mipm_handle hh, int64_t n, const int32_t *Hp, const int32_t *Hj, int index_base)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (n < 0 || !Hp) return fail(h, MIPM_ERR_ARG, "bad argument");
    int64_t nnz = Hp[n] - index_base;
    std::vector<int32_t> rp((size_t)n + 1), cj((size_t)nnz);
    for (int64_t i = 0; i <= n; ++i) rp[(size_t)i] = Hp[i] - index_base;
    for (int64_t p = 0; p < nnz; ++p) {
        cj[(size_t)p] = Hj[p] - index_base;
        if (cj[(size_t)p] < 0 || cj[(size_t)p] >= n) return fail(h, MIPM_ERR_ARG, "column index out of range");
    }
    MIPM_CUDA(h, cudaSetDevice(h->device));
    MIPM_CUDA(h, h->d_hs_rowptr.upload(rp, h->stream));
    MIPM_CUDA(h, h->d_hs_col.upload(cj, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->hs_n = n;
    h->hs_nnz = nnz;
    h->has_hess = true;
    return MIPM_OK;
}

int mipm_hess_spmv(mipm_handle hh, double alpha, const double *d_Hx, const double *d_x, double beta, double *d_y)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_hess) return fail(h, MIPM_ERR_STATE, "mipm_hess_setup has not been called");
    if ((!d_Hx && h->hs_nnz > 0) || !d_x || !d_y) return fail(h, MIPM_ERR_ARG, "null argument");
    if (h->hs_n > 0) {
#define MIPM_SPMV_H(G) k_spmv_csr<G><<<grid_for(h->hs_n * G, 256), 256, 0, h->stream>>>(h->hs_n, h->d_hs_rowptr.p, h->d_hs_col.p, d_Hx, d_x, alpha, beta, d_y)
        switch (spmv_group(h->hs_nnz, h->hs_n)) {
        case 2: MIPM_SPMV_H(2); break;
        case 4: MIPM_SPMV_H(4); break;
        case 8: MIPM_SPMV_H(8); break;
        case 16: MIPM_SPMV_H(16); break;
        default: MIPM_SPMV_H(32); break;
        }
#undef MIPM_SPMV_H
        MIPM_CHECK_LAUNCH(h);
    }
    return MIPM_OK;
}

// ---------------------------------------------------------------- linear solver front-end
int mipm_ls_analyze(mipm_handle hh, int64_t n, const int32_t *colptr, const int32_t *rowval, int index_base,
                    int kind, int ordering, const int32_t *user_perm)
{
    Handle *h = (Handle *)hh;
    if (h && !h->host_only) use_handle(h);
    if (!h || n < 0 || !colptr || (kind != MIPM_CHOLESKY && kind != MIPM_LDL && kind != MIPM_LDL_DEFINITE)) return fail(h, MIPM_ERR_ARG, "bad argument");
    int64_t nnz = colptr[n] - index_base;
    if (nnz < 0 || (nnz > 0 && !rowval)) return fail(h, MIPM_ERR_ARG, "bad column pointer");
    std::vector<int32_t> up;
    if (ordering == MIPM_ORDER_USER && user_perm) {
        up.resize((size_t)n);
        for (int64_t k = 0; k < n; ++k) up[(size_t)k] = user_perm[k] - index_base;
    }
    LsOptions opt;
    opt.kind = (kind == MIPM_LDL_DEFINITE) ? MIPM_CHOLESKY : kind;      // the analysis of a definite matrix has no K2 rule
    opt.ordering = ordering;
    opt.host_a2l = h->host_only;
    if (const char *s = std::getenv("MIPM_ND_LEAF")) opt.nd_leaf = std::max(1, atoi(s));
    if (const char *s = std::getenv("MIPM_LDL_DELAY")) opt.ldl_delay_all = (std::strcmp(s, "first") != 0);
    if (const char *s = std::getenv("MIPM_RELAX")) {
        // "always,k1,z1,k2,z2,z3"
        sscanf(s, "%d,%d,%lf,%d,%lf,%lf", &opt.relax_always, &opt.relax_k1, &opt.relax_z1, &opt.relax_k2, &opt.relax_z2, &opt.relax_z3);
    }
    h->has_ls = false;
    h->factorized = false;
    std::string e = ls_analyze(n, colptr, rowval, index_base, opt, up.empty() ? nullptr : up.data(), h->sym);
    if (!e.empty()) return fail(h, MIPM_ERR_ARG, e);
    h->ldl_definite = (kind == MIPM_LDL_DEFINITE);
    if (h->ldl_definite) h->sym.kind = MIPM_LDL;
    h->has_ls = true;
    if (!h->host_only) return ls_device_setup(h);
    return MIPM_OK;
}

int mipm_ls_factorize_async(mipm_handle hh, const double *d_nzval)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze has not been called");
    if (!d_nzval && h->sym.nnz_a > 0) return fail(h, MIPM_ERR_ARG, "null values");
    return ls_factorize_impl(h, d_nzval);
}

int mipm_ls_status(mipm_handle hh, int *status)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!status) return fail(h, MIPM_ERR_ARG, "null argument");
    if (!h->factorized) return fail(h, MIPM_ERR_STATE, "no factorization");
    int info[4];
    MIPM_CUDA(h, cudaMemcpyAsync(info, h->d_info.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    *status = (info[0] == 0) ? MIPM_OK : MIPM_ERR_NOT_FACTORIZED;
    return MIPM_OK;
}

int mipm_ls_factorize(mipm_handle hh, const double *d_nzval, int *status)
{
    int rc = mipm_ls_factorize_async(hh, d_nzval);
    if (rc != MIPM_OK) return rc;
    if (status) return mipm_ls_status(hh, status);
    return MIPM_OK;
}

int mipm_ls_solve(mipm_handle hh, double *d_x, int ir_steps)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->has_ls || !h->factorized) return fail(h, MIPM_ERR_STATE, "solve before factorize");
    if (!d_x && h->sym.n > 0) return fail(h, MIPM_ERR_ARG, "null argument");
    return ls_solve_impl(h, d_x, ir_steps < 0 ? 0 : ir_steps);
}

int mipm_ls_inertia(mipm_handle hh, int64_t *num_pos, int64_t *num_zero, int64_t *num_neg)
{
    Handle *h = (Handle *)hh;
    MIPM_NEED_DEVICE(h);
    if (!h->factorized) return fail(h, MIPM_ERR_STATE, "no factorization");
    int info[4];
    MIPM_CUDA(h, cudaMemcpyAsync(info, h->d_info.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MIPM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (num_neg) *num_neg = info[1];
    if (num_zero) *num_zero = info[2];
    if (num_pos) *num_pos = h->sym.n - info[1] - info[2];
    return MIPM_OK;
}

int mipm_ls_stats(mipm_handle hh, mipm_ls_stats_t *out)
{
    Handle *h = (Handle *)hh;
    if (!h || !out) return MIPM_ERR_ARG;
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze has not been called");
    const LsSymbolic &S = h->sym;
    out->n = S.n;
    out->nnz_a = S.nnz_a;
    out->nnz_l = S.nnz_l;
    out->nnz_l_exact = S.nnz_l_exact;
    out->flops = S.flops;
    out->n_supernodes = S.ns;
    out->n_levels = S.n_levels;
    out->max_front_cols = S.max_front_cols;
    out->max_front_rows = S.max_front_rows;
    out->update_doubles = S.update_doubles;
    out->n_launches = h->n_launch_factor;
    return MIPM_OK;
}

int mipm_ls_symbolic(mipm_handle hh, int32_t **perm, int64_t *n_sn, int32_t **sn_ptr, int32_t **sn_parent,
                     int64_t **row_ptr, int32_t **row_idx)
{
    Handle *h = (Handle *)hh;
    if (!h || !perm || !n_sn || !sn_ptr || !sn_parent || !row_ptr || !row_idx) return MIPM_ERR_ARG;
    if (!h->has_ls) return fail(h, MIPM_ERR_STATE, "mipm_ls_analyze has not been called");
    const LsSymbolic &S = h->sym;
    *perm = host_copy(S.perm, 0);
    *n_sn = S.ns;
    *sn_ptr = host_copy(S.sn_ptr, 0);
    *sn_parent = host_copy(S.sn_parent, 0);
    *row_ptr = host_copy(S.row_ptr, 0);
    *row_idx = host_copy(S.row_idx, 0);
    if (!*perm || !*sn_ptr || !*sn_parent || !*row_ptr || !*row_idx) return fail(h, MIPM_ERR_ALLOC, "host allocation failed");
    return MIPM_OK;
}

}  // extern "C"
