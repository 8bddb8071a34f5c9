// Stand-alone development harness for the panel-blocked 64x64 diagonal-block task (csrc/factor.cu: task_diag):
// correctness against a host Cholesky / LDL^T and timing alone (1 CTA) and saturated (3 CTAs per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/diag_v2_bench.bin tools/diag_v2_bench.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
constexpr int NB = 64;
#include "../madipm_jl_b200/csrc/diag_block.cuh"

template <bool LDL>
__global__ void __launch_bounds__(256, 3) k(double *A, int N, int nb, double *Dinv, int *info, long long *cyc, int reps)
{
    extern __shared__ double smem[];
    double *P = A + (size_t)blockIdx.x * N * NB;
    long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
        mipm_diag::diag_block<LDL, true>(P, N, nb, Dinv + (size_t)blockIdx.x * NB * NB, NB, 1e-13, info, smem, false, (blockIdx.x == 0 && gridDim.x == 1) ? cyc + 8 : nullptr);
        __syncthreads();
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - c0;
}

template <bool LDL>
static int run(int nblk, int nb, bool check, int special = 0)
{
    const int N = 96;
    std::vector<double> h((size_t)nblk * N * NB, 0.0), ref;
    srand(1);
    for (int b = 0; b < nblk; ++b) {
        double *M = &h[(size_t)b * N * NB];
        for (int c = 0; c < nb; ++c)
            for (int r = c; r < nb; ++r) {
                double v = (double)rand() / RAND_MAX - 0.5;
                M[c * N + r] = (r == c) ? (LDL && (c % 3 == 1) ? -(8.0 + v) : 8.0 + v) : v * 0.5;
            }
    }
    if (special == 1) h[20 * N + 20] = -3.0;                      // Cholesky breakdown: flagged, run stays finite
    if (special == 2) {                                           // LDL^T: exact zero pivot -> perturbed to 1e-13
        for (int c = 0; c < nb; ++c) for (int r = c; r < nb; ++r) if (r == 7 || c == 7) h[c * N + r] = 0.0;
    }
    ref = h;
    double *A, *Dinv; int *info; long long *cyc;
    cudaMalloc(&A, h.size() * 8); cudaMalloc(&Dinv, (size_t)nblk * NB * NB * 8); cudaMalloc(&info, 16); cudaMalloc(&cyc, (nblk + 16) * 8);
    cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(info, 0, 16);
    size_t smem = (2 * 64 * 68 + 256) * 8;
    cudaFuncSetAttribute(k<LDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<LDL><<<nblk, 256, smem>>>(A, N, nb, Dinv, info, cyc, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> hc(nblk);
    cudaMemcpy(hc.data(), cyc, nblk * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : hc) mx = v > mx ? v : mx;
    if (nblk == 1) { long long q[7]; cudaMemcpy(q, cyc + 8, 56, cudaMemcpyDeviceToHost);
        printf("   cycles: load %lld | factor16 %lld | trsm %lld | update %lld | inv diag %lld | inv rows %lld | store %lld\n", q[0], q[1], q[2], q[3], q[4], q[5], q[6]); }
    int hinfo[4]; cudaMemcpy(hinfo, info, 16, cudaMemcpyDeviceToHost);
    printf("%s nblk=%d nb=%d: max cycles %lld (%.2f us @1.965GHz) info=%d,%d,%d\n", LDL ? "LDL" : "CHOL", nblk, nb, mx, mx / 1965.0, hinfo[0], hinfo[1], hinfo[2]);
    if (check) {
        std::vector<double> L((size_t)N * NB), Di((size_t)NB * NB);
        cudaMemcpy(L.data(), A, L.size() * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(Di.data(), Dinv, Di.size() * 8, cudaMemcpyDeviceToHost);
        const double *M = ref.data();
        double err = 0, erri = 0;
        for (int c = 0; c < nb; ++c)
            for (int r = c; r < nb; ++r) {
                double s = 0;
                for (int q = 0; q <= c; ++q) {
                    double lr = (r == q) ? (LDL ? 1.0 : L[q * N + r]) : L[q * N + r];
                    double lc = (c == q) ? (LDL ? 1.0 : L[q * N + c]) : L[q * N + c];
                    s += LDL ? lr * L[q * N + q] * lc : lr * lc;
                }
                err = fmax(err, fabs(s - M[c * N + r]));
            }
        // L * Dinv = I (LDL: unit-lower L)
        for (int c = 0; c < nb; ++c)
            for (int r = 0; r < nb; ++r) {
                double s = 0;
                for (int q = 0; q <= r; ++q) {
                    double l = (q == r) ? (LDL ? 1.0 : L[q * N + r]) : L[q * N + r];
                    s += l * ((q >= c) ? Di[c * NB + q] : 0.0);
                }
                erri = fmax(erri, fabs(s - (r == c ? 1.0 : 0.0)));
            }
        double pad = 0;   // entries of Dinv outside nb must be the identity
        for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) if (r >= nb || c >= nb) pad = fmax(pad, fabs(Di[c * NB + r] - ((r == c) ? 1.0 : 0.0)));
        printf("   |L L' - A| = %.3e   |L Linv - I| = %.3e   pad err %.1e\n", err, erri, pad);
    }
    cudaFree(A); cudaFree(Dinv); cudaFree(info); cudaFree(cyc);
    return 0;
}

int main()
{
    run<false>(1, 64, true); run<false>(1, 37, true); run<true>(1, 64, true); run<true>(1, 50, true);
    run<false>(1, 5, true); run<false>(1, 16, true); run<true>(1, 20, true); run<true>(1, 33, true); run<false>(1, 48, true); run<false>(444, 20, false);
    run<false>(1, 64, false, 1); run<true>(1, 64, true, 2);
    run<false>(1, 64, false); run<false>(1, 64, false); run<false>(148, 64, false); run<false>(444, 64, false); run<false>(888, 64, false);
    return 0;
}
