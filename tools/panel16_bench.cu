// What costs what in the 16 x 16 warp-level panel factorization (diag_block.cuh step 1): one warp, variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/panel16_bench.bin tools/panel16_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int PW = 16;
// branch-free reciprocal (normal, finite, non-zero argument): MUFU seed (2^-23) + two Newton steps
__device__ __forceinline__ double fast_rcp(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    return x;
}
// V bit0: no reciprocal (multiply instead)   bit1: no shared-memory exchange (use own register)
// V bit2: no pivot test                       bit3: no next-pivot shuffle (use own value)
// V bit4: branch-free reciprocal (fast_rcp) instead of __drcp_rn
template <int V>
__global__ void k(const double *A, double *out, long long *cyc, int reps)
{
    __shared__ double S[PW * 17], cbuf[2 * PW], dv[PW], invd[PW];
    const int lane = threadIdx.x, i = lane & 15;
    const bool fac = lane < PW;
    for (int t = lane; t < PW * PW; t += 32) S[(t / PW) * 17 + (t % PW)] = A[t];
    __syncwarp();
    const unsigned FULL = 0xffffffffu;
    double sink = 0.0;
    int nbad = 0;
    long long c0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        double a[PW];
#pragma unroll
        for (int kk = 0; kk < PW; ++kk) a[kk] = fac ? S[kk * 17 + i] : ((kk == i) ? 1.0 : 0.0);
        double d = __shfl_sync(FULL, a[0], 0);
        double inv = (V & 1) ? d * 0.01 : ((V & 16) ? fast_rcp(d) : __drcp_rn(d));
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const double aj = a[j];
            const double lij = aj * inv;
            double *cj = cbuf + (j & 1) * PW;
            if (!(V & 2)) { if (fac) cj[i] = aj; }
            double d_n = 1.0, inv_n = 1.0;
            if (j + 1 < PW) {
                const double own = fma(-lij, aj, a[j + 1]);
                d_n = (V & 8) ? own : __shfl_sync(FULL, own, j + 1);
                inv_n = (V & 1) ? d_n * 0.01 : ((V & 16) ? fast_rcp(d_n) : __drcp_rn(d_n));
            }
            if (!(V & 2)) __syncwarp();
#pragma unroll
            for (int kk = j + 1; kk < PW; ++kk) a[kk] = fma(-lij, (V & 2) ? aj * (1.0 + kk) : cj[kk], a[kk]);
            if (fac) { if (i > j) a[j] = lij; else if (i == j) a[j] = d; }
            if (lane == j) { dv[j] = d; invd[j] = inv; }
            if (j + 1 < PW) {
                if (!(V & 4)) {
                    if (!(d_n > 0.0) || !(d_n < 1e300)) { nbad++; d_n = 1.0; inv_n = 1.0; if (lane == j + 1) a[j + 1] = d_n; }
                }
                d = d_n; inv = inv_n;
            }
        }
#pragma unroll
        for (int kk = 0; kk < PW; ++kk) sink += a[kk];
    }
    long long c1 = clock64();
    if (lane == 0) cyc[0] = c1 - c0;
    out[lane] = sink + nbad + dv[i] + invd[i];
}
template <int V> void run(const char *name, const double *A, double *out, long long *cyc)
{
    const int reps = 200;
    k<V><<<1, 32>>>(A, out, cyc, reps);
    k<V><<<1, 32>>>(A, out, cyc, reps);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %.0f cycles per panel (%.1f per column)\n", name, (double)h / reps, (double)h / reps / PW);
}
int main()
{
    double hA[PW * PW];
    for (int c = 0; c < PW; ++c) for (int r = 0; r < PW; ++r) hA[c * PW + r] = (r == c) ? 20.0 + r : 1.0 / (1 + r + c);
    double *A, *out; long long *cyc; cudaMalloc(&A, sizeof(hA)); cudaMalloc(&out, 512); cudaMalloc(&cyc, 8);
    cudaMemcpy(A, hA, sizeof(hA), cudaMemcpyHostToDevice);
    run<0>("full", A, out, cyc);
    run<1>("no reciprocal", A, out, cyc);
    run<2>("no smem exchange", A, out, cyc);
    run<4>("no pivot test", A, out, cyc);
    run<8>("no pivot shuffle", A, out, cyc);
    run<15>("none of them (DFMA work only)", A, out, cyc);
    run<16>("fast_rcp", A, out, cyc);
    run<20>("fast_rcp, no pivot test", A, out, cyc);
    run<3>("no rcp, no exchange", A, out, cyc);
    run<6>("no exchange, no test", A, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
