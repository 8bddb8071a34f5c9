// Micro-benchmark of the 64x64 diagonal-block elimination loop used by task_diag (csrc/factor.cu):
// which part of the per-column critical path costs what. One CTA, clock64 around the loop.
// Variants: 0 full; 1 no division; 2 no owner write-back; 3 no barrier (wrong results); 4 barrier + loads only.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 64;
template <int V>
__global__ void __launch_bounds__(256) k(double *A, long long *cyc, double *sink)
{
    __shared__ double colbuf[2 * NB];
    __shared__ double dv[NB];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double a[4][4];
    for (int jj = 0; jj < 4; ++jj)
        for (int i = 0; i < 4; ++i) { int rr = 4 * ty + i, cc = 4 * tx + jj; a[i][jj] = (rr >= cc) ? A[cc * NB + rr] : 0.0; }
    __syncthreads();
    long long c0 = clock64();
    for (int j4 = 0; j4 < NB / 4; ++j4) {
#pragma unroll
        for (int js = 0; js < 4; ++js) {
            const int j = 4 * j4 + js;
            double *cb = colbuf + (j & 1) * NB;
            if (tx == j4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = 4 * ty + i;
                    cb[rr] = (rr > j) ? a[i][js] : 0.0;
                    if (rr == j) dv[j] = a[i][js];
                }
            }
            if (V != 3) __syncthreads();
            double d = dv[j];
            if (!(d > 0.0) || !(d < 1.0e300)) d = 1.0;
            const double scale = (V == 1) ? d * 0.999 : 1.0 / d;
            double lr[4], lc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) lr[i] = cb[4 * ty + i] * scale;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) lc[jj] = cb[4 * tx + jj];
            if (V != 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) a[i][jj] = fma(-lr[i], lc[jj], a[i][jj]);
            } else { a[0][0] += lr[0] + lc[0]; }
            if (V != 2 && tx == j4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = 4 * ty + i;
                    if (rr > j) a[i][js] = lr[i];
                    else if (rr == j) a[i][js] = d;
                }
            }
        }
    }
    long long c1 = clock64();
    double s = 0;
    for (int jj = 0; jj < 4; ++jj) for (int i = 0; i < 4; ++i) s += a[i][jj];
    sink[tid] = s;
    if (tid == 0) cyc[0] = c1 - c0;
}
template <int V> void run(double *A, long long *cyc, double *sink)
{
    long long h = 0, best = 1LL << 60;
    for (int rep = 0; rep < 5; ++rep) {
        k<V><<<1, 256>>>(A, cyc, sink);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (h < best) best = h;
    }
    printf("variant %d: %lld cycles total, %.1f per column\n", V, best, best / 64.0);
}
int main()
{
    double *A, *sink; long long *cyc;
    cudaMalloc(&A, NB * NB * 8); cudaMalloc(&sink, 256 * 8); cudaMalloc(&cyc, 8);
    double h[NB * NB];
    for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 100.0 + r : 1.0 / (1 + r + c);
    cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>(A, cyc, sink); run<1>(A, cyc, sink); run<2>(A, cyc, sink); run<3>(A, cyc, sink); run<4>(A, cyc, sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
