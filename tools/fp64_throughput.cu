// FP64 issue throughput per SM: independent DFMA / DMUL streams and DMMA m8n8k4, 8 warps per CTA, one CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_throughput.bin tools/fp64_throughput.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(double *out, long long *cyc, double seed, int n)
{
    double x[8], y = seed;
    for (int u = 0; u < 8; ++u) x[u] = seed + u + threadIdx.x * 1e-9;
    double c0r = 0, c1r = 0;
    __syncthreads();
    long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) x[u] = fma(x[u], y, 1e-9);
            if (OP == 1) x[u] = x[u] * y;
            if (OP == 2) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0r), "+d"(c1r) : "d"(x[u]), "d"(y));
            }
            if (OP == 3) { double a = x[u], b = x[(u + 1) & 7];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(x[u]), "=d"(x[(u+1)&7]) : "d"(y), "d"(y), "d"(a), "d"(b)); }
        }
    }
    __syncthreads();
    long long c1 = clock64();
    if (threadIdx.x == 0) cyc[0] = c1 - c0;
    double s = c0r + c1r; for (int u = 0; u < 8; ++u) s += x[u];
    out[threadIdx.x] = s;
}
template <int OP> void run(const char *name, int threads, double *out, long long *cyc)
{
    const int n = 512;
    k<OP><<<1, threads>>>(out, cyc, 1.0, n);
    k<OP><<<1, threads>>>(out, cyc, 1.0, n);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double winstr = (double)n * 8 * (threads / 32);
    printf("%-34s threads=%3d  %.2f cycles per warp-instr per SM  (%.1f lanes/clk/SM)\n", name, threads, (double)h / winstr, 32.0 * winstr / (double)h);
}
int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
    run<0>("DFMA independent x8", 32, out, cyc);
    run<0>("DFMA independent x8", 128, out, cyc);
    run<0>("DFMA independent x8", 256, out, cyc);
    run<0>("DFMA independent x8", 512, out, cyc);
    run<1>("DMUL independent x8", 256, out, cyc);
    run<2>("DMMA m8n8k4 same-acc chain", 256, out, cyc);
    run<3>("DMMA m8n8k4 independent", 256, out, cyc);
    run<3>("DMMA m8n8k4 independent", 512, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
