// Dependent-chain latencies (cycles per op, one warp) of the FP64 building blocks the diagonal-block task is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency.bin tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, long long *cyc, double seed, int n)
{
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = seed;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    int idx = threadIdx.x & 31;
    long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) x = fma(x, y, 1e-9);
            if (OP == 1) x = x * y;
            if (OP == 2) x = __drcp_rn(x);
            if (OP == 3) x = 1.0 / x;
            if (OP == 4) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
            if (OP == 5) x = sqrt(x);
            if (OP == 6) { idx = (int)sm[idx] ; }                     // LDS + cvt chain
            if (OP == 7) x = (x > 0.5 && x < 1e300) ? x : y;          // compare + select
            if (OP == 8) x = x + y;
            if (OP == 9) { x = fma(x, y, 1e-9); y = fma(y, x, 1e-9); }   // 2 interleaved... still dependent
            if (OP == 10) { double c1 = y; asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(c1) : "d"(1e-3), "d"(1e-3)); y = c1; }
            if (OP == 11) { double c1 = y, d0 = x * 0.5, d1 = y * 0.5;     // two independent accumulator pairs
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(c1) : "d"(1e-3), "d"(1e-3));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(1e-3), "d"(1e-3));
                y = c1 + d0 + d1; }
        }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0) cyc[0] = c1 - c0;
    out[threadIdx.x] = x + idx + y;
}
template <int OP> void run(const char *name, double seed, double *out, long long *cyc)
{
    const int n = 256;
    k<OP><<<1, 32>>>(out, cyc, seed, n);
    k<OP><<<1, 32>>>(out, cyc, seed, n);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles/op\n", name, (double)h / (n * 8));
}
int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 2048); cudaMalloc(&cyc, 8);
    run<0>("DFMA dependent", 1.0, out, cyc);
    run<1>("DMUL dependent", 1.0, out, cyc);
    run<8>("DADD dependent", 1.0, out, cyc);
    run<2>("__drcp_rn dependent", 1.7, out, cyc);
    run<3>("1.0/x dependent", 1.7, out, cyc);
    run<4>("shfl.f64 dependent", 1.0, out, cyc);
    run<5>("sqrt dependent", 1.7, out, cyc);
    run<6>("LDS.64 + cvt dependent", 3.0, out, cyc);
    run<7>("cmp+select dependent", 1.0, out, cyc);
    run<10>("DMMA m8n8k4 dependent", 1.0, out, cyc);
    run<11>("2 DMMA independent + 2 DADD", 1.0, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
