// Which SM does each CTA of a cooperative (256 threads, 3 CTAs/SM, 69,632 B dynamic smem) launch land on?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/smid_map.bin tools/smid_map.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(256, 3) k(int *smid)
{
    extern __shared__ double sm[];
    if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid[blockIdx.x] = (int)s; sm[0] = 1.0; }
    cg::this_grid().sync();
}
int main()
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    size_t smem = 69632;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k, 256, smem);
    int grid = per * nsm;
    int *d; cudaMalloc(&d, grid * sizeof(int));
    void *args[] = {&d};
    cudaLaunchCooperativeKernel((void *)k, dim3(grid), dim3(256), args, smem, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> h(grid); cudaMemcpy(h.data(), d, grid * sizeof(int), cudaMemcpyDeviceToHost);
    printf("err=%d nsm=%d per=%d grid=%d\n", (int)e, nsm, per, grid);
    for (int i = 0; i < grid; ++i) printf("%d%c", h[i], (i % 37 == 36) ? '\n' : ' ');
    printf("\n");
    return 0;
}
