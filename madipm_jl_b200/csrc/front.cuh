// Shared device-side declarations of the supernodal factorization (factor.cu) and the triangular solves (solve.cu).
//
// Data layout (all FP64, column-major):
//   panel of supernode s : (k+r) x k at L + lp[s], leading dimension ld = k+r rounded up to even
//                          (every panel column then starts on a 16-byte boundary: TMA bulk copies)
//   update matrix of s   : r x r at U + up[s], ld = r (lower triangle used)
//   inverse diagonal blocks: 64 x 64 with column stride XS = 68 at Dinv + 64*68*(dinv[s] + jb/64)  (inverse of the
//                          L11 block, lower): the padded stride is the shared-memory operand layout, so a block is ONE TMA bulk copy
//   LDL^T only: the pivots d_j of column c0 + j at Dg[c0 + j]
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mipm {

constexpr int NB = 64;          // block width of the dense partial factorization
constexpr int LDS = 65;         // smem leading dimension for NB x NB blocks (odd: conflict-free rows)
constexpr int TILE = 64;        // GEMM tile
constexpr int XS = 68;          // smem row stride of staged operands (68 % 16 == 4: conflict-free DMMA fragment loads)
constexpr int SL_K = 8;         // small leaf front: no children, at most SL_K columns ...
constexpr int SL_N = 32;        // ... and at most SL_N rows: one warp does the whole front in registers

// Everything a task needs to know about a front, in one 64-byte record (4 x 16-byte loads).
struct __align__(16) FrontInfo {
    int32_t k, r, c0, nchild;
    int64_t lp, up;
    int64_t dinv, rowp;
    int64_t childp;
    int32_t parent, total;      // parent front (-1: root); number of completions (children + own tasks) that finish the front
};
static_assert(sizeof(FrontInfo) == 64, "FrontInfo must match Handle::FI64");

__host__ __device__ __forceinline__ int front_ld(int k, int r) { return (k + r + 1) & ~1; }

__device__ __forceinline__ int read_smid()
{
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return (int)s;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace mipm
