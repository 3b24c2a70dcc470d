// Microbenchmark (exploration): which warps of a block share an SM sub-partition?  Two warps of a 256-thread block run
// an FP64-pipe-bound loop, the others exit at once; a pair on the same sub-partition takes twice as long as a pair
// on different ones.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_smsp_map warp_smsp_map.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int wa, int wb, int iters, double a, double b)
{
    const int w = threadIdx.x >> 5;
    if (w != wa && w != wb) return;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out;
    cudaMalloc(&out, 148 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wb = 0; wb < 8; ++wb) {
        k<<<148, 256>>>(out, 0, wb, 1000, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k<<<148, 256>>>(out, 0, wb, 200000, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("warps 0 and %d: %.3f ms  (%.2f cycles per DFMA per warp)\n", wb, ms, ms * 1e-3 * 1.965e9 / (200000.0 * 8));
    }
    return 0;
}
