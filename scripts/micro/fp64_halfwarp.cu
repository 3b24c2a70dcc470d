// Microbenchmark (exploration): does an FP64 warp instruction with only 16 (or 8) active lanes occupy the 16-lane
// FP64 pipe of an sm_100a sub-partition for fewer cycles than a full warp?   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double a, int active_lanes, int pattern)
{
    const int lane = threadIdx.x & 31;
    // pattern 0: lanes [0, active); 1: even lanes only (active = 16); 2: lanes [16, 16 + active)
    const bool on = pattern == 0 ? lane < active_lanes : pattern == 1 ? (lane & 1) == 0 : (lane >= 16 && lane < 16 + active_lanes);
    if (!on) return;
    double x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = 1e-9 * (i + 1); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, y[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
void run(int w, int active, int pattern, double *out)
{
    const int iters = 20000, blocks = 148 * 4 * w;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<8><<<blocks, 32>>>(out, 100, 1.0000001, active, pattern);
    cudaEventRecord(e0);
    k<8><<<blocks, 32>>>(out, iters, 1.0000001, active, pattern);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SMSP %d  active lanes %2d pattern %d: %.2f cycles per DFMA per warp\n", w, active, pattern,
           ms * 1e-3 * 1.965e9 / ((double)iters * 64) / w);
}
int main()
{
    double *out;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8);
    for (int w : {2, 4, 8}) {
        run(w, 32, 0, out); run(w, 16, 0, out); run(w, 8, 0, out); run(w, 16, 1, out); run(w, 16, 2, out); run(w, 1, 0, out);
    }
    return 0;
}
