// Microbenchmark (exploration): does an FP64 warp instruction on sm_100a occupy the issue port for 2 cycles, i.e.
// do interleaved integer/select instructions cost extra time in an FP64-bound loop?   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int CH>
__global__ void k(double *out, int *iout, int iters, double a, double b, int ia)
{
    double x[CH];
    int y[8];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int i = 0; i < NI; ++i) y[i % 8] = y[i % 8] * ia + (int)threadIdx.x;   // IMAD
        }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) t += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int NI, int CH>
void run(int warps_per_smsp, double *out, int *iout)
{
    const int iters = 20000, blocks = 148 * 4 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NI, CH><<<blocks, 32>>>(out, iout, 100, 1.0000001, 1e-9, 3);
    cudaEventRecord(e0);
    k<NI, CH><<<blocks, 32>>>(out, iout, iters, 1.0000001, 1e-9, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 8) / warps_per_smsp;   // cycles per (CH DFMA + NI IMAD) per warp
    printf("warps/SMSP %d  chains %d  DFMA %d + IMAD %d per group: %.2f cycles per group per warp  (%.2f per DFMA)\n",
           warps_per_smsp, CH, CH, NI, cyc, cyc / CH);
}
int main()
{
    double *out; int *iout;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8); cudaMalloc(&iout, 148 * 4 * 16 * 32 * 4);
    for (int w : {1, 2, 3, 4, 8}) {
        run<0, 8>(w, out, iout);
        run<4, 8>(w, out, iout);
        run<8, 8>(w, out, iout);
        run<16, 8>(w, out, iout);
        run<0, 2>(w, out, iout);
        run<0, 4>(w, out, iout);
    }
    return 0;
}
