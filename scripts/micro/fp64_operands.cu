// Microbenchmark (exploration): FP64 pipe throughput on sm_100a as a function of the operand pattern --
// DFMA with two operands shared by all chains vs three distinct register operands per instruction vs a
// constant-bank operand, and DMUL/DADD mixes.   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[16] = {1.0000001, 1.0000002, 1.0000003, 1.0000004, 1.0000005, 1.0000006, 1.0000007, 1.0000008,
                              1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};
template <int MODE, int CH>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[CH], y[CH], z[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * (i + 1) * threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);                       // two operands shared
                if (MODE == 1) x[i] = fma(x[i], y[i], z[i]);                 // three distinct registers
                if (MODE == 2) x[i] = fma(x[i], kc[i % 8], kc[8 + i % 8]);    // constant-bank operand(s)
                if (MODE == 3) x[i] = fma(x[i], y[(i + 1) % CH], x[(i + CH / 2) % CH] );   // cross-chain operands
                if (MODE == 4) x[i] = x[i] * y[i];                            // DMUL two distinct
                if (MODE == 5) x[i] = x[i] + z[i];                            // DADD
                if (MODE == 6) x[i] = fma(x[i], y[i], b);                     // two distinct + one shared
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int CH>
void run(int w, double *out)
{
    const int iters = 20000, blocks = 148 * 4 * w;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, CH><<<blocks, 32>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<MODE, CH><<<blocks, 32>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 8 * CH) / w;
    const char *names[] = {"fma(x,a,b) shared a,b", "fma(x,y,z) 3 distinct regs", "fma(x,c[],c[]) const bank", "fma cross-chain", "mul(x,y)", "add(x,z)", "fma(x,y,b)"};
    printf("warps/SMSP %d chains %2d  %-28s %.2f cycles per FP64 instr\n", w, CH, names[MODE], cyc);
}
int main()
{
    double *out;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8);
    for (int w : {1, 2, 4}) {
        run<0, 8>(w, out); run<1, 8>(w, out); run<2, 8>(w, out); run<3, 8>(w, out); run<4, 8>(w, out); run<5, 8>(w, out); run<6, 8>(w, out);
        run<1, 16>(w, out); run<3, 16>(w, out);
    }
    return 0;
}
