// Microbenchmark (exploration): does an FP64 instruction whose DESTINATION is a fresh register (not one of its sources)
// cost more issue time than an in-place accumulate?   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[2] = {1e-9, 0.5};
template <int MODE>
__global__ void k(double *out, int iters, double e)
{
    constexpr int CH = 6;
    double a[CH], b[CH], s[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = 1e-3 * (threadIdx.x + i + 1); b[i] = 1.0 + 1e-6 * (threadIdx.x * (i + 1)); s[i] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                a[i] += e;                                                   // DADD in place
                if (MODE == 0) s[i] = fma(a[i], b[i], s[i]);                 // DFMA in place, three registers
                if (MODE == 1) { const double t = fma(a[i], b[i], kc[1]); s[i] += t; }   // DFMA fresh dest (2 regs + UR), DADD in place
                if (MODE == 2) { const double t = a[i] * b[i]; s[i] += t; }  // DMUL fresh dest, DADD in place
                if (MODE == 3) s[i] += a[i];                                 // DADD only (reference: 2 DADD)
            }
        }
    }
    double z = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) z += s[i] + a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = z;
}
template <int MODE>
void run(int w, double *out)
{
    const int iters = 20000, blocks = 148 * 4 * w;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 32>>>(out, 100, 1e-12);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 32>>>(out, iters, 1e-12);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const char *names[] = {"a+=e; s=fma(a,b,s)", "a+=e; t=fma(a,b,UR); s+=t", "a+=e; t=a*b; s+=t", "a+=e; s+=a"};
    printf("warps/SMSP %d  %-28s %.2f cycles per chain-step\n", w, names[MODE], ms * 1e-3 * 1.965e9 / ((double)iters * 24) / w);
}
int main()
{
    double *out;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8);
    for (int w : {1, 2, 4}) { run<0>(w, out); run<1>(w, out); run<2>(w, out); run<3>(w, out); }
    return 0;
}
