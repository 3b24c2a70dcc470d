// Microbenchmark (exploration): issue cost of the FP64 forms the rollout kernel is made of -- Horner steps with the
// coefficient in a uniform register / in a register, and matrix-style accumulations.   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[8] = {1.589e-10, -2.505e-08, 2.755e-06, -1.984e-04, 8.333e-03, -1.666e-01, 0.5, 1.0};
template <int MODE>
__global__ void k(double *out, int iters, double seed)
{
    constexpr int CH = 8;
    double z[CH], p[CH], a[CH], b[CH];
    double cr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cr[i] = kc[i] * (1.0 + seed * threadIdx.x);        // per-thread copies: registers
#pragma unroll
    for (int i = 0; i < CH; ++i) { z[i] = 1e-3 * (threadIdx.x + i + 1); p[i] = 0.1 * i; a[i] = 1.0 + 1e-9 * i; b[i] = 1e-9 * threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) p[i] = fma(z[i], p[i], kc[r]);            // Horner, coefficient uniform (UR / const bank)
                if (MODE == 1) p[i] = fma(z[i], p[i], cr[r]);            // Horner, coefficient in a register
                if (MODE == 2) p[i] = fma(a[i], b[(i + 1) % CH], p[i]);  // accumulate a product of two registers
                if (MODE == 3) p[i] = fma(z[i], kc[r], p[i]);            // register * uniform + accumulator
                if (MODE == 4) p[i] = fma(z[0], p[i], kc[r]);            // Horner, z shared by all chains (.reuse)
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(int w, double *out)
{
    const int iters = 20000, blocks = 148 * 4 * w;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 32>>>(out, 100, 0.0);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 32>>>(out, iters, 0.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const char *names[] = {"p=fma(z,p,UR)", "p=fma(z,p,Rconst)", "p=fma(a,b,p)", "p=fma(z,UR,p)", "p=fma(z0,p,UR) z shared"};
    printf("warps/SMSP %d  %-26s %.2f cycles per DFMA\n", w, names[MODE], ms * 1e-3 * 1.965e9 / ((double)iters * 48) / w);
}
int main()
{
    double *out;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8);
    for (int w : {1, 2, 3, 4}) { run<0>(w, out); run<1>(w, out); run<2>(w, out); run<3>(w, out); run<4>(w, out); }
    return 0;
}
