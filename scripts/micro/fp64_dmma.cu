// Microbenchmark (exploration): throughput and dependent latency of the FP64 tensor-core instruction
// mma.sync.aligned.m8n8k4.row.col.f64 on sm_100a, against the DFMA rate measured by fp64_issue.cu
// (one DFMA warp instruction = 32 FMA per 2 cycles per SM sub-partition = 16 FMA/cycle).
// One DMMA = 8 x 8 x 4 = 256 FMA: at the vector rate it would issue every 16 cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_dmma fp64_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CH independent accumulator tiles per warp; every trip issues CH DMMAs (chain length = iters per tile)
template <int CH, int NF>
__global__ void k(double *out, int iters, double a0, double b0)
{
    double c[CH > 0 ? CH : 1][2], f[NF > 0 ? NF : 1];
    const double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = i;
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);          // NF DFMAs next to the DMMAs: same pipe?
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, int NF>
void run(int warps_per_smsp, double *out)
{
    const int iters = 4000, blocks = 148 * 4 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH, NF><<<blocks, 32>>>(out, 50, 1e-3, 1e-3);
    cudaEventRecord(e0);
    k<CH, NF><<<blocks, 32>>>(out, iters, 1e-3, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc_trip = ms * 1e-3 * 1.965e9 / iters;                       // cycles per trip of one sub-partition
    const double per_dmma = cyc_trip / ((double)CH * warps_per_smsp);
    const double tflops = 2.0 * (256.0 * CH + 32.0 * NF) * iters * blocks / (ms * 1e-3) * 1e-12;
    printf("warps/SMSP %d  tiles %d  +DFMA %d : %8.1f cycles per trip, %6.2f cycles per DMMA per sub-partition, %6.2f TFLOP/s%s\n",
           warps_per_smsp, CH, NF, cyc_trip, per_dmma, tflops, cudaGetLastError() == cudaSuccess ? "" : "  (CUDA error)");
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8);
    printf("# dependent chain (1 tile, 1 warp): latency; many tiles / warps: throughput\n");
    run<1, 0>(1, out);
    run<2, 0>(1, out);
    run<4, 0>(1, out);
    run<8, 0>(1, out);
    for (int w : {2, 4, 8}) { run<1, 0>(w, out); run<4, 0>(w, out); run<8, 0>(w, out); }
    printf("# DMMA + DFMA mixed (do they share the pipe?)\n");
    run<4, 8>(2, out);
    run<4, 32>(2, out);
    run<8, 16>(4, out);
    run<0, 16>(4, out);
    return 0;
}
