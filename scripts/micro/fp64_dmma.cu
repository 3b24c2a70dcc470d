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

// like the LTV kernel: 4 accumulator tiles, 4 k-steps each, every DMMA with its own A and B registers (8 + 8 operands)
__global__ void k16(double *out, int iters, double a0, double b0)
{
    double c[4][2], a[2][4], b[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i / 4][i % 4] = a0 + (threadIdx.x + i) * 1e-9; b[i / 4][i % 4] = b0 - (threadIdx.x + 3 * i) * 1e-9; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dmma(c[mt * 2 + nt][0], c[mt * 2 + nt][1], a[mt][ks], b[nt][ks]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
void run16(int warps_per_smsp, double *out)
{
    const int iters = 4000, blocks = 148 * 4 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k16<<<blocks, 32>>>(out, 50, 1e-3, 1e-3);
    cudaEventRecord(e0);
    k16<<<blocks, 32>>>(out, iters, 1e-3, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc_trip = ms * 1e-3 * 1.965e9 / iters;
    printf("warps/SMSP %d  16 DMMAs (4 tiles x 4 k-steps, distinct operands): %8.1f cycles per trip, %6.2f cycles per DMMA per sub-partition\n",
           warps_per_smsp, cyc_trip, cyc_trip / (16.0 * warps_per_smsp));
}

// the LTV step in caricature: 32 DMMAs (two products, 4 tiles x 4 k-steps, the second reading the first's accumulators)
// followed by ONE dependent chain of L scalar FP64 instructions (the 4 x 4 LU) whose result feeds the next trip's operands
template <int L>
__global__ void kstep(double *out, int iters, double a0, double b0)
{
    double c[4][2], d[4][2], a[2][4];
    double x = a0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i / 4][i % 4] = a0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = b0 + i * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dmma(c[mt * 2 + nt][0], c[mt * 2 + nt][1], a[mt][ks], d[nt * 2 + (ks >> 1)][ks & 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dmma(d[mt * 2 + nt][0], d[mt * 2 + nt][1], a[mt][ks], c[nt * 2 + (ks >> 1)][ks & 1]);
        x = d[3][1] * 1e-30 + x;
#pragma unroll
        for (int i = 0; i < L; ++i) x = fma(x, 0.999999, 1e-9);
        d[0][0] += x * 1e-30;
    }
    double s = x;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int L>
void runstep(int warps_per_smsp, double *out)
{
    const int iters = 2000, blocks = 148 * 4 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kstep<L><<<blocks, 32>>>(out, 50, 1e-3, 1e-3);
    cudaEventRecord(e0);
    kstep<L><<<blocks, 32>>>(out, iters, 1e-3, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / warps_per_smsp;
    printf("warps/SMSP %d  32 DMMA + chain of %3d DFMA: %7.1f cycles per step per sub-partition (pipe work %d)\n",
           warps_per_smsp, L, cyc, 32 * 16 + (L + 2) * 2);
}

// do integer / select instructions issue while a DMMA occupies the FP64 pipe?  4 DMMAs + NI LOP3/IMADs per trip
template <int NI>
__global__ void kint(double *out, int *iout, int iters, double a0, double b0, int ia)
{
    double c[4][2];
    int y[8];
    const double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = i;
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < NI / 4; ++j) y[(i * (NI / 4) + j) % 8] = (y[(i * (NI / 4) + j) % 8] ^ ia) + (int)threadIdx.x;
        }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) t += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int NI>
void runint(int warps_per_smsp, double *out)
{
    const int iters = 4000, blocks = 148 * 4 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kint<NI><<<blocks, 32>>>(out, (int *)(out + 148 * 4 * 8 * 32), 50, 1e-3, 1e-3, 5);
    cudaEventRecord(e0);
    kint<NI><<<blocks, 32>>>(out, (int *)(out + 148 * 4 * 8 * 32), iters, 1e-3, 1e-3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / warps_per_smsp;
    printf("warps/SMSP %d  4 DMMA + %3d integer ops (2 SASS instructions each): %7.1f cycles per trip per warp  (DMMA alone: 64)\n",
           warps_per_smsp, NI, cyc);
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
    printf("# the LTV kernel's pattern\n");
    for (int w : {1, 2, 4, 6}) run16(w, out);
    printf("# products + dependent scalar chain per step\n");
    for (int w : {1, 2, 3, 4, 6, 8}) { runstep<60>(w, out); runstep<120>(w, out); }
    printf("# DMMA + integer instructions\n");
    for (int w : {2, 4}) { runint<0>(w, out); runint<16>(w, out); runint<32>(w, out); runint<64>(w, out); runint<128>(w, out); }
    printf("# DMMA + DFMA mixed (do they share the pipe?)\n");
    run<4, 8>(2, out);
    run<4, 32>(2, out);
    run<8, 16>(4, out);
    run<0, 16>(4, out);
    return 0;
}
