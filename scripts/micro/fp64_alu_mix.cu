// Microbenchmark (round 2): what does an ALU-pipe instruction (LOP3 / IADD3 / SEL, as the quadrant
// selects and address arithmetic of the rollout loop) cost next to FP64 instructions on sm_100a?  The round-1 model
// (2 cycles per FP64 instruction, integer work free) predicts ~700 cycles per warp-step for rollout_kernel, the
// kernel saturates at ~880-950.  Each kernel runs CH independent two-register DFMA chains plus NA extra instructions
// of one kind per group, at 1..8 warps per SM sub-partition.     nvcc -arch=sm_100a -O3 -o fp64_alu_mix fp64_alu_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
enum { K_NONE, K_IMAD, K_LOP3, K_IADD, K_SEL };   // K_IADD: ptxas emits a mix of IADD3 (ALU pipe) and IMAD.IADD (FMA pipe)
template <int KIND, int NA, int CH>
__global__ void k(double *out, int *iout, int iters, double a, double b, int ia, int ib)
{
    double x[CH];
    int y[8];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                int &v = y[i % 8];
                const int w = y[(i + 3) % 8];
                if (KIND == K_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(ia), "r"(w));
                if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(ia), "r"(w));
                if (KIND == K_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(w));
                if (KIND == K_SEL) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(v) : "r"(w), "r"(ib));
            }
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
template <int KIND, int NA, int CH>
void run(const char *name, int warps_per_smsp, double *out, int *iout, double ghz)
{
    const int iters = 20000, blocks = 148 * 4 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, NA, CH><<<blocks, 32>>>(out, iout, 100, 1.0000001, 1e-9, 3, 1);
    cudaEventRecord(e0);
    k<KIND, NA, CH><<<blocks, 32>>>(out, iout, iters, 1.0000001, 1e-9, 3, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * ghz * 1e9 / ((double)iters * 8) / warps_per_smsp;
    printf("warps/SMSP %d  DFMA %d + %2d %-7s per group: %6.2f cycles per group per warp  (model FP64-only %d, FP64+2*ALU %d, issue %d)\n",
           warps_per_smsp, CH, NA, name, cyc, 2 * CH, 2 * CH + 2 * NA, CH + NA);
}
int main()
{
    double *out; int *iout;
    cudaMalloc(&out, 148 * 4 * 16 * 32 * 8); cudaMalloc(&iout, 148 * 4 * 16 * 32 * 4);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = 1.965;
    printf("device clock attribute %d kHz; cycles computed at %.3f GHz\n", khz, ghz);
    for (int w : {1, 2, 4, 8}) {
        run<K_NONE, 0, 8>("none", w, out, iout, ghz);
        run<K_IMAD, 4, 8>("IMAD", w, out, iout, ghz);
        run<K_IMAD, 8, 8>("IMAD", w, out, iout, ghz);
        run<K_LOP3, 4, 8>("LOP3", w, out, iout, ghz);
        run<K_LOP3, 8, 8>("LOP3", w, out, iout, ghz);
        run<K_IADD, 4, 8>("IADD", w, out, iout, ghz);
        run<K_IADD, 8, 8>("IADD", w, out, iout, ghz);
        run<K_SEL, 4, 8>("SETP+SEL", w, out, iout, ghz);
        run<K_SEL, 8, 8>("SETP+SEL", w, out, iout, ghz);
        run<K_LOP3, 16, 8>("LOP3", w, out, iout, ghz);
        run<K_NONE, 0, 4>("none", w, out, iout, ghz);
        run<K_LOP3, 4, 4>("LOP3", w, out, iout, ghz);
    }
    return 0;
}
