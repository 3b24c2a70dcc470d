// Microbenchmark (round 2, VERDICT item 4 / DESIGN section 9): where do the ~950 cycles per warp-step of rollout_kernel
// go?  The kernel's per-step work -- control law, stage cost, one RK4 step of the UA double pendulum: 314 FP64
// instructions -- is replayed from the SAME device functions (csrc/ilqr_systems.cuh) in three forms:
//   MODE 0  arithmetic only: nominal, gains in registers, nothing stored
//   MODE 1  + the 10 global loads per step (x_old, u_old, k, K) with the kernel's batch-innermost addressing
//   MODE 2  + the 5 global stores per step (candidate x, u)                    [= rollout_kernel's loop]
// at 1..6 warps per SM sub-partition (one-warp blocks, as the solver launches them).  Cycles per warp-step = time x
// clock / (steps x warps per sub-partition).  If MODE 0 already saturates near 950 the FP64 stream itself (operand
// forms, register banks) is the limit; if it runs near the 2-cycles-per-instruction model (~630-700) the difference is
// the price of addressing, loads and stores next to the FP64 pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../iterative-linear-quadratic-regulator_b200/csrc \
//        -I../../include -o rollout_replay rollout_replay.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ilqr_systems.cuh"
using namespace ilqr;
typedef DoublePendulumSys<double, 1> Sys;
typedef DiagCost<double, 4, 1> Cost;

template <int MODE>
__global__ void replay(const __grid_constant__ Sys sys, const __grid_constant__ Cost qc, int N, int B, double alpha,
                       const double *__restrict__ x0, const double *__restrict__ X_old, const double *__restrict__ U_old,
                       const double *__restrict__ k, const double *__restrict__ K, double *__restrict__ Xw,
                       double *__restrict__ Uw, double *__restrict__ cost_out)
{
    constexpr int n = 4;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double x[n], cost = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = x0[(size_t)i * B + b];
    double xo[n], uo, kk, Kr[n];
#pragma unroll
    for (int i = 0; i < n; ++i) { xo[i] = X_old[(size_t)i * B + b]; Kr[i] = K[(size_t)i * B + b]; }
    uo = U_old[b];
    kk = k[b];
    for (int t = 0; t < N; ++t) {
        if (MODE >= 1) {
#pragma unroll
            for (int i = 0; i < n; ++i) {
                xo[i] = X_old[((size_t)t * n + i) * B + b];
                Kr[i] = K[((size_t)t * n + i) * B + b];
            }
            uo = U_old[(size_t)t * B + b];
            kk = k[(size_t)t * B + b];
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i) s += Kr[i] * (x[i] - xo[i]);
        double u[1] = { uo + alpha * kk + s };
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < n; ++i) Xw[((size_t)t * n + i) * B + b] = x[i];
            Uw[(size_t)t * B + b] = u[0];
        }
        cost += qc.stage(x, u);
        double xn[n];
        step<RK4>(sys, qc.dt, x, u, xn, 0.0);
#pragma unroll
        for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    cost_out[b] = cost + qc.terminal(x);
}

// two step sizes of one trajectory per thread: the nominal / gains are loaded once, the two rollouts interleave
template <int MODE>
__global__ void replay2(const __grid_constant__ Sys sys, const __grid_constant__ Cost qc, int N, int B, double alpha0, double alpha1,
                        const double *__restrict__ x0, const double *__restrict__ X_old, const double *__restrict__ U_old,
                        const double *__restrict__ k, const double *__restrict__ K, double *__restrict__ Xw,
                        double *__restrict__ Uw, double *__restrict__ cost_out)
{
    constexpr int n = 4;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double x[2][n], cost[2] = { 0.0, 0.0 };
#pragma unroll
    for (int i = 0; i < n; ++i) x[0][i] = x[1][i] = x0[(size_t)i * B + b];
    double xo[n], uo, kk, Kr[n];
#pragma unroll
    for (int i = 0; i < n; ++i) { xo[i] = X_old[(size_t)i * B + b]; Kr[i] = K[(size_t)i * B + b]; }
    uo = U_old[b];
    kk = k[b];
    for (int t = 0; t < N; ++t) {
        if (MODE >= 1) {
#pragma unroll
            for (int i = 0; i < n; ++i) {
                xo[i] = X_old[((size_t)t * n + i) * B + b];
                Kr[i] = K[((size_t)t * n + i) * B + b];
            }
            uo = U_old[(size_t)t * B + b];
            kk = k[(size_t)t * B + b];
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) s += Kr[i] * (x[a][i] - xo[i]);
            double u[1] = { uo + (a ? alpha1 : alpha0) * kk + s };
            if (MODE >= 2) {
#pragma unroll
                for (int i = 0; i < n; ++i) Xw[(((size_t)a * (N + 1) + t) * n + i) * B + b] = x[a][i];
                Uw[((size_t)a * N + t) * B + b] = u[0];
            }
            cost[a] += qc.stage(x[a], u);
            double xn[n];
            step<RK4>(sys, qc.dt, x[a], u, xn, 0.0);
#pragma unroll
            for (int i = 0; i < n; ++i) x[a][i] = xn[i];
        }
    }
    cost_out[b] = cost[0] + qc.terminal(x[0]);
    cost_out[B + b] = cost[1] + qc.terminal(x[1]);
}

int main()
{
    const int N = 500, WMAX = 6, BMAX = 148 * 4 * 32 * WMAX;
    Sys sys;
    sys.c = 1.0; sys.m11_0 = 1.25 + 1.0 + 0.25 + 1.0 / 6; sys.m12_0 = 0.25 + 1.0 / 12; sys.g1 = 9.81 / 2; sys.g2 = 9.81 * 1.5;
    sys.d1 = sys.d2 = 0.1;
    QuadCost<double, 4, 1> q;
    q.dt = 0.01;
    for (int i = 0; i < 4; ++i) {
        q.xt[i] = i == 0 ? 3.14159265358979 : 0.0;
        for (int j = 0; j < 4; ++j) { q.Qs[i][j] = i == j ? (i < 2 ? 1.0 : 0.1) : 0.0; q.Qfs[i][j] = i == j ? (i < 2 ? 1000.0 : 100.0) : 0.0; }
    }
    q.Rs[0][0] = 1.0; q.diag = 1; q.monotone = 1;
    Cost qc(q);
    double *x0, *X, *U, *k, *K, *Xw, *Uw, *c;
    cudaMalloc(&x0, 8ull * 4 * BMAX); cudaMalloc(&X, 8ull * 4 * (N + 1) * BMAX); cudaMalloc(&U, 8ull * N * BMAX);
    cudaMalloc(&k, 8ull * N * BMAX); cudaMalloc(&K, 8ull * 4 * N * BMAX); cudaMalloc(&Xw, 2 * 8ull * 4 * (N + 1) * BMAX);
    cudaMalloc(&Uw, 2 * 8ull * N * BMAX); cudaMalloc(&c, 2 * 8ull * BMAX);
    cudaMemset(x0, 0, 8ull * 4 * BMAX); cudaMemset(X, 0, 8ull * 4 * (N + 1) * BMAX); cudaMemset(U, 0, 8ull * N * BMAX);
    cudaMemset(k, 0, 8ull * N * BMAX); cudaMemset(K, 0, 8ull * 4 * N * BMAX);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double ghz = 1.965;
    for (int w : {1, 2, 3, 4, 6}) {
        const int B = 148 * 4 * 32 * w;
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) replay<0><<<B / 32, 32>>>(sys, qc, N, B, 0.5, x0, X, U, k, K, Xw, Uw, c);
                if (mode == 1) replay<1><<<B / 32, 32>>>(sys, qc, N, B, 0.5, x0, X, U, k, K, Xw, Uw, c);
                if (mode == 2) replay<2><<<B / 32, 32>>>(sys, qc, N, B, 0.5, x0, X, U, k, K, Xw, Uw, c);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("warps/SMSP %d  mode %d (%s): %7.3f ms  %7.1f cycles per warp-step  (%s)\n", w, mode,
                   mode == 0 ? "arithmetic only" : mode == 1 ? "+ loads" : "+ loads + stores", best,
                   best * 1e-3 * ghz * 1e9 / (N * (double)w), cudaGetErrorString(cudaGetLastError()));
        }
    }
    // two rollouts per thread (one warp then carries 64 rollouts): cycles per warp-step for BOTH
    for (int w : {1, 2, 3}) {
        const int B = 148 * 4 * 32 * w;
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) replay2<0><<<B / 32, 32>>>(sys, qc, N, B, 0.5, 0.25, x0, X, U, k, K, Xw, Uw, c);
                if (mode == 1) replay2<1><<<B / 32, 32>>>(sys, qc, N, B, 0.5, 0.25, x0, X, U, k, K, Xw, Uw, c);
                if (mode == 2) replay2<2><<<B / 32, 32>>>(sys, qc, N, B, 0.5, 0.25, x0, X, U, k, K, Xw, Uw, c);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("PAIRED warps/SMSP %d  mode %d: %7.3f ms  %7.1f cycles per warp-step for two rollouts = %7.1f per rollout  (%s)\n", w, mode,
                   best, best * 1e-3 * ghz * 1e9 / (N * (double)w), best * 1e-3 * ghz * 1e9 / (N * (double)w) / 2,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
