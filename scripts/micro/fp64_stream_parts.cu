// Microbenchmark (round 2): which part of the rollout step's FP64 stream costs more than 2 cycles per instruction?
// The step (rollout_replay.cu, arithmetic only: ~850 cycles per warp-step for ~314 FP64 instructions) is split into
//   PART 0  the 8 sincos_t of an RK4 step (2 angles x 4 stages), results folded back into the angles
//   PART 1  everything else: 4 x the double-pendulum acceleration from GIVEN sines/cosines (cheap stand-ins), the RK4
//           combination, the control law and the stage cost
//   PART 2  the whole step (= rollout_replay mode 0)
// each at 4 warps per SM sub-partition; FP64 instruction counts per loop trip are read from the SASS
// (scripts/sass_flops.py scripts/micro/fp64_stream_parts <kernel>), cycles per trip are printed here.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../iterative-linear-quadratic-regulator_b200/csrc \
//        -I../../include -o fp64_stream_parts fp64_stream_parts.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ilqr_systems.cuh"
using namespace ilqr;
typedef DoublePendulumSys<double, 1> Sys;
typedef DiagCost<double, 4, 1> Cost;

// DoublePendulumSys::acc with the trigonometry supplied by the caller
template <bool REAL_TRIG>
__device__ __forceinline__ void acc_core(const Sys &s, const double *x, const double *u, double *a)
{
    const double q1d = x[2], q2d = x[3];
    double s1, c1, s2, c2;
    if (REAL_TRIG) {
        sincos_t(x[0], &s1, &c1);
        sincos_t(x[1], &s2, &c2);
    } else {                                     // 2 FP64 instructions per pair instead of 22
        s1 = x[0]; c1 = fma(-0.5 * x[0], x[0], 1.0);
        s2 = x[1]; c2 = fma(-0.5 * x[1], x[1], 1.0);
    }
    const double s12 = s1 * c2 + c1 * s2;
    const double m11 = s.m11_0 + s.c * c2, m12 = s.m12_0 + 0.5 * s.c * c2, m22 = s.m12_0;
    const double inv = rcp_t(m11 * m22 - m12 * m12);
    const double hcs2 = (0.5 * s.c) * s2, g = s.g1 * s12;
    const double w = fma(2.0, q1d, q2d) * q2d;
    double h1 = fma(hcs2, w, u[0]) - g;
    h1 = fma(-s.g2, s1, h1);
    h1 = fma(-s.d1, q1d, h1);
    double h2 = -fma(hcs2, q1d * q1d, g);
    h2 = fma(-s.d2, q2d, h2);
    a[0] = inv * (m22 * h1 - m12 * h2);
    a[1] = inv * (m11 * h2 - m12 * h1);
}

template <bool REAL_TRIG>
__device__ __forceinline__ void rk4(const Sys &s, double dt, const double *x, const double *u, double *xn)
{
    double k1[4], k2[4], k3[4], k4[4], xs[4], a[2];
    acc_core<REAL_TRIG>(s, x, u, a); k1[0] = x[2]; k1[1] = x[3]; k1[2] = a[0]; k1[3] = a[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + (dt * 0.5) * k1[i];
    acc_core<REAL_TRIG>(s, xs, u, a); k2[0] = xs[2]; k2[1] = xs[3]; k2[2] = a[0]; k2[3] = a[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + (dt * 0.5) * k2[i];
    acc_core<REAL_TRIG>(s, xs, u, a); k3[0] = xs[2]; k3[1] = xs[3]; k3[2] = a[0]; k3[3] = a[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = x[i] + dt * k3[i];
    acc_core<REAL_TRIG>(s, xs, u, a); k4[0] = xs[2]; k4[1] = xs[3]; k4[2] = a[0]; k4[3] = a[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = x[i] + (dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
}

template <int PART>
__global__ void parts(const __grid_constant__ Sys sys, const __grid_constant__ Cost qc, int N, double alpha, const double *__restrict__ in,
                      double *__restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double x[4], cost = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = in[i * 32 + (b & 31)] + 1e-3 * b;
    const double xo[4] = { in[128], in[129], in[130], in[131] }, Kr[4] = { in[132], in[133], in[134], in[135] };
    const double uo = in[136], kk = in[137];
    for (int t = 0; t < N; ++t) {
        if (PART == 0) {
            double ang[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) ang[s] = x[s & 1] + 0.005 * s * x[2 + (s & 1)];
            double acc = 0.0;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                double sn, cs;
                sincos_t(ang[s], &sn, &cs);
                acc = fma(sn, cs, acc);
            }
            x[0] = fma(1e-3, acc, x[0]);
            x[1] = fma(-1e-3, acc, x[1]);
        } else {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) s += Kr[i] * (x[i] - xo[i]);
            double u[1] = { uo + alpha * kk + s };
            cost += qc.stage(x, u);
            double xn[4];
            if (PART == 1) rk4<false>(sys, qc.dt, x, u, xn);
            else rk4<true>(sys, qc.dt, x, u, xn);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = xn[i];
        }
    }
    out[b] = cost + x[0] + x[1] + x[2] + x[3];
}

int main()
{
    const int N = 2000, W = 4, B = 148 * 4 * 32 * W;
    Sys sys;
    sys.c = 1.0; sys.m11_0 = 1.25 + 1.0 + 0.25 + 1.0 / 6; sys.m12_0 = 0.25 + 1.0 / 12; sys.g1 = 9.81 / 2; sys.g2 = 9.81 * 1.5;
    sys.d1 = sys.d2 = 0.1;
    QuadCost<double, 4, 1> q;
    q.dt = 0.01;
    for (int i = 0; i < 4; ++i) {
        q.xt[i] = i == 0 ? 3.14159265358979 : 0.0;
        for (int j = 0; j < 4; ++j) { q.Qs[i][j] = i == j ? (i < 2 ? 1.0 : 0.1) : 0.0; q.Qfs[i][j] = i == j ? 100.0 : 0.0; }
    }
    q.Rs[0][0] = 1.0; q.diag = 1; q.monotone = 1;
    Cost qc(q);
    double h[160];
    for (int i = 0; i < 160; ++i) h[i] = 0.01 * (i % 17) - 0.05;
    double *in, *out;
    cudaMalloc(&in, sizeof h); cudaMalloc(&out, 8ull * B);
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[3] = { "8 sincos", "step without trigonometry", "whole step" };
    for (int part = 0; part < 3; ++part) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (part == 0) parts<0><<<B / 32, 32>>>(sys, qc, N, 0.5, in, out);
            if (part == 1) parts<1><<<B / 32, 32>>>(sys, qc, N, 0.5, in, out);
            if (part == 2) parts<2><<<B / 32, 32>>>(sys, qc, N, 0.5, in, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("part %d (%-26s): %7.3f ms  %7.1f cycles per warp-trip at %d warps per sub-partition  (%s)\n", part, names[part], best,
               best * 1e-3 * 1.965e9 / (N * (double)W), W, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
