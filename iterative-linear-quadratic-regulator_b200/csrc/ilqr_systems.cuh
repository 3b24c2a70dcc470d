// ilqr_systems.cuh -- device-side system definitions (continuous dynamics, analytic Jacobians,
// integrators, quadratic cost) for the batched iLQR kernels.  sm_100a; FP64 on the CUDA cores -- the per-step matrices
// are n<=12, m<=4 (BASELINE.json north_star) -- except the n=12 LTV recursion, which runs on FP64 DMMA
// (ilqr_kernels_ltv_mma.cuh).
//
// What each piece replaces in the reference (paths relative to /root/reference/python/):
//   PendulumSys            class_files/systems/pendulum_sys.py:60-75
//   DoublePendulumSys<M>   class_files/systems/double_pendulum_sys.py:84-111,138-206 (M=2)
//                          class_files/systems/UA_double_pendulum_sys.py:84-111,140-208 (M=1)
//   step<INTEG>            class_files/systems/system_base.py:50-74 (euler/midpoint/rk4), :88-140 (backward Euler)
//   step_jac<INTEG>        system_base.py:203-205 (jacfwd of the step) and :146-188 (IFT for backward Euler),
//                          evaluated in closed form instead of by autodiff
//   QuadCost               pendulum_sys.py:77-98 and the same block in the two double-pendulum files;
//                          derivatives system_base.py:212-219
#pragma once
#ifndef __CUDACC_RTC__      // NVRTC (user-defined systems, class_files/codegen.py) has the device API built in
#include <cuda_runtime.h>
#endif
#include "ilqr_trig_table.cuh"

namespace ilqr {

enum Model { PENDULUM = 0, DOUBLE_PENDULUM = 1, UA_DOUBLE_PENDULUM = 2, LTV = 3 };
enum Integ { EULER = 0, MIDPOINT = 1, RK4 = 2, BACKWARD_EULER = 3 };

#define ILQR_DEV __device__ __forceinline__

// minimax coefficients of the fdlibm sin/cos kernels on [-pi/4, pi/4]; in constant memory so the
// DFMAs take them as c[bank][offset] operands instead of materialising 64-bit immediates
__constant__ double kTrig[12] = {
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

// sin and cos of one FP64 argument sharing a single two-constant Cody-Waite reduction (exact under
// FMA: the first product-difference is rounded once, the second constant carries the next 53 bits of
// pi/2) and the fdlibm minimax kernels on [-pi/4, pi/4].  ~22 FP64 instructions, no slow-path call;
// absolute error ~1e-16 for |x| up to ~1e9 (the generic CUDA sincos() costs about twice the
// instructions because of its Payne-Hanek guard).  NaN/Inf propagate as NaN.
ILQR_DEV void sincos_t(double x, double *s, double *c)
{
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52: rounds to nearest integer
    const double t = fma(x, 0.63661977236758138243, magic);  // x * 2/pi
    const int k = __double2loint(t);
    const double kd = t - magic;
    double r = fma(-kd, 1.57079632679489655800e+00, x);
    r = fma(-kd, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = fma(z, kTrig[0], kTrig[1]);
    ps = fma(z, ps, kTrig[2]);
    ps = fma(z, ps, kTrig[3]);
    ps = fma(z, ps, kTrig[4]);
    ps = fma(z, ps, kTrig[5]);
    double pc = fma(z, kTrig[6], kTrig[7]);
    pc = fma(z, pc, kTrig[8]);
    pc = fma(z, pc, kTrig[9]);
    pc = fma(z, pc, kTrig[10]);
    pc = fma(z, pc, kTrig[11]);
    // operand forms chosen for the FP64 pipe's register ports (a DFMA with three distinct register sources
    // takes 3 issue cycles instead of 2 on sm_100a; scripts/micro/fp64_operands.cu): r + r*(z*ps) reads r twice,
    // and the cosine continues the Horner chain through the constants -1/2 and 1
    const double sn = fma(r, z * ps, r);
    const double cs = fma(z, fma(z, pc, -0.5), 1.0);
    const double a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
    // quadrant signs by flipping the sign bit in the integer pipe (a negate-and-select would spend two FP64
    // pipe slots per sincos on DADDs)
    *s = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
    *c = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
}
// Table form for the hand-written models' FP64 kernels (ILQR_TRIG_TABLE): x = k pi/256 + r, |r| <= pi/512, with
// sin, cos of k pi/256 from a 512-entry table in SHARED memory (trig_table_init; the table is exact to half an ulp, the
// reduction exact as above) and sin r, cos r - 1 from two-term fits: 16 FP64 instructions instead of 22 and no quadrant
// selects; absolute error 1.1e-16 (scripts/gen_trig_table.py), i.e. the same as sincos_t.  The per-lane table read is an
// LDS.128 that runs beside the polynomials.
#ifndef ILQR_TRIG_TABLE
#define ILQR_TRIG_TABLE 1
#endif
#if ILQR_TRIG_TABLE
__shared__ double2 s_trig_table[ILQR_TRIG_N];
// every thread of the block, before the first sincos_tab
ILQR_DEV void trig_table_init()
{
    for (int i = threadIdx.x; i < ILQR_TRIG_N; i += blockDim.x) s_trig_table[i] = g_trig_table[i];
    __syncthreads();
}
ILQR_DEV void sincos_tab(double x, double *s, double *c)
{
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52: rounds to nearest integer
    const double t = fma(x, ILQR_TRIG_INV_H, magic);
    const double2 sc = s_trig_table[__double2loint(t) & (ILQR_TRIG_N - 1)];
    const double kd = t - magic;
    double r = fma(-kd, ILQR_TRIG_H_HI, x);
    r = fma(-kd, ILQR_TRIG_H_LO, r);
    const double z = r * r;
    const double sr = fma(r * z, fma(z, ILQR_TRIG_S2, ILQR_TRIG_S1), r);        // sin r
    const double cm = z * fma(z, ILQR_TRIG_C2, ILQR_TRIG_C1);                   // cos r - 1
    *s = sc.x + fma(sc.x, cm, sc.y * sr);
    *c = sc.y + fma(sc.y, cm, -(sc.x * sr));
}
#endif
// the sincos of a model's dynamics: the table form where the kernel has set the table up (FP64, hand-written models)
template <bool TAB> ILQR_DEV void sincos_m(double x, double *s, double *c)
{
#if ILQR_TRIG_TABLE
    if constexpr (TAB) sincos_tab(x, s, c);
    else
#endif
        sincos_t(x, s, c);
}
// FP32 counterpart (optional 1e-4 mode): three-constant Cody-Waite reduction and the cephes sinf/cosf
// polynomials on [-pi/4, pi/4]; ~2e-7 absolute error for |x| up to ~1e5, no slow path.
ILQR_DEV void sincos_t(float x, float *s, float *c)
{
    const float magic = 12582912.0f;                         // 1.5 * 2^23: rounds to nearest integer
    const float t = fmaf(x, 0.636619772f, magic);
    const int k = __float_as_int(t);                         // the low mantissa bits hold the integer
    const float kd = t - magic;
    float r = fmaf(-kd, 1.5703125f, x);                      // pi/2 split in three parts (cephes DP1..DP3 * 2)
    r = fmaf(-kd, 4.837512969970703125e-4f, r);
    r = fmaf(-kd, 7.54978995489188216e-8f, r);
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    const float sn = fmaf(z * r, ps, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    const float cs = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
    const float a = (k & 1) ? cs : sn, b = (k & 1) ? sn : cs;
    *s = (k & 2) ? -a : a;
    *c = ((k + 1) & 2) ? -b : b;
}
ILQR_DEV double sin_t(double x) { double s, c; sincos_t(x, &s, &c); return s; }
ILQR_DEV float sin_t(float x) { return sinf(x); }
// reciprocal to ~1 ulp: MUFU seed + two Newton steps, no special-case path (the callers' arguments
// are a positive-definite 2x2 determinant or Q_uu, never 0/Inf/denormal in a healthy solve)
ILQR_DEV double rcp_t(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}
ILQR_DEV float rcp_t(float d)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    return fmaf(y, fmaf(-d, y, 1.0f), y);                    // one Newton step: ~1 ulp
}
ILQR_DEV double fma_t(double a, double b, double c) { return fma(a, b, c); }
ILQR_DEV float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
ILQR_DEV double sqrt_t(double x) { return sqrt(x); }
ILQR_DEV float sqrt_t(float x) { return sqrtf(x); }
ILQR_DEV double abs_t(double x) { return fabs(x); }
ILQR_DEV float abs_t(float x) { return fabsf(x); }
// used by generated user systems (class_files/codegen.py)
ILQR_DEV double cos_t(double x) { double s, c; sincos_t(x, &s, &c); return c; }
ILQR_DEV float cos_t(float x) { return cosf(x); }
ILQR_DEV double tan_t(double x) { double s, c; sincos_t(x, &s, &c); return s / c; }
ILQR_DEV float tan_t(float x) { return tanf(x); }
ILQR_DEV double exp_t(double x) { return exp(x); }
ILQR_DEV float exp_t(float x) { return expf(x); }
ILQR_DEV double log_t(double x) { return log(x); }
ILQR_DEV float log_t(float x) { return logf(x); }
ILQR_DEV double tanh_t(double x) { return tanh(x); }
ILQR_DEV float tanh_t(float x) { return tanhf(x); }
ILQR_DEV double pow_t(double x, double y) { return pow(x, y); }
ILQR_DEV float pow_t(float x, float y) { return powf(x, y); }
ILQR_DEV double atan2_t(double y, double x) { return atan2(y, x); }
ILQR_DEV float atan2_t(float y, float x) { return atan2f(y, x); }
ILQR_DEV double atan_t(double x) { return atan(x); }
ILQR_DEV float atan_t(float x) { return atanf(x); }
ILQR_DEV double asin_t(double x) { return asin(x); }
ILQR_DEV float asin_t(float x) { return asinf(x); }
ILQR_DEV double acos_t(double x) { return acos(x); }
ILQR_DEV float acos_t(float x) { return acosf(x); }
ILQR_DEV double sinh_t(double x) { return sinh(x); }
ILQR_DEV float sinh_t(float x) { return sinhf(x); }
ILQR_DEV double cosh_t(double x) { return cosh(x); }
ILQR_DEV float cosh_t(float x) { return coshf(x); }

// ------------------------------------------------------------------------------------------
// Second-order mechanical systems: x = [q, qd], xdot = [qd, qdd(x,u)].  A system provides
//   acc(x,u,a)                 a = qdd
//   acc_jac(x,u,a,J,Bq)        J = d qdd / d x  (NQ x N),  Bq = d qdd / d u  (NQ x M)
// The continuous Jacobian is then A_c = [[0, I],[J]], B_c = [[0],[Bq]], and products with A_c
// only need the NQ dense rows (the zero/identity rows are never multiplied).
// ------------------------------------------------------------------------------------------

template <typename T>
struct PendulumSys {
    static constexpr bool TRIG_TABLE = false;
    static constexpr int NQ = 1, N = 2, M = 1;
    static constexpr bool FIRST_ORDER = false, GENERIC = false;
    T gl, d;   // g/l, damping
    ILQR_DEV T time_scalar(int, T) const { return T(0); }
    ILQR_DEV void acc(const T *x, const T *u, T *a) const
    {
        a[0] = u[0] - d * x[1] - gl * sin_t(x[0]);
    }
    ILQR_DEV void acc_jac(const T *x, const T *u, T *a, T (*J)[N], T (*Bq)[M]) const
    {
        T s, c;
        sincos_t(x[0], &s, &c);
        a[0] = u[0] - d * x[1] - gl * s;
        J[0][0] = -gl * c;
        J[0][1] = -d;
        Bq[0][0] = T(1);
    }
};

// TAB_: sines and cosines through the shared-memory table (FP64 only; the kernels call trig_table_init).  Faster where the
// kernels are bound by the FP64 pipe (B=131072: rollouts 7.08 -> 6.50 ms, fused backward 5.64 -> 5.48 ms per iteration),
// slower where one or two warps per sub-partition wait on the table read's latency (B=4096: rollouts 0.502 -> 0.522 ms):
// a handle uses it from ILQR_TRIG_TABLE_MIN trajectories up (ilqr_b200.cu), in ALL its kernels.
template <typename T, int M_, bool TAB_ = false>
struct DoublePendulumSys {
    static constexpr int NQ = 2, N = 4, M = M_;
    static constexpr bool FIRST_ORDER = false, GENERIC = false;
    static constexpr bool TRIG_TABLE = ILQR_TRIG_TABLE && TAB_ && sizeof(T) == 8;
    ILQR_DEV static void sc(T x, T *s, T *c)
    {
        if constexpr (sizeof(T) == 8) sincos_m<TRIG_TABLE>(x, s, c);
        else sincos_t(x, s, c);
    }
    ILQR_DEV T time_scalar(int, T) const { return T(0); }
    // derived constants (host, double precision):
    //   c = m2 l1 l2, m11_0 = m1 l1^2/4 + m2 l1^2 + m2 l2^2/4 + th1 + th2, m12_0 = m22 = m2 l2^2/4 + th2,
    //   g1 = m2 g l2/2, g2 = m2 g l1 + m1 g l1/2
    T c, m11_0, m12_0, g1, g2, d1, d2;

    ILQR_DEV void acc(const T *x, const T *u, T *a) const
    {
        const T q1d = x[2], q2d = x[3];
        T s1, c1, s2, c2;
        sc(x[0], &s1, &c1);
        sc(x[1], &s2, &c2);
        const T s12 = s1 * c2 + c1 * s2;
        const T m11 = m11_0 + c * c2, m12 = m12_0 + T(0.5) * c * c2, m22 = m12_0;
        const T inv = rcp_t(m11 * m22 - m12 * m12);
        // h1 = u1 + (c s2 / 2)(2 q1' q2' + q2'^2) - g1 s12 - g2 s1 - d1 q1',  h2 = [u2] - (c s2 / 2) q1'^2 - g1 s12 - d2 q2'
        // written so that most FMAs read at most two registers besides a constant (see sincos_t)
        const T hcs2 = (T(0.5) * c) * s2, g = g1 * s12;
        const T w = fma_t(T(2), q1d, q2d) * q2d;
        T h1 = fma_t(hcs2, w, u[0]) - g;
        h1 = fma_t(-g2, s1, h1);
        h1 = fma_t(-d1, q1d, h1);
        T h2 = -fma_t(hcs2, q1d * q1d, g);
        h2 = fma_t(-d2, q2d, h2);
        if (M == 2) h2 += u[M - 1];
        a[0] = inv * (m22 * h1 - m12 * h2);
        a[1] = inv * (m11 * h2 - m12 * h1);
    }

    ILQR_DEV void acc_jac(const T *x, const T *u, T *a, T (*J)[N], T (*Bq)[M]) const
    {
        const T q1d = x[2], q2d = x[3];
        T s1, c1, s2, c2;
        sc(x[0], &s1, &c1);
        sc(x[1], &s2, &c2);
        const T s12 = s1 * c2 + c1 * s2, c12 = c1 * c2 - s1 * s2;
        const T m11 = m11_0 + c * c2, m12 = m12_0 + T(0.5) * c * c2, m22 = m12_0;
        const T inv = rcp_t(m11 * m22 - m12 * m12);
        const T i00 = inv * m22, i01 = -inv * m12, i11 = inv * m11;   // M^-1 (symmetric)
        const T cs2 = c * s2, cc2 = c * c2;
        const T w = T(2) * q1d * q2d + q2d * q2d;
        T h1 = u[0] + T(0.5) * cs2 * w - g1 * s12 - g2 * s1 - d1 * q1d;
        T h2 = -T(0.5) * cs2 * (q1d * q1d) - g1 * s12 - d2 * q2d;
        if (M == 2) h2 += u[M - 1];
        const T a1 = i00 * h1 + i01 * h2, a2 = i01 * h1 + i11 * h2;
        a[0] = a1; a[1] = a2;
        // r_z = dh/dz - (dM/dz) qdd ;  dM/dq2 = -c s2 [[1, 1/2],[1/2, 0]]
        T r0[4], r1[4];
        r0[0] = -g1 * c12 - g2 * c1;
        r1[0] = -g1 * c12;
        r0[1] = T(0.5) * cc2 * w - g1 * c12 + cs2 * (a1 + T(0.5) * a2);
        r1[1] = -T(0.5) * cc2 * (q1d * q1d) - g1 * c12 + T(0.5) * cs2 * a1;
        r0[2] = cs2 * q2d - d1;
        r1[2] = -cs2 * q1d;
        r0[3] = cs2 * (q1d + q2d);
        r1[3] = -d2;
#pragma unroll
        for (int z = 0; z < 4; ++z) {
            J[0][z] = i00 * r0[z] + i01 * r1[z];
            J[1][z] = i01 * r0[z] + i11 * r1[z];
        }
        Bq[0][0] = i00; Bq[1][0] = i01;
        if (M == 2) { Bq[0][M - 1] = i01; Bq[1][M - 1] = i11; }
    }
};

// ------------------------------------------------------------------------------------------
// Synthetic linear time-varying system (BASELINE.json config 4; closest reference
// artefact matlab/CLASSES/Linear_iLQR_CLASS.m): first order, n = 12, m = 4,
//     x_dot = (Ac + w E) x + Bc u,   w = amp * sin(2 pi t / N + phi_b)  (per trajectory phase),
// discretised with forward Euler only.  The matrices are generated in the kernel from the constants,
// never stored per trajectory.
// ------------------------------------------------------------------------------------------
template <typename T>
struct LtvSys {
    static constexpr bool TRIG_TABLE = false;
    static constexpr int NQ = 0, N = 12, M = 4;
    static constexpr bool FIRST_ORDER = true, GENERIC = false;
    T Ac[N][N], E[N][N], Bc[N][M];
    T amp, two_pi_over_N;
    ILQR_DEV T time_scalar(int t, T phi) const { return amp * sin_t(two_pi_over_N * T(t) + phi); }
    ILQR_DEV void xdot(T w, const T *x, const T *u, T *xd) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            T s1 = T(0), s2 = T(0);
#pragma unroll
            for (int j = 0; j < N; ++j) { s1 += Ac[i][j] * x[j]; s2 += E[i][j] * x[j]; }
            T s3 = T(0);
#pragma unroll
            for (int j = 0; j < M; ++j) s3 += Bc[i][j] * u[j];
            xd[i] = (s1 + w * s2) + s3;
        }
    }
};

// ------------------------------------------------------------------------------------------
// small dense LU with partial pivoting, fully unrolled (used by backward Euler and by the
// Q_uu solve for m >= 2).  Rows are swapped by value so all indices stay compile-time.
// ------------------------------------------------------------------------------------------
// FAST_RCP: reciprocals of the pivots by rcp_t (MUFU seed + Newton, ~1 ulp) instead of IEEE division.
template <int n, int nrhs, typename T, bool FAST_RCP = false>
ILQR_DEV void lu_solve_inplace(T (*a)[n], T (*b)[nrhs])
{
#pragma unroll
    for (int j = 0; j < n; ++j) {
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
            // bubble the largest |a[.][j]| of rows j..n-1 into row j (first maximum wins, as getrf)
            const bool sw = abs_t(a[i][j]) > abs_t(a[j][j]);
#pragma unroll
            for (int c = 0; c < n; ++c) { const T u = a[j][c], v = a[i][c]; a[j][c] = sw ? v : u; a[i][c] = sw ? u : v; }
#pragma unroll
            for (int c = 0; c < nrhs; ++c) { const T u = b[j][c], v = b[i][c]; b[j][c] = sw ? v : u; b[i][c] = sw ? u : v; }
        }
        const T r = FAST_RCP ? rcp_t(a[j][j]) : T(1) / a[j][j];
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
            const T l = a[i][j] * r;
#pragma unroll
            for (int c = j + 1; c < n; ++c) a[i][c] -= l * a[j][c];
#pragma unroll
            for (int c = 0; c < nrhs; ++c) b[i][c] -= l * b[j][c];
        }
    }
#pragma unroll
    for (int i = n - 1; i >= 0; --i) {
        const T r = FAST_RCP ? rcp_t(a[i][i]) : T(1) / a[i][i];
#pragma unroll
        for (int c = 0; c < nrhs; ++c) {
            T s = b[i][c];
#pragma unroll
            for (int k = i + 1; k < n; ++k) s -= a[i][k] * b[k][c];
            b[i][c] = s * r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// integrators
// ------------------------------------------------------------------------------------------
template <class Sys, typename T>
ILQR_DEV void f_cont(const Sys &s, const T *x, const T *u, T *xd)
{
    if constexpr (Sys::GENERIC) {
        s.f(x, u, xd);
    } else {
        constexpr int NQ = Sys::NQ;
        T a[NQ];
        s.acc(x, u, a);
#pragma unroll
        for (int i = 0; i < NQ; ++i) { xd[i] = x[NQ + i]; xd[NQ + i] = a[i]; }
    }
}

// ------------------------------------------------------------------------------------------
// Generic first-order systems x_dot = f(x,u) with dense continuous Jacobians -- the form generated
// for user-defined System subclasses (class_files/codegen.py): the system provides
//   f(x,u,xd)                 xd = f(x,u)
//   f_jac(x,u,xd,Ac,Bc)       xd = f(x,u), Ac = df/dx (N x N), Bc = df/du (N x M)
// ------------------------------------------------------------------------------------------

// backward Euler quasi-Newton, system_base.py:101-140 (same scheme as backward_euler_step below)
template <class Sys, typename T>
ILQR_DEV void backward_euler_step_generic(const Sys &s, T dt, const T *x, const T *u, T *xn)
{
    constexpr int n = Sys::N, M = Sys::M;
    T f[n], F[n], Ac[n][n], Bc[n][M], Jr[n][n], Inv[n][n];
    s.f(x, u, f);
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = x[i] + dt * f[i];                     // explicit-Euler guess (:124)
    s.f_jac(xn, u, f, Ac, Bc);                                               // residual Jacobian frozen at the guess (:130-135)
    T fn = T(0);
#pragma unroll
    for (int i = 0; i < n; ++i) { F[i] = xn[i] - x[i] - dt * f[i]; fn += F[i] * F[i]; }
    fn = sqrt_t(fn);
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j) {
            Jr[i][j] = ((i == j) ? T(1) : T(0)) - dt * Ac[i][j];
            Inv[i][j] = (i == j) ? T(1) : T(0);
        }
    lu_solve_inplace<n, n>(Jr, Inv);
    int k = 0;
    while (fn > T(1e-5) && k < 20) {                                         // (:105-120)
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T d = T(0);
#pragma unroll
            for (int j = 0; j < n; ++j) d -= Inv[i][j] * F[j];
            xn[i] += d;
        }
        s.f(xn, u, f);
        fn = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { F[i] = xn[i] - x[i] - dt * f[i]; fn += F[i] * F[i]; }
        fn = sqrt_t(fn);
        ++k;
    }
}

// discrete Jacobians of a generic system: chain rule through the integrator stages (what the reference
// obtains with jacfwd of the step, system_base.py:203-205; implicit function theorem for backward Euler,
// :146-188)
template <int INTEG, class Sys, typename T>
ILQR_DEV void step_jac_generic(const Sys &s, T dt, const T *x, const T *u, T (*A)[Sys::N], T (*Bd)[Sys::M])
{
    constexpr int n = Sys::N, M = Sys::M;
    T k[n], Ac[n][n], Bc[n][M];
    if (INTEG == EULER) {
        s.f_jac(x, u, k, Ac, Bc);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[i][j] = ((i == j) ? T(1) : T(0)) + dt * Ac[i][j];
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = dt * Bc[i][j];
        }
    } else if (INTEG == BACKWARD_EULER) {
        T xn[n], Jr[n][n], R[n][n + M];
        backward_euler_step_generic(s, dt, x, u, xn);
        s.f_jac(xn, u, k, Ac, Bc);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                Jr[i][j] = ((i == j) ? T(1) : T(0)) - dt * Ac[i][j];
                R[i][j] = (i == j) ? T(1) : T(0);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) R[i][n + j] = dt * Bc[i][j];
        }
        lu_solve_inplace<n, n + M>(Jr, R);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[i][j] = R[i][j];
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = R[i][n + j];
        }
    } else {
        constexpr int S = (INTEG == MIDPOINT) ? 2 : 4;
        const T cs[4] = { T(0), T(0.5), (INTEG == RK4) ? T(0.5) : T(0), T(1) };
        const T ws[4] = { (INTEG == RK4) ? T(1) : T(0), (INTEG == RK4) ? T(2) : T(1), T(2), T(1) };
        T Kx[n][n], Ku[n][M], Ax[n][n], Au[n][M], xs[n];
#pragma unroll
        for (int st = 0; st < S; ++st) {
#pragma unroll
            for (int i = 0; i < n; ++i) xs[i] = st == 0 ? x[i] : x[i] + (cs[st] * dt) * k[i];
            s.f_jac(xs, u, k, Ac, Bc);
            if (st == 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Kx[i][j] = Ac[i][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Ku[i][j] = Bc[i][j];
                }
            } else {
                // Kx <- Ac (I + c dt Kx), Ku <- Ac (c dt Ku) + Bc
                T Sx[n][n], Su[n][M];
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Sx[i][j] = ((i == j) ? T(1) : T(0)) + (cs[st] * dt) * Kx[i][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Su[i][j] = (cs[st] * dt) * Ku[i][j];
                }
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) {
                        T acc = T(0);
#pragma unroll
                        for (int l = 0; l < n; ++l) acc += Ac[i][l] * Sx[l][j];
                        Kx[i][j] = acc;
                    }
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        T acc = T(0);
#pragma unroll
                        for (int l = 0; l < n; ++l) acc += Ac[i][l] * Su[l][j];
                        Ku[i][j] = acc + Bc[i][j];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j < n; ++j) Ax[i][j] = st == 0 ? ws[0] * Kx[i][j] : Ax[i][j] + ws[st] * Kx[i][j];
#pragma unroll
                for (int j = 0; j < M; ++j) Au[i][j] = st == 0 ? ws[0] * Ku[i][j] : Au[i][j] + ws[st] * Ku[i][j];
            }
        }
        const T h = (INTEG == RK4) ? dt / T(6) : dt;
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[i][j] = ((i == j) ? T(1) : T(0)) + h * Ax[i][j];
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = h * Au[i][j];
        }
    }
}

// backward Euler quasi-Newton (system_base.py:101-140): explicit-Euler guess, Jacobian of the
// residual frozen at the guess, iterate while ||F||_2 > 1e-5 and k < 20.
template <class Sys, typename T>
ILQR_DEV void backward_euler_step(const Sys &s, T dt, const T *x, const T *u, T *xn)
{
    constexpr int n = Sys::N, NQ = Sys::NQ, M = Sys::M;
    T f[n], F[n][1], a[NQ], J[NQ][n], Bq[NQ][M], Jr[n][n];
    f_cont(s, x, u, f);
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = x[i] + dt * f[i];
    s.acc_jac(xn, u, a, J, Bq);
    T fn = T(0);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        F[i][0] = xn[i] - x[i] - dt * xn[NQ + i];
        F[NQ + i][0] = xn[NQ + i] - x[NQ + i] - dt * a[i];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) fn += F[i][0] * F[i][0];
    fn = sqrt_t(fn);
    // residual Jacobian I - dt*A_c at the guess; inverted once (n<=4) instead of an LU object
    T Inv[n][n];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const T ac = (i < NQ) ? ((j == i + NQ) ? T(1) : T(0)) : J[i - NQ][j];
            Jr[i][j] = ((i == j) ? T(1) : T(0)) - dt * ac;
            Inv[i][j] = (i == j) ? T(1) : T(0);
        }
    lu_solve_inplace<n, n>(Jr, Inv);
    int k = 0;
    while (fn > T(1e-5) && k < 20) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T d = T(0);
#pragma unroll
            for (int j = 0; j < n; ++j) d -= Inv[i][j] * F[j][0];
            xn[i] += d;
        }
        s.acc(xn, u, a);
        fn = T(0);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            F[i][0] = xn[i] - x[i] - dt * xn[NQ + i];
            F[NQ + i][0] = xn[NQ + i] - x[NQ + i] - dt * a[i];
        }
#pragma unroll
        for (int i = 0; i < n; ++i) fn += F[i][0] * F[i][0];
        fn = sqrt_t(fn);
        ++k;
    }
}

template <int INTEG, class Sys, typename T>
ILQR_DEV void step(const Sys &s, T dt, const T *x, const T *u, T *xn, T w = T(0))
{
    constexpr int n = Sys::N;
    if constexpr (Sys::FIRST_ORDER) {
        T xd[n];
        s.xdot(w, x, u, xd);
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = x[i] + xd[i] * dt;
    } else
    if (INTEG == EULER) {
        T k1[n];
        f_cont(s, x, u, k1);
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = x[i] + k1[i] * dt;
    } else if (INTEG == MIDPOINT) {
        T k1[n], k2[n], xs[n];
        f_cont(s, x, u, k1);
#pragma unroll
        for (int i = 0; i < n; ++i) xs[i] = x[i] + (dt * T(0.5)) * k1[i];
        f_cont(s, xs, u, k2);
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = x[i] + dt * k2[i];
    } else if (INTEG == RK4) {
        T k1[n], k2[n], k3[n], k4[n], xs[n];
        f_cont(s, x, u, k1);
#pragma unroll
        for (int i = 0; i < n; ++i) xs[i] = x[i] + (dt * T(0.5)) * k1[i];
        f_cont(s, xs, u, k2);
#pragma unroll
        for (int i = 0; i < n; ++i) xs[i] = x[i] + (dt * T(0.5)) * k2[i];
        f_cont(s, xs, u, k3);
#pragma unroll
        for (int i = 0; i < n; ++i) xs[i] = x[i] + dt * k3[i];
        f_cont(s, xs, u, k4);
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = x[i] + (dt / T(6)) * (k1[i] + T(2) * k2[i] + T(2) * k3[i] + k4[i]);
    } else {
        if constexpr (Sys::GENERIC) backward_euler_step_generic(s, dt, x, u, xn);
        else backward_euler_step(s, dt, x, u, xn);
    }
}

// out = A_c * S for a second-order system: rows 0..NQ-1 are rows NQ..N-1 of S, the rest J*S
template <int NQ, int n, int cols, typename T>
ILQR_DEV void mul_Ac(const T (*J)[n], const T (*S)[cols], T (*out)[cols])
{
#pragma unroll
    for (int c = 0; c < cols; ++c) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) out[i][c] = S[NQ + i][c];
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < n; ++k) acc += J[i][k] * S[k][c];
            out[NQ + i][c] = acc;
        }
    }
}

// discrete Jacobians A = d step/dx (n x n), Bd = d step/du (n x m)
template <int INTEG, class Sys, typename T>
ILQR_DEV void step_jac(const Sys &s, T dt, const T *x, const T *u, T (*A)[Sys::N], T (*Bd)[Sys::M], T w = T(0))
{
    constexpr int n = Sys::N, NQ = Sys::NQ, M = Sys::M;
    if constexpr (Sys::GENERIC) {
        step_jac_generic<INTEG>(s, dt, x, u, A, Bd);
    } else if constexpr (Sys::FIRST_ORDER) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[i][j] = ((i == j) ? T(1) : T(0)) + dt * (s.Ac[i][j] + w * s.E[i][j]);
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = dt * s.Bc[i][j];
        }
    } else {
    T a[NQ], J[NQ][n], Bq[NQ][M];
    if (INTEG == EULER) {
        s.acc_jac(x, u, a, J, Bq);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const T ac = (i < NQ) ? ((j == i + NQ) ? T(1) : T(0)) : J[i - NQ][j];
                A[i][j] = ((i == j) ? T(1) : T(0)) + dt * ac;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = (i < NQ) ? T(0) : dt * Bq[i - NQ][j];
        }
    } else if (INTEG == BACKWARD_EULER) {
        // IFT at the converged step: A = (I - dt A_c(x+))^-1, B = A dt B_c(x+)
        T xn[n], Jr[n][n], R[n][n + M];
        backward_euler_step(s, dt, x, u, xn);
        s.acc_jac(xn, u, a, J, Bq);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const T ac = (i < NQ) ? ((j == i + NQ) ? T(1) : T(0)) : J[i - NQ][j];
                Jr[i][j] = ((i == j) ? T(1) : T(0)) - dt * ac;
                R[i][j] = (i == j) ? T(1) : T(0);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) R[i][n + j] = (i < NQ) ? T(0) : dt * Bq[i - NQ][j];
        }
        lu_solve_inplace<n, n + M>(Jr, R);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[i][j] = R[i][j];
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = R[i][n + j];
        }
    } else {
        // explicit Runge-Kutta stages with ZOH on u:
        //   Kx_s = A_c(x_s) (I + c_s dt Kx_{s-1}),  Ku_s = A_c(x_s) c_s dt Ku_{s-1} + B_c(x_s)
        constexpr int S = (INTEG == MIDPOINT) ? 2 : 4;
        const T cs[4] = { T(0), T(0.5), (INTEG == RK4) ? T(0.5) : T(0), T(1) };
        const T ws[4] = { (INTEG == RK4) ? T(1) : T(0), (INTEG == RK4) ? T(2) : T(1), T(2), T(1) };
        T kprev[n], Kx[n][n], Ku[n][M], Ax[n][n], Au[n][M], xs[n];
#pragma unroll
        for (int st = 0; st < S; ++st) {
            if (st == 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) xs[i] = x[i];
            } else {
#pragma unroll
                for (int i = 0; i < n; ++i) xs[i] = x[i] + (cs[st] * dt) * kprev[i];
            }
            s.acc_jac(xs, u, a, J, Bq);
#pragma unroll
            for (int i = 0; i < NQ; ++i) { kprev[i] = xs[NQ + i]; kprev[NQ + i] = a[i]; }
            if (st == 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Kx[i][j] = (i < NQ) ? ((j == i + NQ) ? T(1) : T(0)) : J[i - NQ][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Ku[i][j] = (i < NQ) ? T(0) : Bq[i - NQ][j];
                }
            } else {
                T Sx[n][n], Su[n][M], Nx[n][n], Nu[n][M];
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Sx[i][j] = ((i == j) ? T(1) : T(0)) + (cs[st] * dt) * Kx[i][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Su[i][j] = (cs[st] * dt) * Ku[i][j];
                }
                mul_Ac<NQ, n, n>(J, Sx, Nx);
                mul_Ac<NQ, n, M>(J, Su, Nu);
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Kx[i][j] = Nx[i][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Ku[i][j] = Nu[i][j] + ((i < NQ) ? T(0) : Bq[i - NQ][j]);
                }
            }
            if (st == 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Ax[i][j] = ws[0] * Kx[i][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Au[i][j] = ws[0] * Ku[i][j];
                }
            } else {
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                    for (int j = 0; j < n; ++j) Ax[i][j] += ws[st] * Kx[i][j];
#pragma unroll
                    for (int j = 0; j < M; ++j) Au[i][j] += ws[st] * Ku[i][j];
                }
            }
        }
        const T h = (INTEG == RK4) ? dt / T(6) : dt;
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[i][j] = ((i == j) ? T(1) : T(0)) + h * Ax[i][j];
#pragma unroll
            for (int j = 0; j < M; ++j) Bd[i][j] = h * Au[i][j];
        }
    }
    }
}

// ------------------------------------------------------------------------------------------
// quadratic cost  l = (1/2 dx'Q dx + 1/2 u'R u) dt,  l_f = 1/2 dx'Q_f dx,  dx = x - x_target.
// Qs/Rs/Qfs are the symmetrised weights (the derivatives the reference gets from autodiff are
// those of the symmetric part).
// ------------------------------------------------------------------------------------------
template <typename T, int n, int m>
struct QuadCost {
    static constexpr bool QUADRATIC = true;
    T dt;
    T xt[n];
    T Qs[n][n], Rs[m][m], Qfs[n][n];
    int diag;       // all three weights diagonal (true for every reference script)
    int monotone;   // diagonal and non-negative: running cost sums are monotone (enables early rejection)

    ILQR_DEV T stage(const T *x, const T *u) const
    {
        T cx = T(0), cu = T(0);
        if (diag) {
#pragma unroll
            for (int i = 0; i < n; ++i) { const T d = x[i] - xt[i]; cx += (T(0.5) * d) * Qs[i][i] * d; }
#pragma unroll
            for (int j = 0; j < m; ++j) cu += (T(0.5) * u[j]) * Rs[j][j] * u[j];
        } else {
            T dx[n];
#pragma unroll
            for (int i = 0; i < n; ++i) dx[i] = x[i] - xt[i];
#pragma unroll
            for (int j = 0; j < n; ++j) {
                T s = T(0);
#pragma unroll
                for (int i = 0; i < n; ++i) s += (T(0.5) * dx[i]) * Qs[i][j];
                cx += s * dx[j];
            }
#pragma unroll
            for (int j = 0; j < m; ++j) {
                T s = T(0);
#pragma unroll
                for (int i = 0; i < m; ++i) s += (T(0.5) * u[i]) * Rs[i][j];
                cu += s * u[j];
            }
        }
        return (cx + cu) * dt;
    }
    ILQR_DEV T terminal(const T *x) const
    {
        T dx[n], c = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) dx[i] = x[i] - xt[i];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            T s = T(0);
#pragma unroll
            for (int i = 0; i < n; ++i) s += (T(0.5) * dx[i]) * Qfs[i][j];
            c += s * dx[j];
        }
        return c;
    }
    // l_x = dt Qs dx, l_u = dt Rs u
    ILQR_DEV void grad(const T *x, const T *u, T *lx, T *lu) const
    {
        T dx[n];
#pragma unroll
        for (int i = 0; i < n; ++i) dx[i] = x[i] - xt[i];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T s = T(0);
            if (diag) s = Qs[i][i] * dx[i];
            else {
#pragma unroll
                for (int j = 0; j < n; ++j) s += Qs[i][j] * dx[j];
            }
            lx[i] = s * dt;
        }
#pragma unroll
        for (int i = 0; i < m; ++i) {
            T s = T(0);
            if (diag) s = Rs[i][i] * u[i];
            else {
#pragma unroll
                for (int j = 0; j < m; ++j) s += Rs[i][j] * u[j];
            }
            lu[i] = s * dt;
        }
    }
    ILQR_DEV void terminal_grad(const T *x, T *vx) const
    {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < n; ++j) s += Qfs[i][j] * (x[j] - xt[j]);
            vx[i] = s;
        }
    }
};

// The rollout kernel's view of a QuadCost whose three weights are diagonal (every reference script): the
// diagonals only, so the kernel keeps ~13 constants in uniform registers instead of three dense matrices,
// and the stage cost has no runtime branch.  Same expressions as QuadCost::stage / ::terminal with diag set.
template <typename T, int n, int m>
struct DiagCost {
    static constexpr bool QUADRATIC = true;
    T dt;
    T xt[n], hq[n], hr[m], hqf[n];   // halved diagonals: (0.5 d) q d == ((0.5 q) d) d bit for bit (scaling by 1/2 is exact)
    int monotone;
#ifndef __CUDACC_RTC__     // host-side conversion (launch_rollout); NVRTC compiles device code only
    explicit DiagCost(const QuadCost<T, n, m> &c) : dt(c.dt), monotone(c.monotone)
    {
        for (int i = 0; i < n; ++i) { xt[i] = c.xt[i]; hq[i] = T(0.5) * c.Qs[i][i]; hqf[i] = T(0.5) * c.Qfs[i][i]; }
        for (int j = 0; j < m; ++j) hr[j] = T(0.5) * c.Rs[j][j];
    }
#endif
    ILQR_DEV T stage(const T *x, const T *u) const
    {
        T cx = T(0), cu = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { const T d = x[i] - xt[i]; cx += (hq[i] * d) * d; }
#pragma unroll
        for (int j = 0; j < m; ++j) cu += (hr[j] * u[j]) * u[j];
        return (cx + cu) * dt;
    }
    ILQR_DEV T terminal(const T *x) const
    {
        T c = T(0);
#pragma unroll
        for (int j = 0; j < n; ++j) { const T d = x[j] - xt[j]; c += (hqf[j] * d) * d; }
        return c;
    }
};


}  // namespace ilqr
