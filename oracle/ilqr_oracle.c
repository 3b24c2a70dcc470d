/* ilqr_oracle.c -- TEST INFRASTRUCTURE ONLY.  See ilqr_oracle.h for the header note.
 *
 * Plain-C float64 restatement of the reference hot path.  Citations are relative to
 * /root/reference/python/.  The reference obtains f_x/f_u and the cost derivatives by
 * autodiff (class_files/systems/system_base.py:203-219); this file evaluates the same
 * derivatives in closed form, which is equal up to rounding and is
 * checked against the reference's own autodiff outputs in tests/golden/derivs_*.npz.
 * Compile with -ffp-contract=off so the operation sequence is the one written here.
 */
#include "ilqr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define NX ORC_NMAX
#define MX ORC_MMAX

/* ------------------------------------------------------------------ small dense helpers */

/* LU with partial pivoting (LAPACK getrf semantics: first largest |a_ij| in the column),
 * in place on row-major a (n x n).  This is what jnp.linalg.solve / lu_factor call. */
static void lu_factor(int n, double *a, int *piv)
{
    for (int j = 0; j < n; ++j) {
        int p = j;
        double best = fabs(a[j * n + j]);
        for (int i = j + 1; i < n; ++i) {
            double v = fabs(a[i * n + j]);
            if (v > best) { best = v; p = i; }
        }
        piv[j] = p;
        if (p != j)
            for (int c = 0; c < n; ++c) { double t = a[j * n + c]; a[j * n + c] = a[p * n + c]; a[p * n + c] = t; }
        double d = a[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double l = a[i * n + j] / d;
            a[i * n + j] = l;
            for (int c = j + 1; c < n; ++c) a[i * n + c] -= l * a[j * n + c];
        }
    }
}

/* solve for nrhs right-hand sides stored row-major b (n x nrhs), in place */
static void lu_solve(int n, const double *lu, const int *piv, double *b, int nrhs)
{
    for (int j = 0; j < n; ++j) {
        int p = piv[j];
        if (p != j)
            for (int c = 0; c < nrhs; ++c) { double t = b[j * nrhs + c]; b[j * nrhs + c] = b[p * nrhs + c]; b[p * nrhs + c] = t; }
    }
    for (int c = 0; c < nrhs; ++c) {
        for (int i = 1; i < n; ++i) {
            double s = b[i * nrhs + c];
            for (int k = 0; k < i; ++k) s -= lu[i * n + k] * b[k * nrhs + c];
            b[i * nrhs + c] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = b[i * nrhs + c];
            for (int k = i + 1; k < n; ++k) s -= lu[i * n + k] * b[k * nrhs + c];
            b[i * nrhs + c] = s / lu[i * n + i];
        }
    }
}

/* C(r x c) = A(r x k) * B(k x c) */
static void matmul(int r, int k, int c, const double *A, const double *B, double *C)
{
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * c + j];
            C[i * c + j] = s;
        }
}

/* C(k x c) = A^T (A is r x k) * B(r x c) */
static void matmul_tn(int r, int k, int c, const double *A, const double *B, double *C)
{
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < c; ++j) {
            double s = 0.0;
            for (int l = 0; l < r; ++l) s += A[l * k + i] * B[l * c + j];
            C[i * c + j] = s;
        }
}

/* ------------------------------------------------------------------ continuous dynamics */

typedef struct { double m11, m12, m22, h1, h2, s1, s2, s12, c1, c2, c12, qdd1, qdd2; int lu[2]; double f[4]; } dp_eval;

/* double_pendulum_sys.py:138-160 (mass matrix), :162-206 (rhs), :107 (2x2 LU solve);
 * UA_double_pendulum_sys.py:140-162, :164-208 (f_act = [tau[0], 0], :204). */
static void dp_eval_point(const orc_problem *p, const double *x, const double *u, dp_eval *e)
{
    const double g = p->phys[0], m1 = p->phys[1], m2 = p->phys[2], l1 = p->phys[3], l2 = p->phys[4];
    const double d1 = p->phys[5], d2 = p->phys[6], th1 = p->phys[7], th2 = p->phys[8];
    const double q1 = x[0], q2 = x[1], q1d = x[2], q2d = x[3];
    e->s1 = sin(q1); e->s2 = sin(q2); e->s12 = sin(q1 + q2);
    e->c1 = cos(q1); e->c2 = cos(q2); e->c12 = cos(q1 + q2);
    e->m11 = (m1 * (l1 * l1)) / 4 + m2 * (l1 * l1) + (m2 * (l2 * l2)) / 4 + m2 * l1 * l2 * e->c2 + th1 + th2;
    e->m12 = (m2 * (l2 * l2)) / 4 + (m2 * l1 * l2 * e->c2) / 2 + th2;
    e->m22 = (m2 * (l2 * l2)) / 4 + th2;
    double fc1 = (m2 * l1 * l2 * e->s2 * (2 * q1d * q2d + q2d * q2d)) / 2;
    double fc2 = -(m2 * l1 * l2 * e->s2 * (q1d * q1d)) / 2;
    double fg1 = -m2 * g * (l2 * e->s12 / 2 + l1 * e->s1) - (m1 * g * l1 * e->s1) / 2;
    double fg2 = -m2 * g * (l2 * e->s12) / 2;
    double fd1 = -d1 * q1d, fd2 = -d2 * q2d;
    double fa1 = u[0], fa2 = (p->model == ORC_DOUBLE_PENDULUM) ? u[1] : 0.0;
    e->h1 = fa1 + fc1 + fg1 + fd1;
    e->h2 = fa2 + fc2 + fg2 + fd2;
    double M[4] = { e->m11, e->m12, e->m12, e->m22 };
    lu_factor(2, M, e->lu);
    memcpy(e->f, M, sizeof M);
    double b[2] = { e->h1, e->h2 };
    lu_solve(2, e->f, e->lu, b, 1);
    e->qdd1 = b[0]; e->qdd2 = b[1];
}

static void ltv_matrix(const orc_problem *p, int t, double phi, double *A)
{
    const int n = p->n;
    const double w = p->ltv_amp * sin(2.0 * M_PI * (double)t / (double)p->N + phi);
    for (int i = 0; i < n * n; ++i) A[i] = p->Ac[i] + w * p->E[i];
}

void orc_f_cont(const orc_problem *p, int t, double phi, const double *x, const double *u, double *xdot)
{
    switch (p->model) {
    case ORC_PENDULUM: {   /* pendulum_sys.py:60-75 */
        const double g = p->phys[0], l = p->phys[1], d = p->phys[2];
        xdot[0] = x[1];
        xdot[1] = u[0] - d * x[1] - (g / l) * sin(x[0]);
        break;
    }
    case ORC_DOUBLE_PENDULUM:
    case ORC_UA_DOUBLE_PENDULUM: {   /* double_pendulum_sys.py:84-111 */
        dp_eval e;
        dp_eval_point(p, x, u, &e);
        xdot[0] = x[2]; xdot[1] = x[3]; xdot[2] = e.qdd1; xdot[3] = e.qdd2;
        break;
    }
    default: {   /* synthetic LTV, BASELINE config 4 */
        double A[NX * NX], t1[NX], t2[NX];
        ltv_matrix(p, t, phi, A);
        matmul(p->n, p->n, 1, A, x, t1);
        matmul(p->n, p->m, 1, p->Bc, u, t2);
        for (int i = 0; i < p->n; ++i) xdot[i] = t1[i] + t2[i];
    }
    }
}

/* closed form of jacfwd(_f_cont_fcn) (system_base.py:209-210); SURVEY.md Appendix B */
void orc_f_cont_jac(const orc_problem *p, int t, double phi, const double *x, const double *u,
                    double *Ac, double *Bc)
{
    const int n = p->n, m = p->m;
    memset(Ac, 0, sizeof(double) * n * n);
    memset(Bc, 0, sizeof(double) * n * m);
    switch (p->model) {
    case ORC_PENDULUM: {
        const double g = p->phys[0], l = p->phys[1], d = p->phys[2];
        Ac[0 * 2 + 1] = 1.0;
        Ac[1 * 2 + 0] = -(g / l) * cos(x[0]);
        Ac[1 * 2 + 1] = -d;
        Bc[1] = 1.0;
        break;
    }
    case ORC_DOUBLE_PENDULUM:
    case ORC_UA_DOUBLE_PENDULUM: {
        const double g = p->phys[0], m1 = p->phys[1], m2 = p->phys[2], l1 = p->phys[3], l2 = p->phys[4];
        const double d1 = p->phys[5], d2 = p->phys[6];
        const double q1d = x[2], q2d = x[3];
        dp_eval e;
        dp_eval_point(p, x, u, &e);
        const double c = m2 * l1 * l2;
        /* rhs_z = dh/dz - (dM/dz) qdd, columns z = q1,q2,q1d,q2d, then unit torque columns */
        double rhs[2 * 6];
        const int nr = 4 + m;
        double r0[6], r1[6];
        r0[0] = -m2 * g * (l2 * e.c12 / 2 + l1 * e.c1) - (m1 * g * l1 * e.c1) / 2;
        r1[0] = -m2 * g * (l2 * e.c12) / 2;
        r0[1] = (c * e.c2 * (2 * q1d * q2d + q2d * q2d)) / 2 - m2 * g * (l2 * e.c12) / 2
                + c * e.s2 * (e.qdd1 + e.qdd2 / 2);
        r1[1] = -(c * e.c2 * (q1d * q1d)) / 2 - m2 * g * (l2 * e.c12) / 2 + c * e.s2 * (e.qdd1 / 2);
        r0[2] = c * e.s2 * q2d - d1;
        r1[2] = -c * e.s2 * q1d;
        r0[3] = c * e.s2 * (q1d + q2d);
        r1[3] = -d2;
        r0[4] = 1.0; r1[4] = 0.0;
        r0[5] = 0.0; r1[5] = 1.0;
        for (int k = 0; k < nr; ++k) { rhs[0 * nr + k] = r0[k]; rhs[1 * nr + k] = r1[k]; }
        lu_solve(2, e.f, e.lu, rhs, nr);
        Ac[0 * 4 + 2] = 1.0; Ac[1 * 4 + 3] = 1.0;
        for (int k = 0; k < 4; ++k) { Ac[2 * 4 + k] = rhs[0 * nr + k]; Ac[3 * 4 + k] = rhs[1 * nr + k]; }
        for (int k = 0; k < m; ++k) { Bc[2 * m + k] = rhs[0 * nr + 4 + k]; Bc[3 * m + k] = rhs[1 * nr + 4 + k]; }
        break;
    }
    default:
        ltv_matrix(p, t, phi, Ac);
        memcpy(Bc, p->Bc, sizeof(double) * n * m);
    }
}

/* ------------------------------------------------------------------ discrete step */

/* system_base.py:88-140.  Returns the quasi-Newton trip count. */
static int backward_euler_step(const orc_problem *p, int t, double phi, const double *x, const double *u, double *xn)
{
    const int n = p->n;
    double f[NX], F[NX], J[NX * NX], Bc[NX * MX], fn;
    int piv[NX];
    orc_f_cont(p, t, phi, x, u, f);
    for (int i = 0; i < n; ++i) xn[i] = x[i] + p->dt * f[i];            /* :124 explicit-Euler guess */
    orc_f_cont(p, t, phi, xn, u, f);
    fn = 0.0;
    for (int i = 0; i < n; ++i) { F[i] = xn[i] - x[i] - p->dt * f[i]; fn += F[i] * F[i]; }   /* :101-103 */
    fn = sqrt(fn);                                                       /* :127 */
    orc_f_cont_jac(p, t, phi, xn, u, J, Bc);                             /* :130 stale Jacobian at the guess */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) J[i * n + j] = (i == j ? 1.0 : 0.0) - p->dt * J[i * n + j];   /* :131 */
    lu_factor(n, J, piv);                                                /* :135 */
    int k = 0;
    while (fn > 1e-5 && k < 20) {                                        /* :105-107 */
        double d[NX];
        for (int i = 0; i < n; ++i) d[i] = -F[i];
        lu_solve(n, J, piv, d, 1);                                       /* :113 */
        for (int i = 0; i < n; ++i) xn[i] = xn[i] + d[i];                /* :115 */
        orc_f_cont(p, t, phi, xn, u, f);
        fn = 0.0;
        for (int i = 0; i < n; ++i) { F[i] = xn[i] - x[i] - p->dt * f[i]; fn += F[i] * F[i]; }
        fn = sqrt(fn);
        ++k;
    }
    return k;
}

int orc_f(const orc_problem *p, int t, double phi, const double *x, const double *u, double *xn)
{
    const int n = p->n;
    const double dt = p->dt;
    double k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
    switch (p->integrator) {
    case ORC_EULER:      /* system_base.py:50-53 */
        orc_f_cont(p, t, phi, x, u, k1);
        for (int i = 0; i < n; ++i) xn[i] = x[i] + k1[i] * dt;
        return 0;
    case ORC_MIDPOINT:   /* system_base.py:55-63 */
        orc_f_cont(p, t, phi, x, u, k1);
        for (int i = 0; i < n; ++i) xs[i] = x[i] + (dt / 2.0) * k1[i];
        orc_f_cont(p, t, phi, xs, u, k2);
        for (int i = 0; i < n; ++i) xn[i] = x[i] + dt * k2[i];
        return 0;
    case ORC_RK4:        /* system_base.py:65-74 */
        orc_f_cont(p, t, phi, x, u, k1);
        for (int i = 0; i < n; ++i) xs[i] = x[i] + dt / 2 * k1[i];
        orc_f_cont(p, t, phi, xs, u, k2);
        for (int i = 0; i < n; ++i) xs[i] = x[i] + dt / 2 * k2[i];
        orc_f_cont(p, t, phi, xs, u, k3);
        for (int i = 0; i < n; ++i) xs[i] = x[i] + dt * k3[i];
        orc_f_cont(p, t, phi, xs, u, k4);
        for (int i = 0; i < n; ++i) xn[i] = x[i] + (dt / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
        return 0;
    default:
        return backward_euler_step(p, t, phi, x, u, xn);
    }
}

/* Discrete Jacobians: chain rule through the integrators above == jacfwd(self._f_fcn)
 * (system_base.py:203-205); backward Euler by the implicit-function theorem
 * (system_base.py:146-188). */
void orc_f_jac(const orc_problem *p, int t, double phi, const double *x, const double *u, double *A, double *B)
{
    const int n = p->n, m = p->m;
    const double dt = p->dt;
    double Ac[NX * NX], Bc[NX * MX];
    if (p->integrator == ORC_EULER) {
        orc_f_cont_jac(p, t, phi, x, u, Ac, Bc);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 1.0 : 0.0) + dt * Ac[i * n + j];
        for (int i = 0; i < n * m; ++i) B[i] = dt * Bc[i];
        return;
    }
    if (p->integrator == ORC_BACKWARD_EULER) {
        double xn[NX], J[NX * NX], rhs[NX * (NX + MX)];
        int piv[NX];
        backward_euler_step(p, t, phi, x, u, xn);                        /* :149 / :170 */
        orc_f_cont_jac(p, t, phi, xn, u, Ac, Bc);                        /* :153, :175-176 */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) J[i * n + j] = (i == j ? 1.0 : 0.0) - dt * Ac[i * n + j];
        const int nr = n + m;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) rhs[i * nr + j] = (i == j ? 1.0 : 0.0);       /* -J_x = I, :160-164 */
            for (int j = 0; j < m; ++j) rhs[i * nr + n + j] = dt * Bc[i * m + j];     /* -J_u, :183-187 */
        }
        lu_factor(n, J, piv);
        lu_solve(n, J, piv, rhs, nr);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) A[i * n + j] = rhs[i * nr + j];
            for (int j = 0; j < m; ++j) B[i * m + j] = rhs[i * nr + n + j];
        }
        return;
    }
    /* midpoint / rk4: stage sensitivities  Kx_s = Ac(x_s)(I + c_s dt Kx_{s-1}),
     *                                      Ku_s = Ac(x_s) c_s dt Ku_{s-1} + Bc(x_s)  */
    const int stages = (p->integrator == ORC_MIDPOINT) ? 2 : 4;
    const double cs[4] = { 0.0, 0.5, (p->integrator == ORC_RK4) ? 0.5 : 0.0, 1.0 };
    double k[4][NX], Kx[4][NX * NX], Ku[4][NX * MX], xs[NX], T1[NX * NX], T2[NX * MX];
    for (int s = 0; s < stages; ++s) {
        if (s == 0) memcpy(xs, x, sizeof(double) * n);
        else for (int i = 0; i < n; ++i) xs[i] = x[i] + cs[s] * dt * k[s - 1][i];
        orc_f_cont(p, t, phi, xs, u, k[s]);
        orc_f_cont_jac(p, t, phi, xs, u, Ac, Bc);
        if (s == 0) {
            memcpy(Kx[0], Ac, sizeof(double) * n * n);
            memcpy(Ku[0], Bc, sizeof(double) * n * m);
        } else {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) T1[i * n + j] = (i == j ? 1.0 : 0.0) + cs[s] * dt * Kx[s - 1][i * n + j];
            matmul(n, n, n, Ac, T1, Kx[s]);
            for (int i = 0; i < n * m; ++i) T2[i] = cs[s] * dt * Ku[s - 1][i];
            matmul(n, n, m, Ac, T2, Ku[s]);
            for (int i = 0; i < n * m; ++i) Ku[s][i] += Bc[i];
        }
    }
    if (stages == 2) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 1.0 : 0.0) + dt * Kx[1][i * n + j];
        for (int i = 0; i < n * m; ++i) B[i] = dt * Ku[1][i];
    } else {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                A[i * n + j] = (i == j ? 1.0 : 0.0)
                    + (dt / 6.0) * (Kx[0][i * n + j] + 2 * Kx[1][i * n + j] + 2 * Kx[2][i * n + j] + Kx[3][i * n + j]);
        for (int i = 0; i < n * m; ++i) B[i] = (dt / 6.0) * (Ku[0][i] + 2 * Ku[1][i] + 2 * Ku[2][i] + Ku[3][i]);
    }
}

/* ------------------------------------------------------------------ cost */

static double quad_half(int n, const double *M, const double *v)
{
    /* 0.5 * v.T @ M @ v evaluated as (0.5*v) @ M @ v: row-vector times matrix, then dot */
    double acc = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += (0.5 * v[i]) * M[i * n + j];
        acc += s * v[j];
    }
    return acc;
}

/* pendulum_sys.py:77-90, double_pendulum_sys.py:114-126, UA_double_pendulum_sys.py:114-128 */
double orc_l(const orc_problem *p, const double *x, const double *u)
{
    double dx[NX];
    for (int i = 0; i < p->n; ++i) dx[i] = x[i] - p->x_target[i];
    double cx = quad_half(p->n, p->Q, dx);
    double cu = quad_half(p->m, p->R, u);
    return (cx + cu) * p->dt;
}

/* pendulum_sys.py:92-98 etc. */
double orc_lf(const orc_problem *p, const double *x)
{
    double dx[NX];
    for (int i = 0; i < p->n; ++i) dx[i] = x[i] - p->x_target[i];
    return quad_half(p->n, p->Qf, dx);
}

/* closed form of grad/hessian/jacfwd(grad) of the quadratic (system_base.py:212-216):
 * l_x = dt*sym(Q) dx, l_u = dt*sym(R) u, l_xx = dt*sym(Q), l_uu = dt*sym(R), l_ux = 0 */
void orc_l_derivs(const orc_problem *p, const double *x, const double *u,
                  double *lx, double *lu, double *lxx, double *luu, double *lux)
{
    const int n = p->n, m = p->m;
    double dx[NX];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - p->x_target[i];
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) {
            double q = 0.5 * (p->Q[i * n + j] + p->Q[j * n + i]);
            lxx[i * n + j] = q * p->dt;
            s += q * dx[j];
        }
        lx[i] = s * p->dt;
    }
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) {
            double r = 0.5 * (p->R[i * m + j] + p->R[j * m + i]);
            luu[i * m + j] = r * p->dt;
            s += r * u[j];
        }
        lu[i] = s * p->dt;
    }
    memset(lux, 0, sizeof(double) * m * n);
}

/* system_base.py:218-219 */
void orc_lf_derivs(const orc_problem *p, const double *x, double *lfx, double *lfxx)
{
    const int n = p->n;
    double dx[NX];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - p->x_target[i];
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) {
            double q = 0.5 * (p->Qf[i * n + j] + p->Qf[j * n + i]);
            lfxx[i * n + j] = q;
            s += q * dx[j];
        }
        lfx[i] = s;
    }
}

/* ------------------------------------------------------------------ passes */

/* iLQR_class.py:79-119 (body), :122-161 (scan, reverse) */
void orc_backward_pass(const orc_problem *p, double phi, const double *X, const double *U, double *U_ff, double *K)
{
    orc_backward_pass_mu(p, phi, 0.0, X, U, U_ff, K);
}

void orc_backward_pass_mu(const orc_problem *p, double phi, double mu, const double *X, const double *U, double *U_ff,
                          double *K)
{
    const int n = p->n, m = p->m, N = p->N;
    double Vx[NX], Vxx[NX * NX], x[NX], u[MX];
    for (int i = 0; i < n; ++i) x[i] = X[i * (N + 1) + N];
    orc_lf_derivs(p, x, Vx, Vxx);                                        /* :136-138 */
    for (int t = N - 1; t >= 0; --t) {
        double lx[NX], lu[MX], lxx[NX * NX], luu[MX * MX], lux[MX * NX], fx[NX * NX], fu[NX * MX];
        double Qx[NX], Qu[MX], Qxx[NX * NX], Qux[MX * NX], Quu[MX * MX], T1[NX * NX], T2[MX * NX];
        for (int i = 0; i < n; ++i) x[i] = X[i * (N + 1) + t];
        for (int j = 0; j < m; ++j) u[j] = U[j * N + t];
        orc_l_derivs(p, x, u, lx, lu, lxx, luu, lux);                    /* :96-97 */
        orc_f_jac(p, t, phi, x, u, fx, fu);
        matmul_tn(n, n, 1, fx, Vx, Qx);                                  /* :100 */
        for (int i = 0; i < n; ++i) Qx[i] = lx[i] + Qx[i];
        matmul_tn(n, m, 1, fu, Vx, Qu);                                  /* :101 */
        for (int j = 0; j < m; ++j) Qu[j] = lu[j] + Qu[j];
        matmul_tn(n, n, n, fx, Vxx, T1);                                 /* (f_x.T @ V_xx) @ f_x, :102 */
        matmul(n, n, n, T1, fx, Qxx);
        for (int i = 0; i < n * n; ++i) Qxx[i] = lxx[i] + Qxx[i];
        matmul_tn(n, m, n, fu, Vxx, T2);                                 /* (f_u.T @ V_xx) @ f_x, :103 */
        matmul(m, n, n, T2, fx, Qux);
        for (int i = 0; i < m * n; ++i) Qux[i] = lux[i] + Qux[i];
        matmul(m, n, m, T2, fu, Quu);                                    /* :104 */
        for (int i = 0; i < m * m; ++i) Quu[i] = luu[i] + Quu[i];
        for (int j = 0; j < m; ++j) Quu[j * m + j] += mu;                /* extension; mu = 0 in the reference */
        /* :109-110  K = -solve(Q_uu, Q_ux), k = -solve(Q_uu, Q_u); LU, no regularisation */
        double lu_m[MX * MX], rhs[MX * (NX + 1)];
        int piv[MX];
        memcpy(lu_m, Quu, sizeof(double) * m * m);
        lu_factor(m, lu_m, piv);
        for (int j = 0; j < m; ++j) {
            for (int i = 0; i < n; ++i) rhs[j * (n + 1) + i] = Qux[j * n + i];
            rhs[j * (n + 1) + n] = Qu[j];
        }
        lu_solve(m, lu_m, piv, rhs, n + 1);
        double Kt[MX * NX], kt[MX];
        for (int j = 0; j < m; ++j) {
            for (int i = 0; i < n; ++i) Kt[j * n + i] = -rhs[j * (n + 1) + i];
            kt[j] = -rhs[j * (n + 1) + n];
        }
        /* :113-114  V_x = Q_x + K.T Q_u ; V_xx = Q_xx + Q_ux.T K  (no symmetrisation) */
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j < m; ++j) s += Kt[j * n + i] * Qu[j];
            Vx[i] = Qx[i] + s;
        }
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < n; ++c) {
                double s = 0.0;
                for (int j = 0; j < m; ++j) s += Qux[j * n + i] * Kt[j * n + c];
                Vxx[i * n + c] = Qxx[i * n + c] + s;
            }
        for (int j = 0; j < m; ++j) {
            U_ff[j * N + t] = kt[j];
            for (int i = 0; i < n; ++i) K[(t * m + j) * n + i] = Kt[j * n + i];
        }
    }
}

/* iLQR_class.py:164-190 (body), :193-247 (scan) */
double orc_forward_pass(const orc_problem *p, double phi, const double *x0, double alpha,
                        const double *X_old, const double *U_old, const double *U_ff,
                        const double *K, double *X_new, double *U_new)
{
    const int n = p->n, m = p->m, N = p->N;
    double x[NX], u[MX], xn[NX], cost = 0.0;
    memcpy(x, x0, sizeof(double) * n);
    for (int t = 0; t < N; ++t) {
        double dx[NX];
        for (int i = 0; i < n; ++i) dx[i] = x[i] - X_old[i * (N + 1) + t];                 /* :181 */
        for (int j = 0; j < m; ++j) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += K[(t * m + j) * n + i] * dx[i];
            u[j] = U_old[j * N + t] + alpha * U_ff[j * N + t] + s;                           /* :182 */
        }
        orc_f(p, t, phi, x, u, xn);                                                          /* :339 */
        cost = cost + orc_l(p, x, u);                                                        /* :187, :340 */
        for (int i = 0; i < n; ++i) X_new[i * (N + 1) + t] = x[i];
        for (int j = 0; j < m; ++j) U_new[j * N + t] = u[j];
        memcpy(x, xn, sizeof(double) * n);
    }
    for (int i = 0; i < n; ++i) X_new[i * (N + 1) + N] = x[i];
    return cost + orc_lf(p, x);                                                              /* :245 */
}

/* iLQR_class.py:250-313 */
double orc_optimize(const orc_problem *p, double phi, const double *x0,
                    double *X, double *U, double *K, double *U_ff,
                    int *iters, int *status, double *cost0,
                    int *trace_alpha_idx, double *trace_cost)
{
    return orc_optimize_ex(p, phi, x0, X, U, K, U_ff, iters, status, cost0, trace_alpha_idx, trace_cost, NULL);
}

/* The regularisation branches (reg != 0) restate the schedule documented in include/ilqr_b200.h; with
 * reg_factor <= 1 this is the reference loop line by line. */
double orc_optimize_ex(const orc_problem *p, double phi, const double *x0,
                       double *X, double *U, double *K, double *U_ff,
                       int *iters, int *status, double *cost0,
                       int *trace_alpha_idx, double *trace_cost, double *mu_out)
{
    const int reg = p->reg_factor > 1.0;
    double mu = reg ? p->reg_init : 0.0;
    int retry = 0;
    const int n = p->n, m = p->m, N = p->N;
    const size_t sx = (size_t)n * (N + 1), su = (size_t)m * N;
    double *Xn = (double *)malloc(sizeof(double) * (sx + su));
    double *Un = Xn + sx;
    /* :257-259 initial rollout, alpha = 0, with the solver's current X, K, U_ff */
    double cost = orc_forward_pass(p, phi, x0, 0.0, X, U, U_ff, K, Xn, Un);
    memcpy(X, Xn, sizeof(double) * sx);
    memcpy(U, Un, sizeof(double) * su);
    if (cost0) *cost0 = cost;
    double cost_prev = cost;
    int it = 0, st = ORC_MAXITER;
    for (int i = 0; i < p->maxiter; ++i) {
        if (i > 0 && !retry && fabs(cost - cost_prev) <= p->tol) { st = ORC_CONVERGED; break; }       /* :267 */
        cost_prev = cost;
        orc_backward_pass_mu(p, phi, mu, X, U, U_ff, K);                                     /* :275 */
        ++it;
        double alpha = 1.0;
        int accepted = -1;
        for (int j = 0; j < p->n_alpha; ++j) {                                               /* :281 */
            double c = orc_forward_pass(p, phi, x0, alpha, X, U, U_ff, K, Xn, Un);
            if (c <= cost) {                                                                 /* :289 */
                memcpy(X, Xn, sizeof(double) * sx);
                memcpy(U, Un, sizeof(double) * su);
                cost = c;
                accepted = j;
                break;
            }
            alpha *= p->alpha_factor;                                                        /* :300 */
            if (alpha < p->min_alpha) break;                                                 /* :301 */
        }
        if (trace_alpha_idx) trace_alpha_idx[i] = accepted;
        if (trace_cost) trace_cost[i] = cost;
        if (accepted < 0) {
            double next = mu * p->reg_factor > p->reg_min ? mu * p->reg_factor : p->reg_min;
            if (!reg || next > p->reg_max) { st = ORC_LS_FAILED; break; }                    /* :304-307 */
            mu = next;                                                                       /* retry the iteration */
            retry = 1;
        } else {
            retry = 0;
            if (reg) { mu = mu / p->reg_factor; if (mu < p->reg_min) mu = 0.0; }
        }
    }
    free(Xn);
    *iters = it;
    *status = st;
    if (mu_out) *mu_out = mu;
    return cost;
}

typedef struct {
    const orc_problem *p; int B; const double *phi, *x0, *U_init;
    double *X, *U, *K, *U_ff, *cost; int *iters, *status;
    atomic_int next;
    int *tr_alpha; double *tr_cost, *cost0;      /* optional per-member traces [B][maxiter] / [B] */
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const orc_problem *p = j->p;
    const int n = p->n, m = p->m, N = p->N;
    const size_t sx = (size_t)n * (N + 1), su = (size_t)m * N, sk = (size_t)N * m * n;
    for (;;) {
        int b = atomic_fetch_add(&j->next, 1);
        if (b >= j->B) break;
        memset(j->X + b * sx, 0, sizeof(double) * sx);          /* iLQR_class.py:55-61 */
        memset(j->K + b * sk, 0, sizeof(double) * sk);
        memset(j->U_ff + b * su, 0, sizeof(double) * su);
        memcpy(j->U + b * su, j->U_init + b * su, sizeof(double) * su);
        const size_t mi = (size_t)(p->maxiter > 0 ? p->maxiter : 1);
        j->cost[b] = orc_optimize(p, j->phi ? j->phi[b] : 0.0, j->x0 + (size_t)b * n, j->X + b * sx,
                                  j->U + b * su, j->K + b * sk, j->U_ff + b * su,
                                  j->iters + b, j->status + b, j->cost0 ? j->cost0 + b : NULL,
                                  j->tr_alpha ? j->tr_alpha + b * mi : NULL, j->tr_cost ? j->tr_cost + b * mi : NULL);
    }
    return NULL;
}

void orc_optimize_batch(const orc_problem *p, int B, const double *phi, const double *x0,
                        const double *U_init, double *X, double *U, double *K, double *U_ff,
                        double *cost, int *iters, int *status, int nthreads)
{
    orc_optimize_batch_trace(p, B, phi, x0, U_init, X, U, K, U_ff, cost, iters, status, nthreads, NULL, NULL, NULL);
}

/* the same with the per-member control flow recorded: accepted try index and cost after every iteration
 * ([B][max(maxiter,1)], untouched beyond iters[b]) and the cost of the alpha = 0 rollout ([B]) */
void orc_optimize_batch_trace(const orc_problem *p, int B, const double *phi, const double *x0,
                              const double *U_init, double *X, double *U, double *K, double *U_ff,
                              double *cost, int *iters, int *status, int nthreads,
                              int *trace_alpha_idx, double *trace_cost, double *cost0)
{
    batch_job j = { p, B, phi, x0, U_init, X, U, K, U_ff, cost, iters, status, 0, trace_alpha_idx, trace_cost, cost0 };
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > B) nthreads = B;
    if (nthreads <= 1) { batch_worker(&j); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, batch_worker, &j);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
}

/* run_iLQR_UA_MPC.py:146-174 */
void orc_mpc(const orc_problem *p_opt, const orc_problem *p_plant, double phi, const double *x0,
             int ticks, double *X_sim, double *U_sim, double *costs, int *iters,
             double *X, double *U, double *K, double *U_ff, double *X_bar_all, double *U_bar_all)
{
    const int n = p_opt->n, m = p_opt->m, N = p_opt->N;
    double cur[NX], uk[MX], xn[NX];
    memcpy(cur, x0, sizeof(double) * n);
    for (int i = 0; i < n; ++i) X_sim[i * (ticks + 1)] = cur[i];
    for (int k = 0; k < ticks; ++k) {
        int st;
        costs[k] = orc_optimize(p_opt, phi, cur, X, U, K, U_ff, iters + k, &st, NULL, NULL, NULL);   /* :148-154 */
        if (X_bar_all) memcpy(X_bar_all + (size_t)k * n * (N + 1), X, sizeof(double) * n * (N + 1));
        if (U_bar_all) memcpy(U_bar_all + (size_t)k * m * N, U, sizeof(double) * m * N);
        for (int j = 0; j < m; ++j) uk[j] = U[j * N];                                         /* :157 */
        orc_f(p_plant, k, phi, cur, uk, xn);                                                  /* :161 */
        for (int j = 0; j < m; ++j) U_sim[j * ticks + k] = uk[j];
        for (int i = 0; i < n; ++i) X_sim[i * (ticks + 1) + k + 1] = xn[i];
        for (int j = 0; j < m; ++j) {                                                         /* :168 shift */
            for (int t = 0; t + 1 < N; ++t) U[j * N + t] = U[j * N + t + 1];
            /* last column repeats U_bar[:, -1] (already in place) */
        }
        memcpy(cur, xn, sizeof(double) * n);
    }
}

int orc_max_threads(void)
{
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}
