/* ilqr_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product).
 *
 * CPU float64 restatement of the reference's iLQR hot path, one trajectory at a time,
 * in the reference's own array conventions ((dim,time) for X/U/U_ff, (time,m,n) for K).
 * Follows /root/reference/python/class_files/iLQR_class.py and systems/<system>.py; every
 * function in ilqr_oracle.c cites the reference lines it restates.
 *
 * Parity status: PINNED against outputs of the unmodified reference sources executed in
 * the build container over oracle/jaxshim (tests/golden/<case>.npz, tests/golden/make_golden.py).
 * The reference itself ships no tests or golden vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#ifndef ILQR_ORACLE_H
#define ILQR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NMAX 12
#define ORC_MMAX 4

enum { ORC_PENDULUM = 0, ORC_DOUBLE_PENDULUM = 1, ORC_UA_DOUBLE_PENDULUM = 2, ORC_LTV = 3 };
enum { ORC_EULER = 0, ORC_MIDPOINT = 1, ORC_RK4 = 2, ORC_BACKWARD_EULER = 3 };
enum { ORC_CONVERGED = 0, ORC_LS_FAILED = 1, ORC_MAXITER = 2 };

typedef struct {
    int model, integrator, n, m, N;
    int n_alpha;        /* line-search tries; 10 in the reference (iLQR_class.py:281) */
    int maxiter;        /* iLQR_class.py:24 */
    double dt, tol, alpha_factor, min_alpha;
    /* pendulum: g,l,d   double pendulum: g,m1,m2,l1,l2,d1,d2,theta1,theta2 */
    double phys[16];
    double Q[ORC_NMAX * ORC_NMAX];   /* row-major n x n */
    double R[ORC_MMAX * ORC_MMAX];   /* row-major m x m */
    double Qf[ORC_NMAX * ORC_NMAX];
    double x_target[ORC_NMAX];
    /* synthetic LTV (BASELINE config 4): x+ = x + dt*((Ac + amp*sin(2*pi*t/N + phi)*E) x + Bc u) */
    double Ac[ORC_NMAX * ORC_NMAX], E[ORC_NMAX * ORC_NMAX], Bc[ORC_NMAX * ORC_MMAX];
    double ltv_amp;
    /* EXTENSION restated from include/ilqr_b200.h (the reference has no regularisation): Levenberg-Marquardt
     * schedule of Q_uu + mu I; reg_factor <= 1 disables it. */
    double reg_init, reg_factor, reg_min, reg_max;
} orc_problem;

/* point functions; t and phi only matter for ORC_LTV */
void orc_f_cont(const orc_problem *p, int t, double phi, const double *x, const double *u, double *xdot);
void orc_f_cont_jac(const orc_problem *p, int t, double phi, const double *x, const double *u,
                    double *Ac /* n*n */, double *Bc /* n*m */);
int  orc_f(const orc_problem *p, int t, double phi, const double *x, const double *u, double *xn);
void orc_f_jac(const orc_problem *p, int t, double phi, const double *x, const double *u,
               double *A /* n*n */, double *B /* n*m */);
double orc_l(const orc_problem *p, const double *x, const double *u);
double orc_lf(const orc_problem *p, const double *x);
void orc_l_derivs(const orc_problem *p, const double *x, const double *u,
                  double *lx, double *lu, double *lxx, double *luu, double *lux);
void orc_lf_derivs(const orc_problem *p, const double *x, double *lfx, double *lfxx);

/* X (n,N+1), U (m,N), U_ff (m,N) row-major (dim,time); K (N,m,n) */
void orc_backward_pass(const orc_problem *p, double phi, const double *X, const double *U,
                       double *U_ff, double *K);
/* same with Q_uu + mu I (extension; mu = 0 is the reference) */
void orc_backward_pass_mu(const orc_problem *p, double phi, double mu, const double *X, const double *U,
                          double *U_ff, double *K);
double orc_forward_pass(const orc_problem *p, double phi, const double *x0, double alpha,
                        const double *X_old, const double *U_old, const double *U_ff,
                        const double *K, double *X_new, double *U_new);

/* optimize_trajectory(): X,U,K,U_ff are the solver's persistent attributes (in/out).
 * trace_alpha_idx[it] = accepted try index (or -1), trace_cost[it] = cost after iteration it.
 * Returns the final cost; *iters = backward passes executed; *status = ORC_*. */
double orc_optimize(const orc_problem *p, double phi, const double *x0,
                    double *X, double *U, double *K, double *U_ff,
                    int *iters, int *status, double *cost0,
                    int *trace_alpha_idx, double *trace_cost);

/* orc_optimize with the regularisation state returned: *mu_out = mu after the last iteration */
double orc_optimize_ex(const orc_problem *p, double phi, const double *x0,
                       double *X, double *U, double *K, double *U_ff,
                       int *iters, int *status, double *cost0,
                       int *trace_alpha_idx, double *trace_cost, double *mu_out);

/* B independent solves in batch-major layout (X[b] is (n,N+1) etc.), fresh solver state
 * (X=K=U_ff=0) per trajectory, fanned out over nthreads POSIX threads. */
void orc_optimize_batch(const orc_problem *p, int B, const double *phi, const double *x0,
                        const double *U_init, double *X, double *U, double *K, double *U_ff,
                        double *cost, int *iters, int *status, int nthreads);

/* The same with the per-member control flow recorded (accepted try index and cost after every iteration,
 * [B][max(maxiter,1)]; cost of the alpha = 0 rollout, [B]); any of the three may be NULL. */
void orc_optimize_batch_trace(const orc_problem *p, int B, const double *phi, const double *x0,
                              const double *U_init, double *X, double *U, double *K, double *U_ff,
                              double *cost, int *iters, int *status, int nthreads,
                              int *trace_alpha_idx, double *trace_cost, double *cost0);

/* Receding-horizon loop, run_iLQR_UA_MPC.py:146-174: one solver object re-used across ticks. */
void orc_mpc(const orc_problem *p_opt, const orc_problem *p_plant, double phi, const double *x0,
             int ticks, double *X_sim /* (n,ticks+1) */, double *U_sim /* (m,ticks) */,
             double *costs, int *iters, double *X, double *U, double *K, double *U_ff,
             double *X_bar_all /* (ticks,n,N+1) or NULL */, double *U_bar_all /* (ticks,m,N) or NULL */);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
