/* ilqr_b200.h -- C ABI of libilqr_b200.so: batched iLQR on B200 (sm_100a).
 *
 * The reference (MohamedAbou-Taleb/Iterative-Linear-Quadratic-Regulator) has no FFI or plugin
 * layer: its boundary is the Python class API of python/class_files.  This
 * header is the C boundary a binding for that API sits on; each entry point names the
 * reference method it stands behind (paths relative to the reference's python/ directory).
 * The Python host side in iterative-linear-quadratic-regulator_b200/class_files binds it with
 * ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every array argument is a DEVICE pointer owned by the caller (e.g. a torch tensor's
 *     data_ptr()); the library allocates nothing but its small handle;
 *   - element type is double (dtype ILQR_F64) or float (ILQR_F32) for every array of a handle;
 *   - batch-innermost layouts (B = batch, b fastest):
 *         x0   [n][B]              X  [N+1][n][B]        U  [N][m][B]
 *         K    [N][m][n][B]        k  [N][m][B]  (the reference's U_ff)
 *         A    [N][n][n][B]        Bd [N][n][m][B]       cost [B]
 *         phi  [B]  (ILQR_LTV only; may be NULL otherwise)
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it;
 *   - return value: 0 on success, negative ILQR_E_* on error; nothing throws across the ABI;
 *   - one handle per stream; handles share no state.
 */
#ifndef ILQR_B200_H
#define ILQR_B200_H

#ifdef __CUDACC_RTC__      /* device-side compile of the kernels by NVRTC: no host headers */
typedef int int32_t;
typedef long long int64_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ILQR_NMAX 12
#define ILQR_MMAX 4
#define ILQR_MAX_ALPHAS 16
#define ILQR_MAX_WAVES 8

enum { ILQR_PENDULUM = 0, ILQR_DOUBLE_PENDULUM = 1, ILQR_UA_DOUBLE_PENDULUM = 2, ILQR_LTV = 3,
       /* a user-defined System subclass: its kernels come from a module (ilqr_module_load, ilqr_create_user) */
       ILQR_USER = 4 };
enum { ILQR_EULER = 0, ILQR_MIDPOINT = 1, ILQR_RK4 = 2, ILQR_BACKWARD_EULER = 3 };
enum { ILQR_F64 = 0, ILQR_F32 = 1 };
/* per-trajectory exit status written by ilqr_solve (iLQR_class.py:265-311) */
enum { ILQR_ST_CONVERGED = 0, ILQR_ST_LS_FAILED = 1, ILQR_ST_MAXITER = 2, ILQR_ST_RUNNING = 3 };
enum {
    ILQR_OK = 0,
    ILQR_E_INVALID = -1,      /* bad argument / unsupported combination */
    ILQR_E_CUDA = -2,         /* a CUDA runtime call failed; see ilqr_last_cuda_error() */
    ILQR_E_WORKSPACE = -3     /* workspace smaller than ilqr_workspace_bytes() */
};

/* Problem definition = System ctor arguments + iLQR ctor arguments
 * (class_files/systems/system_base.py:25-30, pendulum_sys.py:22-32, double_pendulum_sys.py:20-38,
 *  UA_double_pendulum_sys.py:20-38, iLQR_class.py:18-27). */
typedef struct ilqr_problem_t {
    int32_t model;        /* ILQR_PENDULUM ... */
    int32_t integrator;   /* ILQR_EULER ... */
    int32_t dtype;        /* ILQR_F64 / ILQR_F32 */
    int32_t n, m;         /* must match the model: (2,1), (4,2), (4,1), (12,4) for ILQR_LTV; n <= ILQR_NMAX, m <= ILQR_MMAX */
    int32_t N;            /* horizon: len(arange(0, T+dt, dt)) - 1 (iLQR_class.py:46-47) */
    int32_t B;            /* batch of independent trajectories on this device */
    int32_t n_alpha;      /* line-search tries, 10 in the reference (iLQR_class.py:281) */
    int32_t maxiter;      /* iLQR_class.py:24 */
    int32_t reserved;
    double dt, tol, alpha_factor, min_alpha;
    double phys[16];      /* pendulum: g,l,d ; double pendulums: g,m1,m2,l1,l2,d1,d2,theta1,theta2 */
    double Q[ILQR_NMAX * ILQR_NMAX];    /* row-major n x n */
    double R[ILQR_MMAX * ILQR_MMAX];    /* row-major m x m */
    double Qf[ILQR_NMAX * ILQR_NMAX];
    double x_target[ILQR_NMAX];
    /* ILQR_LTV: x+ = x + dt ((Ac + ltv_amp sin(2 pi t/N + phi_b) E) x + Bc u) */
    double Ac[ILQR_NMAX * ILQR_NMAX], E[ILQR_NMAX * ILQR_NMAX], Bc[ILQR_NMAX * ILQR_MMAX];
    double ltv_amp;
    /* EXTENSION (the reference has no regularisation, iLQR_class.py:109-110): per-trajectory Levenberg-
     * Marquardt term, Q_uu + mu I, scheduled on the device.  reg_factor <= 1 (default 0) disables it and
     * keeps the reference behaviour exactly.  Enabled: mu starts at reg_init; an iteration whose line search
     * accepts no step size is retried with mu <- max(mu reg_factor, reg_min) instead of ending the solve
     * (:304-307), which fails only once mu > reg_max; after an accepted step mu <- mu / reg_factor (0 below
     * reg_min). */
    double reg_init, reg_factor, reg_min, reg_max;
} ilqr_problem_t;

typedef struct ilqr_handle_s *ilqr_handle_t;

#ifndef __CUDACC_RTC__     /* the entry points are host functions: not part of a device-side (NVRTC) compile */

/* iLQR.__init__ / System.__init__: validates the problem (unknown integrator or a model/dimension
 * mismatch -> ILQR_E_INVALID, the counterpart of the ValueErrors at iLQR_class.py:50-52 and
 * system_base.py:197-198) and precomputes the device-side constants. */
int ilqr_create(const ilqr_problem_t *problem, ilqr_handle_t *out);
int ilqr_destroy(ilqr_handle_t h);

/* ---- user-defined System subclasses (model ILQR_USER) -------------------------------------------------------------
 * The reference derives and jit-compiles everything a new System needs at construction, in process
 * (system_base.py:203-251).  Here the host side (class_files/codegen.py) traces the subclass's three methods, generates
 * the device model (ilqr::UserSys<T>, ilqr::UserCost<T>), compiles the library's generic kernel templates against it
 * with NVRTC for sm_100a -- in process, no nvcc, no host compiler -- and hands the cubin to the library:
 *   image        cubin (NVRTC output) holding the instantiations of the ILQR_UK_* kernels for ONE (integrator, dtype)
 *   kernel_names their lowered (mangled) names, indexed by ILQR_UK_*
 *   n, m         state / control dimension of the generated model (n <= ILQR_NMAX, m <= ILQR_MMAX)
 * A module serves any number of handles (ilqr_create_user) on the device that was current at load time and must
 * outlive them. */
typedef struct ilqr_module_s *ilqr_module_t;
enum { ILQR_UK_STEP = 0,            /* step_kernel<UserSys, INTEG, T> */
       ILQR_UK_LINEARIZE = 1,       /* commit_linearize_kernel<UserSys, INTEG, T> */
       ILQR_UK_COST_EXPANSION = 2,  /* cost_expansion_kernel<UserCost, T, n, m> */
       ILQR_UK_BACKWARD_SMALL = 3,  /* backward_kernel<UserCost, T, n, m, 8, 32>   (n > 4: <.., 2, 32>) */
       ILQR_UK_BACKWARD_LARGE = 4,  /* backward_kernel<UserCost, T, n, m, 4, 64>   (n > 4: <.., 2, 32>) */
       ILQR_UK_ROLLOUT = 5,         /* rollout_kernel<UserSys, UserCost, INTEG, T> */
       ILQR_N_USER_KERNELS = 6 };
int ilqr_module_load(const void *image, size_t bytes, const char *const *kernel_names, int n, int m, int integrator,
                     int dtype, ilqr_module_t *out);
int ilqr_module_unload(ilqr_module_t mod);
/* ilqr_create for problem->model == ILQR_USER: n, m, integrator and dtype must be the module's */
int ilqr_create_user(const ilqr_problem_t *problem, ilqr_module_t mod, ilqr_handle_t *out);

/* bytes of caller-provided device scratch needed by ilqr_backward_pass / ilqr_solve */
size_t ilqr_workspace_bytes(ilqr_handle_t h);

/* System.f_fcn for a batch (system_base.py:223 f_fcn): xn[n][B] = f(x[n][B], u[m][B]).
 * Used for the MPC plant step (run_iLQR_UA_MPC.py:161).  t is the step index (ILQR_LTV only). */
int ilqr_step(ilqr_handle_t h, int t, const void *phi, const void *x, const void *u, void *xn, void *stream);

/* System.f_x_fcn / f_u_fcn for every timestep of every trajectory at once
 * (iLQR_class.py:318-331 -> system_base.py:203-219): A[N][n][n][B], Bd[N][n][m][B]. */
int ilqr_linearize(ilqr_handle_t h, const void *phi, const void *X, const void *U, void *A, void *Bd,
                   void *stream);

/* System.l_fcn / l_x_fcn / l_u_fcn / l_xx_fcn / l_uu_fcn / l_ux_fcn for every timestep of every trajectory
 * and l_f_fcn / l_f_x_fcn / l_f_xx_fcn at X[N] (system_base.py:212-219,235-245):
 *   l[N][B] lx[N][n][B] lu[N][m][B] lxx[N][n][n][B] luu[N][m][m][B] lux[N][m][n][B] lf[B] lfx[n][B] lfxx[n][n][B].
 * Any output pointer may be NULL.  (ilqr_backward fuses l_x,l_u into the scan; this entry point is the
 * materialised form of the same expansion.) */
int ilqr_cost_expansion(ilqr_handle_t h, const void *X, const void *U, void *l, void *lx, void *lu, void *lxx,
                        void *luu, void *lux, void *lf, void *lfx, void *lfxx, void *stream);

/* The reverse scan of iLQR._backward_pass_scan (iLQR_class.py:79-161) on a given linearization. */
int ilqr_backward(ilqr_handle_t h, const void *X, const void *U, const void *A, const void *Bd, void *K,
                  void *k, void *stream);

/* iLQR.backward_pass(X_nom, U_nom) -> (U_ff, K)  (iLQR_class.py:122-161): linearize + reverse scan. */
int ilqr_backward_pass(ilqr_handle_t h, const void *phi, const void *X, const void *U, void *K, void *k,
                       void *workspace, size_t workspace_bytes, void *stream);

/* iLQR.forward_pass(x_0, alpha, X_old, U_old, U_ff, K) -> (X_new, U_new, cost)
 * (iLQR_class.py:164-247), one alpha for the whole batch. */
int ilqr_rollout(ilqr_handle_t h, const void *phi, const void *x0, double alpha, const void *X_old,
                 const void *U_old, const void *k, const void *K, void *X_new, void *U_new, void *cost,
                 void *stream);

/* Line search of iLQR.optimize_trajectory (iLQR_class.py:278-307) with all n_alpha step sizes
 * rolled out concurrently: Xc[n_alpha][N+1][n][B], Uc[n_alpha][N][m][B], cost_alpha[n_alpha][B];
 * winner[B] (int32) = lowest try index with cost_alpha <= cost, or -1. */
int ilqr_forward_linesearch(ilqr_handle_t h, const void *phi, const void *x0, const void *X, const void *U,
                            const void *k, const void *K, const void *cost, void *Xc, void *Uc,
                            void *cost_alpha, int32_t *winner, void *stream);

/* iLQR.optimize_trajectory() (iLQR_class.py:250-313) for the whole batch with no per-iteration
 * host round trip.  X,U,K,k are the solver's persistent attributes (in/out: the alpha=0 initial
 * rollout uses the incoming X,K,k exactly like :257-259, which is what MPC warm starts rely on).
 * Outputs: cost[B], iters[B] (int32, backward passes executed), status[B] (int32, ILQR_ST_*).
 * total_iters (host pointer, may be NULL) receives sum_b iters[b] after the stream is synchronized
 * by the call.  With NULL the call does not wait for the solve to finish: the device never waits for the host, but
 * the HOST may block inside the call until all but the last block of 8 iterations has run (it polls the
 * device-side active counter one block behind to stop enqueuing once every trajectory has finished), so the call
 * cannot be captured into a CUDA graph. */
int ilqr_solve(ilqr_handle_t h, const void *phi, const void *x0, void *X, void *U, void *K, void *k,
               void *cost, int32_t *iters, int32_t *status, void *workspace, size_t workspace_bytes,
               void *stream, int64_t *total_iters);

/* Schedule of the line search inside ilqr_solve.  The reference tries its step sizes one after the other
 * and stops at the first acceptable one (iLQR_class.py:279-302).  On the GPU the tries of a WAVE are rolled
 * out concurrently; trajectories that accepted none of them go on a compacted list and only those are
 * rolled out in the next wave.  sizes[n_waves] are the tries per wave (their sum is clamped to the number of
 * tries actually made).  n_waves = 0 selects the eager schedule (all tries in one or two dense waves), which
 * is the default for small batches where the rollout is latency bound; large batches default to waves of
 * 2,2,2,rest.  The accepted step size of every trajectory is the same under every schedule. */
int ilqr_set_linesearch_waves(ilqr_handle_t h, int n_waves, const int32_t *sizes);
/* The schedule in force: returns the number of lazy waves (0 = eager) and, when sizes != NULL, writes the tries per
 * wave into sizes[ILQR_MAX_WAVES] (eager: sizes[0] = tries of the first, dense wave). */
int ilqr_get_linesearch_waves(ilqr_handle_t h, int32_t *sizes);

/* Optional device buffer mu[B] (element type of the handle) for the regularisation state of ilqr_solve:
 * initialised to reg_init at the start of every solve and left holding the final values.  NULL (default)
 * keeps the state in the workspace.  Ignored while reg_factor <= 1. */
int ilqr_set_mu_buffer(ilqr_handle_t h, void *mu);

/* Optional per-iteration trace written by ilqr_solve (the information the reference prints when
 * verbose=True, iLQR_class.py:262,296,306): alpha_idx[maxiter][B] (int32: accepted try index, -1 = line
 * search failed, untouched where the trajectory did not run) and cost_trace[maxiter+1][B] (row 0 = cost of
 * the initial rollout, row it+1 = cost after iteration it).  NULL disables. */
int ilqr_set_trace(ilqr_handle_t h, int32_t *alpha_idx, void *cost_trace);

/* MPC warm-start shift, run_iLQR_UA_MPC.py:168: U[t] <- U[t+1], last column repeated; also
 * returns u0[m][B] = U[0] before the shift (run_iLQR_UA_MPC.py:157). */
int ilqr_mpc_shift(ilqr_handle_t h, void *U, void *u0, void *stream);

/* Optional per-kernel timing of ilqr_solve: CUDA events are chained between its launches on the
 * caller's stream and folded into per-class totals when the solve synchronizes (total_iters != NULL).
 * Classes index the arrays of ilqr_get_kernel_times (ms[ILQR_N_KERNEL_CLASSES], launches[...]).
 * ILQR_KC_OTHER collects the select/convergence kernel and the gaps between launches. */
enum { ILQR_KC_LINEARIZE = 0, ILQR_KC_BACKWARD = 1, ILQR_KC_ROLLOUT = 2, ILQR_KC_INIT_ROLLOUT = 3,
       ILQR_KC_OTHER = 4, ILQR_N_KERNEL_CLASSES = 5 };
int ilqr_set_profiling(ilqr_handle_t h, int enable);
int ilqr_get_kernel_times(ilqr_handle_t h, double *ms, int64_t *launches);

/* Measured FP64 FMA throughput of the current device in TFLOP/s (best of three ~10 ms runs of independent DFMA
 * chains on every SM): the denominator of bench.py's FP64 roofline for the rollout kernel, which is bound by the FP64
 * pipe, not by HBM.  scratch = any device buffer of >= 8 bytes.  Synchronizes the stream. */
int ilqr_fp64_peak(double *tflops, void *scratch, void *stream);

/* number of kernel launches issued through this handle since creation */
int64_t ilqr_launch_count(ilqr_handle_t h);
/* last cudaError_t seen by this handle (0 = none) and its string */
int ilqr_last_cuda_error(ilqr_handle_t h);
const char *ilqr_strerror(int code);
const char *ilqr_version(void);
#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* ILQR_B200_H */
