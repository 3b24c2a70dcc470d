// ilqr_b200.cu -- batched iLQR kernels for B200 (sm_100a) and the C ABI of include/ilqr_b200.h.
//
// Four kernels per iLQR iteration (BASELINE.json north_star):
//   K1 commit_linearize_kernel  commit the accepted line-search candidate, then analytic A_t,B_t for
//                               every (t,b) at once                 (iLQR_class.py:318-331)
//   K2 backward_kernel          reverse Riccati scan, one trajectory per thread, value function
//                               in registers, coalesced K/k stores  (iLQR_class.py:79-161)
//   K3 rollout_kernel           forward rollout of every (alpha,b) pair concurrently
//                                                                    (iLQR_class.py:164-247,278-302)
//   K4 select_kernel            first-acceptable-alpha selection, convergence test, per-trajectory
//                               status/iteration bookkeeping on device (iLQR_class.py:265-271,289-307)
// All arrays are batch-innermost so that a warp's 32 trajectories touch 32 consecutive elements.
// Paths in comments are relative to /root/reference/python/class_files/.
#include "ilqr_b200.h"
#include "ilqr_systems.cuh"
// A library for ONE user-defined System subclass is this same file compiled with
//   -DILQR_USER_SYS -DILQR_USER_HEADER="<generated>.cuh" -DILQR_USER_INTEG=<0..3> -DILQR_USER_F32=<0|1>
// (class_files/codegen.py): the generated header defines ilqr::UserSys<T> / ilqr::UserCost<T>, and only that
// model, integrator and element type are instantiated.
#ifdef ILQR_USER_SYS
#include ILQR_USER_HEADER
#endif

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <type_traits>
#include <vector>

namespace ilqr {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

struct AlphaList { double a[ILQR_MAX_ALPHAS]; };

// iteration-control block living at the head of the workspace
struct Control {
    unsigned long long total_iters;          // sum over trajectories of backward passes executed
    unsigned int n_active[1];                 // [maxiter + 2], n_active[it] = trajectories entering iteration it
};

// Speculative evaluation of the deferred (second-wave) step sizes.  Trajectories that needed a small
// step in the previous iteration are put on a list by select_kernel; the first-wave rollout launch
// carries `cap * n2` extra threads (the warp slots left over when the wave is sized to the SM
// sub-partitions) that roll out the deferred step sizes for the listed trajectories, so that the
// separate, latency-bound second wave is almost never needed.  Which rollouts are evaluated never
// changes which one is accepted.
struct SpecArgs {
    int cap;                              // list capacity; 0 switches speculation off
    int n2;                               // deferred step sizes per trajectory
    int threshold;                        // accepted try index from which a trajectory is listed
    int *list_cur, *list_next;            // [cap]
    unsigned int *count_cur, *count_next; // entries appended this / next iteration (may exceed cap)
    int *mark;                            // [B]: mark[b] == it + 1 <=> b is on the list of iteration it
};

// Levenberg-Marquardt regularisation of Q_uu, kept per trajectory on the device (an EXTENSION: the
// reference has none, iLQR_class.py:109-110, and with factor <= 1 nothing here changes its behaviour).
// The backward pass solves with Q_uu + mu I.  When the line search of an iteration accepts no step size,
// the reference stops the solve (:304-307); with the schedule enabled the trajectory instead retries the
// iteration with mu <- max(mu * factor, mu_min), and fails only once mu exceeds mu_max.  After an accepted
// step mu <- mu / factor (snapped to 0 below mu_min).  All of it runs in the select kernels.
struct RegArgs {
    void *mu;                 // [B], T; nullptr <=> schedule disabled
    double factor, mu_min, mu_max;
};

template <typename T>
ILQR_DEV bool reg_on_failure(const RegArgs &rg, int b)
{
    // true: retry with a larger mu; false: give up (reference behaviour)
    if (!rg.mu) return false;
    T *mu = (T *)rg.mu;
    const T next = mu[b] * (T)rg.factor > (T)rg.mu_min ? mu[b] * (T)rg.factor : (T)rg.mu_min;
    if (next > (T)rg.mu_max) return false;
    mu[b] = next;
    return true;
}

template <typename T>
ILQR_DEV void reg_on_success(const RegArgs &rg, int b)
{
    if (!rg.mu) return;
    T *mu = (T *)rg.mu;
    const T next = mu[b] / (T)rg.factor;
    mu[b] = next < (T)rg.mu_min ? T(0) : next;
}

template <class Sys, int INTEG, typename T>
__global__ void step_kernel(const __grid_constant__ Sys sys, T dt, int B, int t, const T *__restrict__ phi,
                            const T *__restrict__ x, const T *__restrict__ u, T *__restrict__ xn)
{
    constexpr int n = Sys::N, m = Sys::M;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T xv[n], uv[m], out[n];
#pragma unroll
    for (int i = 0; i < n; ++i) xv[i] = x[(size_t)i * B + b];
#pragma unroll
    for (int j = 0; j < m; ++j) uv[j] = u[(size_t)j * B + b];
    step<INTEG>(sys, dt, xv, uv, out, sys.time_scalar(t, phi ? phi[b] : T(0)));
#pragma unroll
    for (int i = 0; i < n; ++i) xn[(size_t)i * B + b] = out[i];
}

// K1.  One thread per (t,b), t in [0,N].  If winner != nullptr and winner[b] >= 0 the thread first
// copies the accepted candidate (Xc/Uc slab winner[b]) into the nominal X/U; if the trajectory is
// active it then writes the discrete Jacobians about that nominal point.
template <class Sys, int INTEG, typename T>
__global__ void commit_linearize_kernel(const __grid_constant__ Sys sys, T dt, int N, int B,
                                        const T *__restrict__ phi, T *__restrict__ X, T *__restrict__ U,
                                        T *__restrict__ A, T *__restrict__ Bd, const T *__restrict__ Xc,
                                        const T *__restrict__ Uc, const int *__restrict__ winner,
                                        const int *__restrict__ wslot, const int *__restrict__ active, int do_linearize,
                                        const unsigned int *__restrict__ gate0, const unsigned int *__restrict__ gate1)
{
    constexpr int n = Sys::N, m = Sys::M;
    if (gate0 && *gate0 == 0u && *gate1 == 0u) return;   // nobody active now or in the previous iteration
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)(N + 1) * B) return;
    const int t = (int)(gid / B), b = (int)(gid % B);
    const int w = winner ? winner[b] : -1;
    const bool act = do_linearize && (active ? active[b] != 0 : true) && t < N;
    if (w < 0 && !act) return;
    T x[n], u[m];
    if (w >= 0) {
        // lazy line search: candidates of the later waves are stored at the trajectory's list position
        const int col = wslot ? wslot[b] : b;
        const T *xs = Xc + (size_t)w * (N + 1) * n * B, *us = Uc + (size_t)w * N * m * B;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            x[i] = xs[((size_t)t * n + i) * B + col];
            X[((size_t)t * n + i) * B + b] = x[i];
        }
        if (t < N) {
#pragma unroll
            for (int j = 0; j < m; ++j) {
                u[j] = us[((size_t)t * m + j) * B + col];
                U[((size_t)t * m + j) * B + b] = u[j];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < n; ++i) x[i] = X[((size_t)t * n + i) * B + b];
#pragma unroll
        for (int j = 0; j < m; ++j) u[j] = U[((size_t)t * m + j) * B + b];
    }
    if (!act) return;
    T Aj[n][n], Bj[n][m];
    step_jac<INTEG>(sys, dt, x, u, Aj, Bj, sys.time_scalar(t, phi ? phi[b] : T(0)));
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < n; ++j) A[(((size_t)t * n + i) * n + j) * B + b] = Aj[i][j];
#pragma unroll
        for (int j = 0; j < m; ++j) Bd[(((size_t)t * n + i) * m + j) * B + b] = Bj[i][j];
    }
}

// K2.  One thread per trajectory; V_x, V_xx live in registers for the whole scan.  The scan is
// sequential in t, so at small batches a warp cannot hide HBM latency by occupancy: each thread
// streams its own A_t, B_t, x_t, u_t through a DEPTH-deep shared-memory ring with cp.async
// (LDGSTS), DEPTH-1 timesteps ahead of the arithmetic.  A thread only ever reads the ring slots it
// filled itself, so cp.async.wait_group is the only synchronisation needed (no block barrier).
template <typename T, int n, int m>
struct BwdIn { T A[n][n], Bd[n][m], x[n], u[m]; };

template <int BYTES>
ILQR_DEV void cp_async(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
ILQR_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> ILQR_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// rows of one ring stage: A (n*n), Bd (n*m), x (n), u (m); element [row][tid]
template <typename T, int n, int m>
ILQR_DEV void bwd_issue(T *stage, int t, int b, int B, const T *__restrict__ X, const T *__restrict__ U,
                        const T *__restrict__ A, const T *__restrict__ Bd)
{
    const int bd = blockDim.x, tid = threadIdx.x;
    int row = 0;
#pragma unroll
    for (int i = 0; i < n * n; ++i, ++row) cp_async<sizeof(T)>(stage + row * bd + tid, A + ((size_t)t * n * n + i) * B + b);
#pragma unroll
    for (int i = 0; i < n * m; ++i, ++row) cp_async<sizeof(T)>(stage + row * bd + tid, Bd + ((size_t)t * n * m + i) * B + b);
#pragma unroll
    for (int i = 0; i < n; ++i, ++row) cp_async<sizeof(T)>(stage + row * bd + tid, X + ((size_t)t * n + i) * B + b);
#pragma unroll
    for (int i = 0; i < m; ++i, ++row) cp_async<sizeof(T)>(stage + row * bd + tid, U + ((size_t)t * m + i) * B + b);
}

template <typename T, int n, int m>
ILQR_DEV void bwd_read(BwdIn<T, n, m> &d, const T *stage)
{
    const int bd = blockDim.x, tid = threadIdx.x;
    int row = 0;
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j, ++row) d.A[i][j] = stage[row * bd + tid];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < m; ++j, ++row) d.Bd[i][j] = stage[row * bd + tid];
#pragma unroll
    for (int i = 0; i < n; ++i, ++row) d.x[i] = stage[row * bd + tid];
#pragma unroll
    for (int j = 0; j < m; ++j, ++row) d.u[j] = stage[row * bd + tid];
}

template <class Cost, typename T, int n, int m, int DEPTH>
__global__ void backward_kernel(const __grid_constant__ Cost qc, int N, int B, const T *__restrict__ X,
                                const T *__restrict__ U, const T *__restrict__ A, const T *__restrict__ Bd,
                                T *__restrict__ K, T *__restrict__ k, const int *__restrict__ active,
                                const unsigned int *__restrict__ gate, const T *__restrict__ mu)
{
    constexpr int L = n * n + n * m + n + m;
    extern __shared__ __align__(16) unsigned char ring_raw[];
    T *ring = reinterpret_cast<T *>(ring_raw);
    if (gate && *gate == 0u) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (active && !active[b]) return;
    const int stage_elems = L * blockDim.x;
#pragma unroll
    for (int s = 0; s < DEPTH; ++s) {
        if (N - 1 - s >= 0) bwd_issue<T, n, m>(ring + s * stage_elems, N - 1 - s, b, B, X, U, A, Bd);
        cp_async_commit();
    }
    const T mu_b = mu ? mu[b] : T(0);                                    // regularisation (RegArgs), 0 in the reference
    T Vx[n], Vxx[n][n];
    {
        T xN[n];
#pragma unroll
        for (int i = 0; i < n; ++i) xN[i] = X[((size_t)N * n + i) * B + b];
        if constexpr (Cost::QUADRATIC) {
            qc.terminal_grad(xN, Vx);                                    // iLQR_class.py:136-138
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j < n; ++j) Vxx[i][j] = qc.Qfs[i][j];
        } else {
            qc.terminal_expand(xN, Vx, Vxx);
        }
    }
    BwdIn<T, n, m> cur;
    int stage = 0;
    for (int t = N - 1; t >= 0; --t) {
        cp_async_wait<DEPTH - 1>();                                       // the group holding step t has landed
        bwd_read(cur, ring + stage * stage_elems);
        T lx[n], lu[m];
        // quadratic costs: l_xx = Q dt, l_uu = R dt, l_ux = 0 are constants folded into the sums below;
        // generated user costs (ilqr_user.cuh) provide the full state-dependent expansion
        [[maybe_unused]] T lxx[Cost::QUADRATIC ? 1 : n][Cost::QUADRATIC ? 1 : n];
        [[maybe_unused]] T luu[Cost::QUADRATIC ? 1 : m][Cost::QUADRATIC ? 1 : m];
        [[maybe_unused]] T lux[Cost::QUADRATIC ? 1 : m][Cost::QUADRATIC ? 1 : n];
        if constexpr (Cost::QUADRATIC) qc.grad(cur.x, cur.u, lx, lu);
        else qc.expand(cur.x, cur.u, lx, lu, lxx, luu, lux);
        // Q_x = l_x + f_x' V_x ; Q_u = l_u + f_u' V_x                    (:100-101)
        T Qx[n], Qu[m];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += cur.A[l][i] * Vx[l];
            Qx[i] = lx[i] + s;
        }
#pragma unroll
        for (int j = 0; j < m; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += cur.Bd[l][j] * Vx[l];
            Qu[j] = lu[j] + s;
        }
        // T1 = f_x' V_xx, T2 = f_u' V_xx ; Q_xx = l_xx + T1 f_x ; Q_ux = T2 f_x ; Q_uu = l_uu + T2 f_u   (:102-104)
        T T1[n][n], T2[m][n], Qxx[n][n], Qux[m][n], Quu[m][m];
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                T s = T(0);
#pragma unroll
                for (int l = 0; l < n; ++l) s += cur.A[l][i] * Vxx[l][j];
                T1[i][j] = s;
            }
#pragma unroll
        for (int i = 0; i < m; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                T s = T(0);
#pragma unroll
                for (int l = 0; l < n; ++l) s += cur.Bd[l][i] * Vxx[l][j];
                T2[i][j] = s;
            }
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                T s = T(0);
#pragma unroll
                for (int l = 0; l < n; ++l) s += T1[i][l] * cur.A[l][j];
                if constexpr (Cost::QUADRATIC) Qxx[i][j] = qc.Qs[i][j] * qc.dt + s;
                else Qxx[i][j] = lxx[i][j] + s;
            }
#pragma unroll
        for (int i = 0; i < m; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                T s = T(0);
#pragma unroll
                for (int l = 0; l < n; ++l) s += T2[i][l] * cur.A[l][j];
                if constexpr (Cost::QUADRATIC) Qux[i][j] = s;            // l_ux = 0 for the quadratic cost
                else Qux[i][j] = lux[i][j] + s;
            }
#pragma unroll
            for (int j = 0; j < m; ++j) {
                T s = T(0);
#pragma unroll
                for (int l = 0; l < n; ++l) s += T2[i][l] * cur.Bd[l][j];
                if constexpr (Cost::QUADRATIC) Quu[i][j] = qc.Rs[i][j] * qc.dt + s;
                else Quu[i][j] = luu[i][j] + s;
                if (i == j) Quu[i][j] += mu_b;
            }
        }
        // K = -Q_uu^-1 Q_ux, k = -Q_uu^-1 Q_u                            (:109-110; no regularisation)
        T Kt[m][n], kt[m];
        if (m == 1) {
            const T r = -rcp_t(Quu[0][0]);
#pragma unroll
            for (int j = 0; j < n; ++j) Kt[0][j] = Qux[0][j] * r;
            kt[0] = Qu[0] * r;
        } else {
            T rhs[m][n + 1];
#pragma unroll
            for (int i = 0; i < m; ++i) {
#pragma unroll
                for (int j = 0; j < n; ++j) rhs[i][j] = Qux[i][j];
                rhs[i][n] = Qu[i];
            }
            T Lm[m][m];
#pragma unroll
            for (int i = 0; i < m; ++i)
#pragma unroll
                for (int j = 0; j < m; ++j) Lm[i][j] = Quu[i][j];
            lu_solve_inplace<m, n + 1>(Lm, rhs);
#pragma unroll
            for (int i = 0; i < m; ++i) {
#pragma unroll
                for (int j = 0; j < n; ++j) Kt[i][j] = -rhs[i][j];
                kt[i] = -rhs[i][n];
            }
        }
        // V_x = Q_x + K' Q_u ; V_xx = Q_xx + Q_ux' K                      (:113-114; not symmetrised)
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < m; ++j) s += Kt[j][i] * Qu[j];
            Vx[i] = Qx[i] + s;
#pragma unroll
            for (int c = 0; c < n; ++c) {
                T s2 = T(0);
#pragma unroll
                for (int j = 0; j < m; ++j) s2 += Qux[j][i] * Kt[j][c];
                Vxx[i][c] = Qxx[i][c] + s2;
            }
        }
#pragma unroll
        for (int j = 0; j < m; ++j) {
#pragma unroll
            for (int i = 0; i < n; ++i) K[(((size_t)t * m + j) * n + i) * B + b] = Kt[j][i];
            k[((size_t)t * m + j) * B + b] = kt[j];
        }
        if (t - DEPTH >= 0) bwd_issue<T, n, m>(ring + stage * stage_elems, t - DEPTH, b, B, X, U, A, Bd);
        cp_async_commit();
        stage = (stage + 1 == DEPTH) ? 0 : stage + 1;
    }
}

// K2, small-batch variant for n = 4, m = 1 (the double-pendulum headline case): FOUR lanes per
// trajectory.  With a few thousand trajectories the one-thread-per-trajectory scan runs one warp per
// SM and is bound by dependent-instruction issue (~490 instructions per step in one thread).  Here
// lane j of a 4-lane group owns column j: it computes Y[:,j] = V_xx A[:,j], Q_xx[:,j] = l_xx[:,j] +
// A' Y[:,j], Q_ux[j] = B' Y[:,j], Q_x[j], K[j] and the new V_xx[:,j], V_x[j]; Q_uu, Q_u, k are cheap and
// computed redundantly.  Every lane keeps a full copy of V_xx, V_x, re-assembled each step through a
// shared-memory exchange (two __syncwarp per step).  A warp holds 8 trajectories; their A_t,B_t,x_t,u_t
// (25 values each) arrive through a DEPTH-deep cp.async ring filled cooperatively (7 LDGSTS per step
// per warp, 64-byte global segments).  ~115 instructions per lane per step.
template <typename T, int DEPTH>
__global__ void __launch_bounds__(32)
backward_n4m1_lanes_kernel(const __grid_constant__ QuadCost<T, 4, 1> qc, int N, int B, const T *__restrict__ X,
                           const T *__restrict__ U, const T *__restrict__ A, const T *__restrict__ Bd,
                           T *__restrict__ K, T *__restrict__ k, const int *__restrict__ active,
                           const unsigned int *__restrict__ gate, const T *__restrict__ mu)
{
    constexpr int n = 4, L = 25, LP = 26, SLOTS = 8;       // LP: padded slot stride (bank-conflict free LDS.128)
    extern __shared__ __align__(16) unsigned char lanes_raw[];
    T *ring = reinterpret_cast<T *>(lanes_raw);             // [DEPTH][SLOTS][LP]
    T *exQ = ring + DEPTH * SLOTS * LP;                     // [SLOTS][4]   Q_ux exchange
    T *exV = exQ + SLOTS * 4;                               // [SLOTS][20]  V_xx (row-major 16) + V_x (4)
    if (gate && *gate == 0u) return;
    const int lane = threadIdx.x, s = lane >> 2, j = lane & 3;
    const int b_raw = blockIdx.x * SLOTS + s;
    const bool valid = b_raw < B && (!active || active[b_raw < B ? b_raw : B - 1] != 0);
    if (__ballot_sync(0xffffffffu, valid) == 0u) return;
    const int b = b_raw < B ? b_raw : B - 1;                // clamped: out-of-range slots compute on a copy, never store

    // cooperative fill of one ring stage: element e = row * 8 + slot, 32 elements per LDGSTS
    const int f_slot = lane & 7, f_row0 = lane >> 3;
    const int f_b = min(blockIdx.x * SLOTS + f_slot, B - 1);
    auto issue = [&](int stage, int t) {
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int row = f_row0 + 4 * i;
            if (row < L) {
                const T *src = row < 16 ? A + ((size_t)t * 16 + row) * B + f_b
                             : row < 20 ? Bd + ((size_t)t * 4 + (row - 16)) * B + f_b
                             : row < 24 ? X + ((size_t)t * 4 + (row - 20)) * B + f_b
                                        : U + (size_t)t * B + f_b;
                cp_async<sizeof(T)>(ring + (stage * SLOTS + f_slot) * LP + row, src);
            }
        }
    };
#pragma unroll
    for (int st = 0; st < DEPTH; ++st) {
        if (N - 1 - st >= 0) issue(st, N - 1 - st);
        cp_async_commit();
    }
    // per-lane constants: row j of dt*Qs (for l_x[j]) and column j of dt*Qs (for l_xx[:,j])
    T qrow[n], qcol[n];
#pragma unroll
    for (int i = 0; i < n; ++i) { qrow[i] = qc.Qs[j][i] * qc.dt; qcol[i] = qc.Qs[i][j] * qc.dt; }
    const T xtj[n] = { qc.xt[0], qc.xt[1], qc.xt[2], qc.xt[3] };
    const T luu = qc.Rs[0][0] * qc.dt;
    const T mu_b = mu ? mu[b] : T(0);                                    // regularisation (RegArgs), 0 in the reference
    T Vx[n], Vxx[n][n];
    {
        T xN[n];
#pragma unroll
        for (int i = 0; i < n; ++i) xN[i] = X[((size_t)N * n + i) * B + b];
        qc.terminal_grad(xN, Vx);                                        // iLQR_class.py:136-138
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int c = 0; c < n; ++c) Vxx[i][c] = qc.Qfs[i][c];
    }
    int stage = 0;
    for (int t = N - 1; t >= 0; --t) {
        cp_async_wait<DEPTH - 1>();
        __syncwarp();                                                    // other lanes' copies are visible
        const T *in = ring + (stage * SLOTS + s) * LP;
        T Am[n][n], Bv[n], x[n], Acol[n];
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int c = 0; c < n; ++c) Am[i][c] = in[i * 4 + c];
            Bv[i] = in[16 + i];
            x[i] = in[20 + i];
            Acol[i] = in[i * 4 + j];
        }
        const T u = in[24];
        // Y = V_xx A[:,j] ; Q_xx[:,j] = l_xx[:,j] + A' Y ; Q_ux[j] = B' Y          (iLQR_class.py:102-103)
        T Y[n], Qxxc[n], Quxj = T(0), Qxj = T(0), Qu = T(0), Quu = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T sum = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) sum += Vxx[i][l] * Acol[l];
            Y[i] = sum;
        }
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T sum = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) sum += Am[l][i] * Y[l];
            Qxxc[i] = qcol[i] + sum;
            Quxj += Bv[i] * Y[i];
        }
        // Q_uu = l_uu + B' V_xx B, Q_u = l_u + B' V_x (redundant in the 4 lanes) ; Q_x[j]   (:100-101,104)
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T vb = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) vb += Vxx[i][l] * Bv[l];
            Quu += Bv[i] * vb;
            Qu += Bv[i] * Vx[i];
            Qxj += Acol[i] * Vx[i];
        }
        Quu += luu;
        Quu += mu_b;
        Qu += luu * u;
        T lxj = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) lxj += qrow[i] * (x[i] - xtj[i]);
        Qxj += lxj;
        const T r = -rcp_t(Quu);                                         // (:109-110)
        const T Kj = Quxj * r, kk = Qu * r;
        const T Vxj = Qxj + Kj * Qu;                                     // (:113)
        exQ[s * 4 + j] = Quxj;
        __syncwarp();
        T Quxa[n];
#pragma unroll
        for (int i = 0; i < n; ++i) Quxa[i] = exQ[s * 4 + i];
#pragma unroll
        for (int i = 0; i < n; ++i) exV[s * 20 + i * 4 + j] = Qxxc[i] + Quxa[i] * Kj;     // V_xx[:,j]   (:114)
        exV[s * 20 + 16 + j] = Vxj;
        if (valid) {
            K[((size_t)t * n + j) * B + b] = Kj;
            if (j == 0) k[(size_t)t * B + b] = kk;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int c = 0; c < n; ++c) Vxx[i][c] = exV[s * 20 + i * 4 + c];
            Vx[i] = exV[s * 20 + 16 + i];
        }
        if (t - DEPTH >= 0) issue(stage, t - DEPTH);                     // every lane is past its reads of this stage
        cp_async_commit();
        stage = (stage + 1 == DEPTH) ? 0 : stage + 1;
    }
}

// K2 for the synthetic LTV system (n = 12, m = 4; BASELINE.json config 4).  One thread per trajectory
// would need V_xx alone in 288 registers, so SIXTEEN lanes share a trajectory: lane c owns column c of
// [A_t | B_t] (12 + 4 columns).  A_t = I + dt (Ac + w_t E) is generated in the kernel from the constants and
// the trajectory's phase -- it is never read from (or written to) HBM -- and B_t = dt Bc is constant.
// Per step lane c computes
//     W[:,c]  = V_xx [A|B][:,c]                    (V_xx read from shared memory, 16-byte broadcasts)
//     G[:,c]  = [A|B]' W[:,c]                       (A' from shared memory, B' constant)
//               -> c < 12: Q_xx[:,c], Q_ux[:,c]      c >= 12: Q_uu[:,c-12]          (iLQR_class.py:102-104)
//     Q_x[c] / Q_u[c-12]                                                             (:100-101)
// then every lane factors the 4x4 Q_uu (LU, partial pivoting, as the reference's solve) and solves for its
// own right-hand side: K[:,c] (c < 12) or k (:109-110), and writes its column of the new V_xx and V_x[c]
// (:113-114).  TPB trajectories per block; the 52 gain values of a step go through a shared-memory stage so
// that every global store is a row of TPB consecutive trajectories (full 128-byte lines for TPB = 16); x_t,
// u_t arrive the same way, prefetched one step ahead.  One block barrier per step.
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

#ifndef ILQR_LTV_MINBLOCKS
#define ILQR_LTV_MINBLOCKS 2
#endif
template <typename T, int TPB>
__global__ void __launch_bounds__(TPB * 16, ILQR_LTV_MINBLOCKS)
backward_ltv_kernel(const __grid_constant__ LtvSys<T> sys, const __grid_constant__ QuadCost<T, 12, 4> qc, int N, int B,
                    const T *__restrict__ phi, const T *__restrict__ X, const T *__restrict__ U, T *__restrict__ K,
                    T *__restrict__ k, const int *__restrict__ active, const unsigned int *__restrict__ gate,
                    const T *__restrict__ mu)
{
    constexpr int n = 12, m = 4, NT = TPB * 16, ROWS = n * m + m;   // 52 gain rows per step
    using V2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char ltv_raw[];
    T *sm = reinterpret_cast<T *>(ltv_raw);
    T *VxxS = sm;                       // [TPB][12][12]  row-major V_xx
    T *ATS = VxxS + TPB * 144;          // [TPB][12][12]  ATS[i][l] = A[l][i]
    T *VxS = ATS + TPB * 144;           // [TPB][12]
    T *QuxS = VxS + TPB * 12;           // [TPB][4][12]
    T *QuuS = QuxS + TPB * 48;          // [TPB][4][4]
    T *QuS = QuuS + TPB * 16;           // [TPB][4]
    T *xsS = QuS + TPB * 4;             // [2][TPB][16]   x_t (12), u_t (4), double buffered
    T *KS = xsS + 2 * TPB * 16;         // [2][52][TPB]   gain stage, double buffered
    T *BdT = KS + 2 * ROWS * TPB;       // [4][12]        BdT[j][l] = dt Bc[l][j]
    T *QsS = BdT + 48;                  // [12][12]       symmetrised Q
    T *RsS = QsS + 144;                 // [4][4]
    T *AcT = RsS + 16;                  // [12][12]       AcT[c][l] = Ac[l][c]
    T *ET = AcT + 144;                  // [12][12]       ET[c][l]  = E[l][c]
    __shared__ int vflag[TPB];
    if (gate && *gate == 0u) return;
    const int tid = threadIdx.x, s = tid >> 4, c = tid & 15;
    const int b0 = blockIdx.x * TPB;
    const int b_raw = b0 + s;
    const bool valid = b_raw < B && (!active || active[b_raw] != 0);
    if (__syncthreads_or(valid) == 0) return;
    const int b = b_raw < B ? b_raw : B - 1;        // out-of-range / inactive slots compute on a copy, never store
    if (c == 0) vflag[s] = valid;
    for (int e = tid; e < 48; e += NT) BdT[e] = qc.dt * sys.Bc[e % 12][e / 12];
    for (int e = tid; e < 144; e += NT) {
        QsS[e] = qc.Qs[e / 12][e % 12];
        AcT[e] = sys.Ac[e % 12][e / 12];
        ET[e] = sys.E[e % 12][e / 12];
    }
    if (tid < 16) RsS[tid] = qc.Rs[tid >> 2][tid & 3];
    // staged loads: thread (r, bb) fetches row r (x_0..x_11, u_0..u_3) of trajectory b0 + bb
    const int ld_r = tid / TPB, ld_bb = tid % TPB;
    const int ld_b = min(b0 + ld_bb, B - 1);
    auto fetch = [&](int t) -> T {
        if (ld_r < n) return X[((size_t)t * n + ld_r) * B + ld_b];
        return t < N ? U[((size_t)t * m + (ld_r - n)) * B + ld_b] : T(0);
    };
    // column c of [A_t | B_t]: rebuilt every step from AcT/ET (c < 12), constant dt Bc[:,c-12] otherwise
    T M[n];
#pragma unroll
    for (int l = 0; l < n; ++l) M[l] = c < n ? T(0) : qc.dt * sys.Bc[l][c - n];
    const T ph = phi ? phi[b] : T(0);
    const T mu_b = mu ? mu[b] : T(0);                                    // regularisation (RegArgs), 0 in the reference
    T w = sys.time_scalar(N - 1, ph);
    T *Vxx = VxxS + s * 144, *AT = ATS + s * 144, *Vx = VxS + s * 12, *Qux = QuxS + s * 48, *Quu = QuuS + s * 16,
      *Qu = QuS + s * 4;
    // terminal condition (iLQR_class.py:136-138): V_x = Q_f (x_N - x_target), V_xx = Q_f
    xsS[ld_bb * 16 + ld_r] = fetch(N);
    __syncthreads();
    if (c < n) {
        T g = T(0);
#pragma unroll
        for (int j = 0; j < n; ++j) g += qc.Qfs[c][j] * (xsS[s * 16 + j] - qc.xt[j]);
        Vx[c] = g;
#pragma unroll
        for (int i = 0; i < n; ++i) Vxx[i * 12 + c] = qc.Qfs[i][c];
    }
    T pre = fetch(N - 1);
    __syncthreads();
    xsS[TPB * 16 + ld_bb * 16 + ld_r] = pre;        // buffer 1 holds step N-1 (buffer index = (N - t) & 1)
    __syncthreads();
    for (int t = N - 1; t >= 0; --t) {
        const int buf = (N - t) & 1;
        const T *xs = xsS + buf * TPB * 16 + s * 16;
        if (t > 0) pre = fetch(t - 1);
        if (c < n) {
#pragma unroll
            for (int l = 0; l < n; l += 2) {
                const V2 a = *reinterpret_cast<const V2 *>(AcT + c * 12 + l), e = *reinterpret_cast<const V2 *>(ET + c * 12 + l);
                M[l] = ((l == c) ? T(1) : T(0)) + qc.dt * (a.x + w * e.x);
                M[l + 1] = ((l + 1 == c) ? T(1) : T(0)) + qc.dt * (a.y + w * e.y);
                *reinterpret_cast<V2 *>(AT + c * 12 + l) = V2{M[l], M[l + 1]};
            }
        }
        __syncwarp();
        if (t > 0) w = sys.time_scalar(t - 1, ph);      // next step's scalar: independent work for the solve's latency
        // W[:,c] = V_xx [A|B][:,c]
        T W[n];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            T acc = T(0);
#pragma unroll
            for (int l = 0; l < n; l += 2) {
                const V2 v = *reinterpret_cast<const V2 *>(Vxx + i * 12 + l);
                acc += v.x * M[l];
                acc += v.y * M[l + 1];
            }
            W[i] = acc;
        }
        // G = [A|B]' W[:,c]
        T G[n + m];
#pragma unroll
        for (int r = 0; r < n + m; ++r) {
            const T *row = r < n ? AT + r * 12 : BdT + (r - n) * 12;
            T acc = T(0);
#pragma unroll
            for (int l = 0; l < n; l += 2) {
                const V2 v = *reinterpret_cast<const V2 *>(row + l);
                acc += v.x * W[l];
                acc += v.y * W[l + 1];
            }
            G[r] = acc;
        }
        // q = [A|B][:,c]' V_x ; cost gradient entry of this lane
        T q = T(0);
#pragma unroll
        for (int l = 0; l < n; ++l) q += M[l] * Vx[l];
        T rhs[m][1], Qc;
        if (c < n) {
            T g = T(0);
            if (qc.diag) g = QsS[c * 12 + c] * (xs[c] - qc.xt[c]);
            else {
#pragma unroll
                for (int j = 0; j < n; ++j) g += QsS[c * 12 + j] * (xs[j] - qc.xt[j]);
            }
            Qc = g * qc.dt + q;                                          // Q_x[c]
#pragma unroll
            for (int i = 0; i < n; ++i) G[i] = QsS[i * 12 + c] * qc.dt + G[i];          // Q_xx[:,c]
#pragma unroll
            for (int j = 0; j < m; ++j) { Qux[j * 12 + c] = G[n + j]; rhs[j][0] = G[n + j]; }
        } else {
            const int jj = c - n;
            T g = T(0);
            if (qc.diag) g = RsS[jj * 4 + jj] * xs[n + jj];
            else {
#pragma unroll
                for (int i = 0; i < m; ++i) g += RsS[jj * 4 + i] * xs[n + i];
            }
            Qc = g * qc.dt + q;                                          // Q_u[c-12]
            Qu[jj] = Qc;
#pragma unroll
            for (int i = 0; i < m; ++i)
                Quu[i * 4 + jj] = (RsS[i * 4 + jj] * qc.dt + G[n + i]) + (i == jj ? mu_b : T(0));            // Q_uu[:,c-12]
        }
        __syncwarp();
        // K[:,c] = -Q_uu^-1 Q_ux[:,c] (c < 12) ; k = -Q_uu^-1 Q_u (lanes 12..15, lane 12+j keeps k[j])
        T Lm[m][m], Quv[m];
#pragma unroll
        for (int i = 0; i < m; ++i) {
#pragma unroll
            for (int j = 0; j < m; ++j) Lm[i][j] = Quu[i * 4 + j];
            Quv[i] = Qu[i];
            if (c >= n) rhs[i][0] = Quv[i];
        }
        lu_solve_inplace<m, 1, T, true>(Lm, rhs);
        T *ks = KS + ((N - t) & 1) * ROWS * TPB;
        if (c < n) {
            // V_xx[:,c] = Q_xx[:,c] + Q_ux' K[:,c] ; V_x[c] = Q_x[c] + K[:,c]' Q_u
            T Kc[m];
#pragma unroll
            for (int j = 0; j < m; ++j) Kc[j] = -rhs[j][0];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < m; ++j) acc += Qux[j * 12 + i] * Kc[j];
                G[i] += acc;
            }
            T vx = T(0);
#pragma unroll
            for (int j = 0; j < m; ++j) vx += Kc[j] * Quv[j];
            vx = Qc + vx;
#pragma unroll
            for (int i = 0; i < n; ++i) Vxx[i * 12 + c] = G[i];
            Vx[c] = vx;
#pragma unroll
            for (int j = 0; j < m; ++j) ks[(j * n + c) * TPB + s] = Kc[j];
        } else {
            ks[(n * m + (c - n)) * TPB + s] = -rhs[c - n][0];
        }
        if (t > 0) xsS[(buf ^ 1) * TPB * 16 + ld_bb * 16 + ld_r] = pre;
        __syncthreads();
        // coalesced store of the step's gains: rows of TPB consecutive trajectories
        for (int e = tid; e < ROWS * TPB; e += NT) {
            const int row = e / TPB, bb = e % TPB;
            if (vflag[bb]) {
                if (row < n * m) K[((size_t)t * n * m + row) * B + b0 + bb] = ks[e];
                else k[((size_t)t * m + (row - n * m)) * B + b0 + bb] = ks[e];
            }
        }
    }
}

// K3.  One thread per (alpha, b); b fastest so loads of the shared nominal/gains coalesce and are
// served once from L2 for all alphas.
template <typename T, int n, int m>
struct FwdIn { T xo[n], uo[m], kk[m], K[m][n]; };

template <typename T, int n, int m>
ILQR_DEV void fwd_load(FwdIn<T, n, m> &d, int t, int b, int B, const T *__restrict__ X, const T *__restrict__ U,
                       const T *__restrict__ k, const T *__restrict__ K)
{
#pragma unroll
    for (int i = 0; i < n; ++i) d.xo[i] = X[((size_t)t * n + i) * B + b];
#pragma unroll
    for (int j = 0; j < m; ++j) {
        d.uo[j] = U[((size_t)t * m + j) * B + b];
        d.kk[j] = k[((size_t)t * m + j) * B + b];
#pragma unroll
        for (int i = 0; i < n; ++i) d.K[j][i] = K[(((size_t)t * m + j) * n + i) * B + b];
    }
}

// one step of the forward pass: control law (iLQR_class.py:181-182), store, stage cost (:187), dynamics (:185)
template <int INTEG, class Sys, class Cost, typename T>
ILQR_DEV void rollout_step(const Sys &sys, const Cost &qc, const FwdIn<T, Sys::N, Sys::M> &in,
                           T alpha, int t, int bw, int B, T phi, T *x, T &cost, T *__restrict__ Xw,
                           T *__restrict__ Uw)
{
    constexpr int n = Sys::N, m = Sys::M;
    T u[m], xn[n];
#pragma unroll
    for (int j = 0; j < m; ++j) {
        T s = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) s += in.K[j][i] * (x[i] - in.xo[i]);
        u[j] = in.uo[j] + alpha * in.kk[j] + s;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) Xw[((size_t)t * n + i) * B + bw] = x[i];
#pragma unroll
    for (int j = 0; j < m; ++j) Uw[((size_t)t * m + j) * B + bw] = u[j];
    cost += qc.stage(x, u);
    step<INTEG>(sys, qc.dt, x, u, xn, sys.time_scalar(t, phi));
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = xn[i];
}

#ifdef ILQR_ROLLOUT_LB
#define ILQR_ROLLOUT_BOUNDS __launch_bounds__(128, ILQR_ROLLOUT_LB)
#else
#define ILQR_ROLLOUT_BOUNDS
#endif
template <class Sys, class Cost, int INTEG, typename T>
__global__ void ILQR_ROLLOUT_BOUNDS rollout_kernel(const __grid_constant__ Sys sys, const __grid_constant__ Cost qc,
                               int N, int B, int n_alpha, const __grid_constant__ AlphaList alphas,
                               const T *__restrict__ phi,
                               const T *__restrict__ x0, const T *__restrict__ X_old, const T *__restrict__ U_old,
                               const T *__restrict__ k, const T *__restrict__ K, T *__restrict__ Xc,
                               T *__restrict__ Uc, T *__restrict__ cost_alpha, const int *__restrict__ active,
                               const unsigned int *__restrict__ gate, const T *__restrict__ cost_ref,
                               const __grid_constant__ SpecArgs sp, const int *__restrict__ list,
                               const unsigned int *__restrict__ list_count)
{
    constexpr int n = Sys::N, m = Sys::M;
    if (gate && *gate == 0u) return;
    // Warp w of the grid handles step size (w % n_alpha) of trajectory group (w / n_alpha): the warps that
    // re-read the same nominal trajectory and gains run next to each other, so at large batches those
    // reads come from L1/L2 instead of once per step size from HBM.
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t wg = gid >> 5, ngrp = ((size_t)B + 31) >> 5;
    int ai, b, bw;        // bw: column of the candidate slabs / cost_alpha this thread writes
    if (wg < ngrp * n_alpha) {
        ai = (int)(wg % n_alpha);
        const unsigned int idx = (unsigned int)(wg / n_alpha) * 32u + (threadIdx.x & 31u);
        if (list) {                                                      // lazy wave: compacted trajectory list;
            if (idx >= min(*list_count, (unsigned int)B)) return;        // results stored at the list position
            b = list[idx];
        } else {
            if (idx >= (unsigned int)B) return;
            b = (int)idx;
        }
        bw = (int)idx;
    } else {                                                             // speculative extra threads
        const size_t e = gid - ngrp * n_alpha * 32;
        if (list || sp.cap == 0 || e >= (size_t)sp.cap * sp.n2) return;
        const int q = (int)(e % sp.cap);
        const unsigned int cnt = min(*sp.count_cur, (unsigned int)sp.cap);
        if ((unsigned int)q >= cnt) return;
        b = sp.list_cur[q];
        bw = b;
        ai = n_alpha + (int)(e / sp.cap);
    }
    if (active && !active[b]) return;
    const T alpha = (T)alphas.a[ai];
    T *Xw = Xc + (size_t)ai * (N + 1) * n * B, *Uw = Uc + (size_t)ai * N * m * B;
    T x[n], cost = T(0);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = x0[(size_t)i * B + b];
    // Early rejection: with non-negative diagonal weights every stage cost is >= 0 and the running sum
    // is monotone in floating point, so once it exceeds the cost to beat the acceptance test
    // `cost_new <= cost` (iLQR_class.py:289) is already decided.  Exactly the reference's decision,
    // without rolling a diverged candidate to the end of the horizon.
    const T ph = phi ? phi[b] : T(0);
    const bool can_reject = cost_ref != nullptr && qc.monotone;
    const T c_ref = can_reject ? cost_ref[b] : T(0);
    // the time loop is unrolled by two over a ping-pong pair of input buffers so that the next step's
    // nominal/gains are in flight during the current step without register-to-register copies
#ifndef ILQR_UNROLL2
#define ILQR_UNROLL2 1
#endif
#ifndef ILQR_REJECT
#define ILQR_REJECT 0
#endif
    if constexpr (n > 4) {
        // large state (n = 12, m = 4): the nominal and the 48 gains of a step are consumed as they arrive;
        // a register-resident prefetch buffer would spill, and these batches have enough warps per SM to
        // cover the load latency by occupancy
        for (int t = 0; t < N; ++t) {
            T dx[n], u[m], xn[n];
#pragma unroll
            for (int i = 0; i < n; ++i) dx[i] = x[i] - X_old[((size_t)t * n + i) * B + b];
#pragma unroll
            for (int j = 0; j < m; ++j) {
                T s = T(0);
#pragma unroll
                for (int i = 0; i < n; ++i) s += K[(((size_t)t * m + j) * n + i) * B + b] * dx[i];
                u[j] = U_old[((size_t)t * m + j) * B + b] + alpha * k[((size_t)t * m + j) * B + b] + s;
            }
#pragma unroll
            for (int i = 0; i < n; ++i) Xw[((size_t)t * n + i) * B + bw] = x[i];
#pragma unroll
            for (int j = 0; j < m; ++j) Uw[((size_t)t * m + j) * B + bw] = u[j];
            cost += qc.stage(x, u);
            step<INTEG>(sys, qc.dt, x, u, xn, sys.time_scalar(t, ph));
#pragma unroll
            for (int i = 0; i < n; ++i) x[i] = xn[i];
        }
    } else {
#if ILQR_UNROLL2
    FwdIn<T, n, m> in0, in1;
    fwd_load(in0, 0, b, B, X_old, U_old, k, K);
    for (int t = 0; t < N; t += 2) {
        if (t + 1 < N) fwd_load(in1, t + 1, b, B, X_old, U_old, k, K);
        rollout_step<INTEG>(sys, qc, in0, alpha, t, bw, B, ph, x, cost, Xw, Uw);
        if (t + 1 >= N) break;
        if (t + 2 < N) fwd_load(in0, t + 2, b, B, X_old, U_old, k, K);
        rollout_step<INTEG>(sys, qc, in1, alpha, t + 1, bw, B, ph, x, cost, Xw, Uw);
#if ILQR_REJECT
        if (can_reject && !(cost <= c_ref)) {
            cost_alpha[(size_t)ai * B + bw] = cost;                       // already > cost to beat (or NaN): rejected
            return;
        }
#endif
    }
#else
    FwdIn<T, n, m> in0, in1;
    fwd_load(in0, 0, b, B, X_old, U_old, k, K);
    for (int t = 0; t < N; ++t) {
        if (t + 1 < N) fwd_load(in1, t + 1, b, B, X_old, U_old, k, K);
        rollout_step<INTEG>(sys, qc, in0, alpha, t, bw, B, ph, x, cost, Xw, Uw);
        in0 = in1;
#if ILQR_REJECT
        if (can_reject && !(cost <= c_ref)) {
            cost_alpha[(size_t)ai * B + bw] = cost;
            return;
        }
#endif
    }
#endif
    }
#pragma unroll
    for (int i = 0; i < n; ++i) Xw[((size_t)N * n + i) * B + bw] = x[i];
    cost_alpha[(size_t)ai * B + bw] = cost + qc.terminal(x);              // :245
}

// after the alpha = 0 rollout (iLQR_class.py:257-263): everything active, candidate 0 is the nominal
template <typename T>
__global__ void init_kernel(int B, const T *__restrict__ cost_alpha, T *__restrict__ cost, int *__restrict__ winner,
                            int *__restrict__ active, int *__restrict__ iters, int *__restrict__ status, int maxiter,
                            Control *ctl, T *__restrict__ tr_cost, T *__restrict__ mu, T mu_init)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) ctl->n_active[0] = maxiter > 0 ? (unsigned)B : 0u;
    if (b >= B) return;
    cost[b] = cost_alpha[b];
    if (tr_cost) tr_cost[b] = cost_alpha[b];
    winner[b] = 0;
    active[b] = maxiter > 0;
    iters[b] = 0;
    status[b] = maxiter > 0 ? ILQR_ST_RUNNING : ILQR_ST_MAXITER;
    if (mu) mu[b] = mu_init;
}

// K4.  iLQR_class.py:265-271 (convergence), :281-307 (first acceptable alpha, failure => stop)
// The line search may be split in two waves of step sizes (alphas [0,n_first) rolled out eagerly,
// [n_first,n_alpha) only for trajectories that accepted none of the first wave; see ilqr_solve).
//   wave 0: every active trajectory; tries a in [a_lo,a_hi); if none is acceptable and a second wave
//           exists (defer != nullptr) the trajectory is marked in defer[] instead of failing.
//   wave 1: the marked trajectories only; tries the remaining step sizes and finalises.
// n2_count points at the deferred-trajectory counter of this iteration (gate of the second wave).
template <typename T>
__global__ void select_kernel(int B, int a_lo, int a_hi, int wave, const T *__restrict__ cost_alpha,
                              T *__restrict__ cost, int *__restrict__ winner, int *__restrict__ active,
                              int *__restrict__ defer, int *__restrict__ iters, int *__restrict__ status, T tol,
                              int it, int maxiter, Control *ctl, unsigned int *n2_count,
                              int *__restrict__ tr_alpha, T *__restrict__ tr_cost, const __grid_constant__ SpecArgs sp,
                              const __grid_constant__ RegArgs rg)
{
    if (ctl->n_active[it] == 0u) return;
    if (wave == 1 && *n2_count == 0u) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool still = false, ran = false, deferred = false;
    if (b < B) {
        const bool mine = wave == 0 ? active[b] != 0 : defer[b] != 0;
        if (!mine) {
            if (wave == 0) winner[b] = -1;
        } else {
            ran = wave == 0;
            if (wave == 1) defer[b] = 0;
            const T c0 = cost[b];
            int w = -1;
            T cw = c0;
            // a listed trajectory had its deferred step sizes rolled out speculatively in the first wave
            const bool listed = wave == 0 && sp.cap > 0 && sp.mark[b] == it + 1;
            const int hi = listed ? a_hi + sp.n2 : a_hi;
            for (int a = a_lo; a < hi; ++a) {
                const T c = cost_alpha[(size_t)a * B + b];
                if (c <= c0) { w = a; cw = c; break; }                   // NaN compares false, as in Python
            }
            winner[b] = w;
            iters[b] = it + 1;
            if (w < 0 && wave == 0 && defer != nullptr && !listed) {
                defer[b] = 1;                                            // decided by the second wave
                deferred = true;
            } else {
                if (tr_alpha) tr_alpha[(size_t)it * B + b] = w;
                if (tr_cost) tr_cost[(size_t)(it + 1) * B + b] = cw;
                if (w < 0) {
                    if (reg_on_failure<T>(rg, b)) {                      // retry this iteration with a larger mu
                        if (it + 1 >= maxiter) { status[b] = ILQR_ST_MAXITER; active[b] = 0; }
                        else still = true;
                    } else {
                        status[b] = ILQR_ST_LS_FAILED;
                        active[b] = 0;
                    }
                } else {
                    cost[b] = cw;
                    reg_on_success<T>(rg, b);
                    if (it + 1 >= maxiter) { status[b] = ILQR_ST_MAXITER; active[b] = 0; }
                    else if (abs_t(cw - c0) <= tol) { status[b] = ILQR_ST_CONVERGED; active[b] = 0; }
                    else still = true;
                    if (still && sp.cap > 0 && w >= sp.threshold) {      // small step needed: list it for next time
                        const unsigned int pos = atomicAdd(sp.count_next, 1u);
                        if (pos < (unsigned int)sp.cap) { sp.list_next[pos] = b; sp.mark[b] = it + 2; }
                    }
                }
            }
        }
    }
    const unsigned full = 0xffffffffu;
    const unsigned ns = __popc(__ballot_sync(full, still)), nr = __popc(__ballot_sync(full, ran));
    const unsigned nd = __popc(__ballot_sync(full, deferred));
    if ((threadIdx.x & 31) == 0) {
        if (ns) atomicAdd(&ctl->n_active[it + 1], ns);
        if (nr) atomicAdd(&ctl->total_iters, (unsigned long long)nr);
        if (nd) atomicAdd(n2_count, nd);
    }
}

// K4, lazy multi-wave form (large batches).  The step sizes are split into consecutive waves
// [a_lo, a_hi).  Wave 0 covers every active trajectory; a trajectory that accepts none of a wave's step
// sizes is appended to a compacted list (warp-aggregated atomics keep a warp's entries contiguous) and
// only the listed trajectories are rolled out in the next wave.  The decision per trajectory is the
// reference's (lowest-index acceptable step size); only the amount of work changes.
template <typename T>
__global__ void select_lazy_kernel(int B, int a_lo, int a_hi, int wave, int last, const T *__restrict__ cost_alpha,
                                   T *__restrict__ cost, int *__restrict__ winner, int *__restrict__ active,
                                   int *__restrict__ iters, int *__restrict__ status, T tol, int it, int maxiter,
                                   Control *ctl, const int *__restrict__ list_in, const unsigned int *cnt_in,
                                   int *__restrict__ list_out, unsigned int *cnt_out, int *__restrict__ wslot,
                                   int *__restrict__ tr_alpha, T *__restrict__ tr_cost, const __grid_constant__ RegArgs rg)
{
    if (ctl->n_active[it] == 0u) return;
    if (wave > 0 && *cnt_in == 0u) return;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int b = gid;
    bool mine = false;
    if (wave == 0) {
        if (gid < B) {
            mine = active[gid] != 0;
            if (!mine) winner[gid] = -1;
        }
    } else if ((unsigned int)gid < min(*cnt_in, (unsigned int)B)) {
        mine = true;
        b = list_in[gid];
    }
    bool still = false, app = false;
    if (mine) {
        const T c0 = cost[b];
        int w = -1;
        T cw = c0;
        for (int a = a_lo; a < a_hi; ++a) {
            const T c = cost_alpha[(size_t)a * B + gid];                 // stored at the list position (wave 0: gid == b)
            if (c <= c0) { w = a; cw = c; break; }                       // NaN compares false, as in Python
        }
        if (w < 0 && !last) {
            app = true;                                                  // decided by a later wave
        } else {
            winner[b] = w;
            wslot[b] = gid;
            iters[b] = it + 1;
            if (tr_alpha) tr_alpha[(size_t)it * B + b] = w;
            if (tr_cost) tr_cost[(size_t)(it + 1) * B + b] = cw;
            if (w < 0) {
                if (reg_on_failure<T>(rg, b)) {                          // retry this iteration with a larger mu
                    if (it + 1 >= maxiter) { status[b] = ILQR_ST_MAXITER; active[b] = 0; }
                    else still = true;
                } else {
                    status[b] = ILQR_ST_LS_FAILED;
                    active[b] = 0;
                }
            } else {
                cost[b] = cw;
                reg_on_success<T>(rg, b);
                if (it + 1 >= maxiter) { status[b] = ILQR_ST_MAXITER; active[b] = 0; }
                else if (abs_t(cw - c0) <= tol) { status[b] = ILQR_ST_CONVERGED; active[b] = 0; }
                else still = true;
            }
        }
    }
    const unsigned full = 0xffffffffu, lane = threadIdx.x & 31;
    const unsigned ma = __ballot_sync(full, app);
    if (ma) {
        const int leader = __ffs(ma) - 1;
        unsigned int base = 0;
        if ((int)lane == leader) base = atomicAdd(cnt_out, (unsigned int)__popc(ma));
        base = __shfl_sync(full, base, leader);
        if (app) list_out[base + __popc(ma & ((1u << lane) - 1u))] = b;
    }
    const unsigned ns = __popc(__ballot_sync(full, still));
    const unsigned nr = __popc(__ballot_sync(full, mine && wave == 0));
    if (lane == 0) {
        if (ns) atomicAdd(&ctl->n_active[it + 1], ns);
        if (nr) atomicAdd(&ctl->total_iters, (unsigned long long)nr);
    }
}

// winner only (ilqr_forward_linesearch)
template <typename T>
__global__ void winner_kernel(int B, int n_alpha, const T *__restrict__ cost_alpha, const T *__restrict__ cost,
                              int *__restrict__ winner)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const T c0 = cost[b];
    int w = -1;
    for (int a = 0; a < n_alpha; ++a)
        if (cost_alpha[(size_t)a * B + b] <= c0) { w = a; break; }
    winner[b] = w;
}

// materialised cost expansion (system_base.py:212-219); one thread per (t,b), t in [0,N]
template <class Cost, typename T, int n, int m>
__global__ void cost_expansion_kernel(const __grid_constant__ Cost qc, int N, int B, const T *__restrict__ X,
                                      const T *__restrict__ U, T *__restrict__ l, T *__restrict__ lx,
                                      T *__restrict__ lu, T *__restrict__ lxx, T *__restrict__ luu,
                                      T *__restrict__ lux, T *__restrict__ lf, T *__restrict__ lfx,
                                      T *__restrict__ lfxx)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)(N + 1) * B) return;
    const int t = (int)(gid / B), b = (int)(gid % B);
    T x[n], u[m];
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = X[((size_t)t * n + i) * B + b];
    if (t == N) {
        if (lf) lf[b] = qc.terminal(x);
        T g[n], H[n][n];
        if constexpr (Cost::QUADRATIC) {
            qc.terminal_grad(x, g);
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int j = 0; j < n; ++j) H[i][j] = qc.Qfs[i][j];
        } else {
            qc.terminal_expand(x, g, H);
        }
#pragma unroll
        for (int i = 0; i < n; ++i) {
            if (lfx) lfx[(size_t)i * B + b] = g[i];
#pragma unroll
            for (int j = 0; j < n; ++j)
                if (lfxx) lfxx[((size_t)i * n + j) * B + b] = H[i][j];
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < m; ++j) u[j] = U[((size_t)t * m + j) * B + b];
    T gx[n], gu[m], hxx[n][n], huu[m][m], hux[m][n];
    if constexpr (Cost::QUADRATIC) {
        qc.grad(x, u, gx, gu);
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) hxx[i][j] = qc.Qs[i][j] * qc.dt;
#pragma unroll
        for (int i = 0; i < m; ++i) {
#pragma unroll
            for (int j = 0; j < m; ++j) huu[i][j] = qc.Rs[i][j] * qc.dt;
#pragma unroll
            for (int j = 0; j < n; ++j) hux[i][j] = T(0);
        }
    } else {
        qc.expand(x, u, gx, gu, hxx, huu, hux);
    }
    if (l) l[(size_t)t * B + b] = qc.stage(x, u);
#pragma unroll
    for (int i = 0; i < n; ++i) {
        if (lx) lx[((size_t)t * n + i) * B + b] = gx[i];
#pragma unroll
        for (int j = 0; j < n; ++j)
            if (lxx) lxx[(((size_t)t * n + i) * n + j) * B + b] = hxx[i][j];
    }
#pragma unroll
    for (int i = 0; i < m; ++i) {
        if (lu) lu[((size_t)t * m + i) * B + b] = gu[i];
#pragma unroll
        for (int j = 0; j < m; ++j)
            if (luu) luu[(((size_t)t * m + i) * m + j) * B + b] = huu[i][j];
#pragma unroll
        for (int j = 0; j < n; ++j)
            if (lux) lux[(((size_t)t * m + i) * n + j) * B + b] = hux[i][j];
    }
}

// run_iLQR_UA_MPC.py:157,168
template <typename T>
__global__ void mpc_shift_kernel(int N, int m, int B, T *__restrict__ U, T *__restrict__ u0)
{
    // one thread per (j,b): walks the horizon so the in-place shift needs no second buffer
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * B) return;
    const int j = gid / B, b = gid % B;
    T prev = U[((size_t)0 * m + j) * B + b];
    if (u0) u0[(size_t)j * B + b] = prev;
    for (int t = 0; t + 1 < N; ++t) {
        const T v = U[((size_t)(t + 1) * m + j) * B + b];
        U[((size_t)t * m + j) * B + b] = v;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

template <typename T> PendulumSys<T> make_pendulum(const ilqr_problem_t &p)
{
    PendulumSys<T> s;
    s.gl = (T)(p.phys[0] / p.phys[1]);
    s.d = (T)p.phys[2];
    return s;
}

template <typename T, int M> DoublePendulumSys<T, M> make_double(const ilqr_problem_t &p)
{
    const double g = p.phys[0], m1 = p.phys[1], m2 = p.phys[2], l1 = p.phys[3], l2 = p.phys[4];
    const double d1 = p.phys[5], d2 = p.phys[6], th1 = p.phys[7], th2 = p.phys[8];
    DoublePendulumSys<T, M> s;
    s.c = (T)(m2 * l1 * l2);
    s.m11_0 = (T)((m1 * l1 * l1) / 4 + m2 * l1 * l1 + (m2 * l2 * l2) / 4 + th1 + th2);
    s.m12_0 = (T)((m2 * l2 * l2) / 4 + th2);
    s.g1 = (T)(m2 * g * l2 / 2);
    s.g2 = (T)(m2 * g * l1 + (m1 * g * l1) / 2);
    s.d1 = (T)d1;
    s.d2 = (T)d2;
    return s;
}

template <typename T> LtvSys<T> make_ltv(const ilqr_problem_t &p)
{
    LtvSys<T> s;
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) { s.Ac[i][j] = (T)p.Ac[i * 12 + j]; s.E[i][j] = (T)p.E[i * 12 + j]; }
        for (int j = 0; j < 4; ++j) s.Bc[i][j] = (T)p.Bc[i * 4 + j];
    }
    s.amp = (T)p.ltv_amp;
    s.two_pi_over_N = (T)(2.0 * 3.14159265358979323846 / (double)p.N);
    return s;
}

template <typename T, int n, int m> QuadCost<T, n, m> make_cost(const ilqr_problem_t &p)
{
    QuadCost<T, n, m> c;
    c.dt = (T)p.dt;
    bool diag = true;
    for (int i = 0; i < n; ++i) {
        c.xt[i] = (T)p.x_target[i];
        for (int j = 0; j < n; ++j) {
            const double q = 0.5 * (p.Q[i * n + j] + p.Q[j * n + i]), qf = 0.5 * (p.Qf[i * n + j] + p.Qf[j * n + i]);
            c.Qs[i][j] = (T)q;
            c.Qfs[i][j] = (T)qf;
            if (i != j && (q != 0.0 || qf != 0.0)) diag = false;
        }
    }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            const double r = 0.5 * (p.R[i * m + j] + p.R[j * m + i]);
            c.Rs[i][j] = (T)r;
            if (i != j && r != 0.0) diag = false;
        }
    c.diag = diag ? 1 : 0;
    // early rejection in the line search needs every stage cost >= 0 in floating point
    bool mono = diag;
    for (int i = 0; i < n && mono; ++i) mono = p.Q[i * n + i] >= 0.0 && p.Qf[i * n + i] >= 0.0;
    for (int i = 0; i < m && mono; ++i) mono = p.R[i * m + i] >= 0.0;
    c.monotone = mono ? 1 : 0;
    return c;
}

struct Handle {
    ilqr_problem_t p;
    int n_alpha_eff;          // tries actually made: stops once alpha < min_alpha (iLQR_class.py:300-302)
    int n_first;              // step sizes rolled out eagerly (first wave); the rest only where needed
    int spec_cap;             // trajectories whose deferred step sizes ride along speculatively (SpecArgs)
    void *mu_user;            // optional caller buffer for the per-trajectory regularisation (ilqr_set_mu_buffer)
    int lazy;                 // large batches: lazy multi-wave line search over compacted lists (select_lazy_kernel)
    int n_waves;
    int wave_lo[ILQR_MAX_WAVES + 1];
    AlphaList alphas;
    long long launches;
    int last_cuda;
    unsigned int *h_flag;     // pinned, for the pipelined early-exit poll
    int *tr_alpha;            // optional per-iteration trace (ilqr_set_trace)
    void *tr_cost;
    // optional per-kernel timing (ilqr_set_profiling): events chained between the launches of ilqr_solve
    int profiling;
    std::vector<cudaEvent_t> *prof_ev;
    std::vector<int> *prof_kind;          // kernel class that ran between event i and i+1
    double prof_ms[ILQR_N_KERNEL_CLASSES];
    long long prof_cnt[ILQR_N_KERNEL_CLASSES];
    cudaEvent_t ev[2];
};

static inline int grid_for(size_t threads, int bs) { return (int)((threads + bs - 1) / bs); }

// pick a block size that still spreads small batches over all 148 SMs
static inline int block_for(size_t threads)
{
    int bs = 256;
    while (bs > 32 && (threads + bs - 1) / bs < 16 * 148) bs >>= 1;   // small batches: one warp per block balances best
    return bs;
}

struct WsLayout {
    size_t ctl, A, Bd, Xc, Uc, cost_alpha, winner, wslot, active, defer, mark, lists, mu, total;
};

static size_t ctl_bytes(int maxiter)
{
    return sizeof(Control) + sizeof(unsigned int) * (3 + ILQR_MAX_WAVES) * (size_t)(maxiter + 2);
}

static WsLayout ws_layout(const ilqr_problem_t &p, int n_alpha)
{
    const size_t w = p.dtype == ILQR_F64 ? 8 : 4;
    const size_t B = p.B, N = p.N, n = p.n, m = p.m;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    WsLayout L;
    size_t off = 0;
    // n_active[maxiter + 2], then the per-iteration deferred (second-wave) counters [maxiter + 2], then the
    // per-iteration speculation-list counters [maxiter + 2], then the lazy-wave list counters
    // [maxiter + 2][ILQR_MAX_WAVES]
    L.ctl = off; off = al(off + ctl_bytes(p.maxiter));
    // the LTV model generates A_t, B_t inside its kernels: no per-trajectory linearization is stored
    const size_t lin = p.model == ILQR_LTV ? 0 : 1;
    L.A = off; off = al(off + lin * w * N * n * n * B);
    L.Bd = off; off = al(off + lin * w * N * n * m * B);
    L.Xc = off; off = al(off + w * (size_t)n_alpha * (N + 1) * n * B);
    L.Uc = off; off = al(off + w * (size_t)n_alpha * N * m * B);
    L.cost_alpha = off; off = al(off + w * (size_t)n_alpha * B);
    L.winner = off; off = al(off + 4 * B);
    L.wslot = off; off = al(off + 4 * B);
    L.active = off; off = al(off + 4 * B);
    L.defer = off; off = al(off + 4 * B);
    L.mark = off; off = al(off + 4 * B);
    L.lists = off; off = al(off + 4 * 2 * B);     // two speculation lists (capacity <= B each)
    L.mu = off; off = al(off + w * B);
    L.total = off;
    return L;
}

#define ILQR_CHECK_LAUNCH(h)                                         \
    do {                                                             \
        (h)->launches++;                                             \
        cudaError_t e_ = cudaGetLastError();                         \
        if (e_ != cudaSuccess) { (h)->last_cuda = (int)e_; return ILQR_E_CUDA; } \
    } while (0)

// dispatch on (dtype, model, integrator): calls f(T{}, sys, qc, integral_constant<int,INTEG>{})
template <typename T, class Sys, class F> static int dispatch_integ(const Handle *h, const Sys &sys, F &&f)
{
    auto qc = make_cost<T, Sys::N, Sys::M>(h->p);
    switch (h->p.integrator) {
    case ILQR_EULER: return f(T(0), sys, qc, std::integral_constant<int, EULER>{});
    case ILQR_MIDPOINT: return f(T(0), sys, qc, std::integral_constant<int, MIDPOINT>{});
    case ILQR_RK4: return f(T(0), sys, qc, std::integral_constant<int, RK4>{});
    case ILQR_BACKWARD_EULER: return f(T(0), sys, qc, std::integral_constant<int, BACKWARD_EULER>{});
    }
    return ILQR_E_INVALID;
}

#ifndef ILQR_USER_SYS
template <typename T, class F> static int dispatch_model(const Handle *h, F &&f)
{
    switch (h->p.model) {
    case ILQR_PENDULUM: return dispatch_integ<T>(h, make_pendulum<T>(h->p), f);
    case ILQR_DOUBLE_PENDULUM: return dispatch_integ<T>(h, make_double<T, 2>(h->p), f);
    case ILQR_UA_DOUBLE_PENDULUM: return dispatch_integ<T>(h, make_double<T, 1>(h->p), f);
    case ILQR_LTV: {   // forward Euler only (validated in ilqr_create)
        auto sys = make_ltv<T>(h->p);
        auto qc = make_cost<T, 12, 4>(h->p);
        return f(T(0), sys, qc, std::integral_constant<int, EULER>{});
    }
    }
    return ILQR_E_INVALID;
}
#endif

#ifdef ILQR_USER_SYS
#if ILQR_USER_F32
typedef float user_t;
#else
typedef double user_t;
#endif
template <class F> static int dispatch(const Handle *h, F &&f)
{
    UserSys<user_t> sys;
    UserCost<user_t> qc;
    qc.dt = (user_t)h->p.dt;
    return f(user_t(0), sys, qc, std::integral_constant<int, ILQR_USER_INTEG>{});
}
#else
template <class F> static int dispatch(const Handle *h, F &&f)
{
    if (h->p.dtype == ILQR_F64) return dispatch_model<double>(h, f);
    return dispatch_model<float>(h, f);
}
#endif

// ---- launch helpers -----------------------------------------------------------------------

static int launch_commit_linearize(Handle *h, const void *phi, void *X, void *U, void *A, void *Bd, const void *Xc, const void *Uc,
                                   const int *winner, const int *wslot, const int *active, int do_lin, const unsigned int *g0,
                                   const unsigned int *g1, cudaStream_t st)
{
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto integ) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        constexpr int I = decltype(integ)::value;
        const size_t threads = (size_t)(h->p.N + 1) * h->p.B;
        const int bs = 128;
        commit_linearize_kernel<Sys, I, T><<<grid_for(threads, bs), bs, 0, st>>>(
            sys, qc.dt, h->p.N, h->p.B, (const T *)phi, (T *)X, (T *)U, (T *)A, (T *)Bd, (const T *)Xc, (const T *)Uc, winner, wslot,
            active, do_lin, g0, g1);
        ILQR_CHECK_LAUNCH(h);
        return ILQR_OK;
    });
}

template <typename T, int n, int m, int DEPTH, class Cost>
static int launch_backward_depth(Handle *h, int bs, const Cost &qc, const void *X, const void *U,
                                 const void *A, const void *Bd, void *K, void *k, const int *active,
                                 const unsigned int *gate, const void *mu, cudaStream_t st)
{
    constexpr int L = n * n + n * m + n + m;
    const size_t smem = (size_t)DEPTH * L * bs * sizeof(T);
    static size_t configured = 0;           // per instantiation: opt in to > 48 KB dynamic shared memory once
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(backward_kernel<Cost, T, n, m, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
        configured = smem;
    }
    backward_kernel<Cost, T, n, m, DEPTH><<<grid_for(h->p.B, bs), bs, smem, st>>>(
        qc, h->p.N, h->p.B, (const T *)X, (const T *)U, (const T *)A, (const T *)Bd, (T *)K, (T *)k, active, gate,
        (const T *)mu);
    ILQR_CHECK_LAUNCH(h);
    return ILQR_OK;
}

static int launch_backward(Handle *h, const void *X, const void *U, const void *A, const void *Bd, void *K, void *k,
                           const int *active, const unsigned int *gate, cudaStream_t st, const void *mu = nullptr)
{
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        // small batches: one warp per block and a deep ring (latency bound); large batches: shallower
        // ring so that more warps fit per SM (HBM bound)
        if constexpr (Sys::N == 4 && Sys::M == 1 && decltype(qc)::QUADRATIC) {
            // small batches of the n=4, m=1 case: four lanes per trajectory (latency bound regime)
            const char *lanes_env = getenv("ILQR_BACKWARD_LANES");
            const bool lanes = lanes_env ? atoi(lanes_env) != 0 : h->p.B <= 32768;
            if (lanes) {
                constexpr int DEPTH = 8, SLOTS = 8, LP = 26;
                const size_t smem = sizeof(T) * (size_t)(DEPTH * SLOTS * LP + SLOTS * 4 + SLOTS * 20);
                backward_n4m1_lanes_kernel<T, DEPTH><<<grid_for(h->p.B, SLOTS), 32, smem, st>>>(
                    qc, h->p.N, h->p.B, (const T *)X, (const T *)U, (const T *)A, (const T *)Bd, (T *)K, (T *)k, active,
                    gate, (const T *)mu);
                ILQR_CHECK_LAUNCH(h);
                return ILQR_OK;
            }
        }
        if constexpr (Sys::N > 4) {
            // n = 12, m = 4: a ring stage is 208 rows; two stages of one warp fit the 227 KB limit
            return launch_backward_depth<T, Sys::N, Sys::M, 2>(h, 32, qc, X, U, A, Bd, K, k, active, gate, mu, st);
        } else {
            if (h->p.B <= 32768)
                return launch_backward_depth<T, Sys::N, Sys::M, 8>(h, 32, qc, X, U, A, Bd, K, k, active, gate, mu, st);
            return launch_backward_depth<T, Sys::N, Sys::M, 4>(h, 64, qc, X, U, A, Bd, K, k, active, gate, mu, st);
        }
    });
}

// K2 of the LTV model: A_t, B_t generated in the kernel (no linearization buffers)
static int launch_backward_ltv(Handle *h, const void *phi, const void *X, const void *U, void *K, void *k,
                               const int *active, const unsigned int *gate, cudaStream_t st, const void *mu = nullptr)
{
#ifdef ILQR_USER_SYS
    return ILQR_E_INVALID;
#else
    constexpr int TPB = 16;
    auto go = [&](auto tz) -> int {
        using T = decltype(tz);
        const size_t smem = sizeof(T) * (size_t)(504 * TPB + 48 + 144 + 16 + 288);
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(backward_ltv_kernel<T, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
            configured = true;
        }
        backward_ltv_kernel<T, TPB><<<grid_for(h->p.B, TPB), TPB * 16, smem, st>>>(
            make_ltv<T>(h->p), make_cost<T, 12, 4>(h->p), h->p.N, h->p.B, (const T *)phi, (const T *)X, (const T *)U,
            (T *)K, (T *)k, active, gate, (const T *)mu);
        ILQR_CHECK_LAUNCH(h);
        return ILQR_OK;
    };
    return h->p.dtype == ILQR_F64 ? go(double(0)) : go(float(0));
#endif
}

static int launch_rollout(Handle *h, int n_alpha, const AlphaList &al, const void *phi, const void *x0, const void *X, const void *U,
                          const void *k, const void *K, void *Xc, void *Uc, void *cost_alpha, const int *active,
                          const unsigned int *gate, const void *cost_ref, cudaStream_t st, const SpecArgs *spec = nullptr,
                          const int *list = nullptr, const unsigned int *list_count = nullptr)
{
    SpecArgs sp;
    std::memset(&sp, 0, sizeof sp);
    if (spec) sp = *spec;
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto integ) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        constexpr int I = decltype(integ)::value;
        const size_t threads = (size_t)n_alpha * (((size_t)h->p.B + 31) / 32 * 32) + (size_t)sp.cap * sp.n2;
        const char *bs_env = getenv("ILQR_ROLLOUT_BS");
        // <= 128 threads per block: the kernel's ~150 registers then leave room for 12 warps per SM
        int bs = bs_env ? atoi(bs_env) : block_for(threads);
        if (!bs_env && bs > 128) bs = 128;
        // lazy waves: the warps of one trajectory group (one per step size) share a block, hence an L1
        if (!bs_env && h->lazy && n_alpha <= 4 && threads >= (size_t)148 * 16 * 32 * n_alpha) bs = 32 * n_alpha;
        rollout_kernel<Sys, decltype(qc), I, T><<<grid_for(threads, bs), bs, 0, st>>>(
            sys, qc, h->p.N, h->p.B, n_alpha, al, (const T *)phi, (const T *)x0, (const T *)X, (const T *)U, (const T *)k,
            (const T *)K, (T *)Xc, (T *)Uc, (T *)cost_alpha, active, gate, (const T *)cost_ref, sp, list, list_count);
        ILQR_CHECK_LAUNCH(h);
        return ILQR_OK;
    });
}

static int launch_select(Handle *h, int a_lo, int a_hi, int wave, const void *ca, void *cost, int *winner, int *active,
                         int *defer, int *iters, int *status, int it, Control *ctl, unsigned int *n2c, const SpecArgs &sp,
                         const RegArgs &rg, cudaStream_t st)
{
    const int B = h->p.B, bs = 128;
    if (h->p.dtype == ILQR_F64)
        select_kernel<double><<<grid_for(B, bs), bs, 0, st>>>(B, a_lo, a_hi, wave, (const double *)ca, (double *)cost, winner,
                                                              active, defer, iters, status, h->p.tol, it, h->p.maxiter, ctl,
                                                              n2c, h->tr_alpha, (double *)h->tr_cost, sp, rg);
    else
        select_kernel<float><<<grid_for(B, bs), bs, 0, st>>>(B, a_lo, a_hi, wave, (const float *)ca, (float *)cost, winner,
                                                             active, defer, iters, status, (float)h->p.tol, it, h->p.maxiter,
                                                             ctl, n2c, h->tr_alpha, (float *)h->tr_cost, sp, rg);
    ILQR_CHECK_LAUNCH(h);
    return ILQR_OK;
}

static int launch_select_lazy(Handle *h, int a_lo, int a_hi, int wave, int last, const void *ca, void *cost, int *winner,
                              int *active, int *iters, int *status, int it, Control *ctl, const int *list_in,
                              const unsigned int *cnt_in, int *list_out, unsigned int *cnt_out, int *wslot, const RegArgs &rg,
                              cudaStream_t st)
{
    const int B = h->p.B, bs = 128;
    if (h->p.dtype == ILQR_F64)
        select_lazy_kernel<double><<<grid_for(B, bs), bs, 0, st>>>(B, a_lo, a_hi, wave, last, (const double *)ca,
                                                                   (double *)cost, winner, active, iters, status, h->p.tol,
                                                                   it, h->p.maxiter, ctl, list_in, cnt_in, list_out, cnt_out,
                                                                   wslot, h->tr_alpha, (double *)h->tr_cost, rg);
    else
        select_lazy_kernel<float><<<grid_for(B, bs), bs, 0, st>>>(B, a_lo, a_hi, wave, last, (const float *)ca,
                                                                  (float *)cost, winner, active, iters, status,
                                                                  (float)h->p.tol, it, h->p.maxiter, ctl, list_in, cnt_in,
                                                                  list_out, cnt_out, wslot, h->tr_alpha, (float *)h->tr_cost, rg);
    ILQR_CHECK_LAUNCH(h);
    return ILQR_OK;
}

// wave boundaries from per-wave sizes (clamped to the number of tries actually made)
static void set_waves(Handle *h, int n_waves, const int *sizes)
{
    h->n_waves = 0;
    h->wave_lo[0] = 0;
    int lo = 0;
    for (int v = 0; v < n_waves && v < ILQR_MAX_WAVES && lo < h->n_alpha_eff; ++v) {
        int sz = sizes[v] < 1 ? 1 : sizes[v];
        if (v == n_waves - 1 || v == ILQR_MAX_WAVES - 1 || lo + sz > h->n_alpha_eff) sz = h->n_alpha_eff - lo;
        lo += sz;
        h->wave_lo[++h->n_waves] = lo;
    }
    h->lazy = h->n_waves > 0;
}

// default schedule: lazy waves of 2,2,2,rest from 16384 trajectories up (FP64-throughput-bound rollouts),
// eager below (latency-bound rollouts).  ILQR_WAVES="2,2,2,4" / ILQR_WAVES=0 override.
static void default_waves(Handle *h)
{
    int sizes[ILQR_MAX_WAVES] = { 2, 2, 2, ILQR_MAX_ALPHAS, 0, 0, 0, 0 };
    int nw = h->p.B >= 16384 ? 4 : 0;
    if (const char *e = getenv("ILQR_WAVES")) {
        nw = 0;
        for (const char *q = e; *q && nw < ILQR_MAX_WAVES;) {
            const int v = atoi(q);
            if (v <= 0) break;
            sizes[nw++] = v;
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
    }
    set_waves(h, nw, sizes);
}

// How many of the n_alpha step sizes to roll out eagerly.  The rollout kernel is FP64-pipe bound and
// its time is set by the busiest SM sub-partition: warps = ceil(B/32) * n_alpha spread over 148 * 4
// sub-partitions.  When dropping the last (smallest, rarely needed) step sizes from the eager wave
// lowers the warps-per-sub-partition ceiling, they are deferred to a second wave that only runs for
// trajectories that accepted none of the first.  ILQR_FIRST_WAVE overrides.
static int first_wave_size(int B, int n_alpha)
{
    if (const char *e = getenv("ILQR_FIRST_WAVE")) {
        const int v = atoi(e);
        if (v >= 1 && v <= n_alpha) return v;
    }
    const long slots = 148L * 4L;
    const long wb = (B + 31) / 32;
    auto ceil_div = [](long a, long b) { return (a + b - 1) / b; };
    const long full = ceil_div(wb * n_alpha, slots);
    int best = n_alpha;
    // defer at most a third of the step sizes, and only if that removes a whole warp per sub-partition
    for (int n1 = n_alpha - 1; n1 >= 1 && n1 >= n_alpha - n_alpha / 3; --n1)
        if (ceil_div(wb * n1, slots) < full) { best = n1; break; }
    return best;
}

// Speculation capacity: the warp slots left over in the first wave once it is rounded up to whole
// warps per SM sub-partition, shared by the n2 deferred step sizes.  ILQR_SPEC_CAP overrides.
static int spec_capacity(int B, int n1, int n_alpha)
{
    const int n2 = n_alpha - n1;
    if (n2 <= 0) return 0;
    if (const char *e = getenv("ILQR_SPEC_CAP")) {
        const int v = atoi(e);
        return v < 0 ? 0 : (v > B ? B : v);
    }
    const long slots = 148L * 4L, wb = (B + 31) / 32;
    const long warps1 = wb * n1, per = (warps1 + slots - 1) / slots;
    long cap = (per * slots - warps1) * 32 / n2;
    if (cap > B) cap = B;
    return (int)cap;
}

// record a chained event after a launch of kernel class `kind` (no-op unless profiling)
static void prof_mark(Handle *h, int kind, cudaStream_t st)
{
    if (!h->profiling) return;
    const size_t i = h->prof_kind->size();
    if (h->prof_ev->size() <= i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->prof_ev->push_back(e);
    }
    cudaEventRecord((*h->prof_ev)[i], st);
    h->prof_kind->push_back(kind);
}

// fold the chained events of the last solve into per-class totals (stream must be idle)
static void prof_collect(Handle *h)
{
    if (!h->profiling) return;
    for (size_t i = 1; i < h->prof_kind->size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, (*h->prof_ev)[i - 1], (*h->prof_ev)[i]) == cudaSuccess) {
            const int kd = (*h->prof_kind)[i];
            h->prof_ms[kd] += ms;
            h->prof_cnt[kd] += 1;
        }
    }
    h->prof_kind->clear();
}

}  // namespace ilqr

using namespace ilqr;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

#ifdef ILQR_USER_SYS
const char *ilqr_version(void) { return "ilqr_b200 0.1 (sm_100a), user-defined system build"; }
#else
const char *ilqr_version(void) { return "ilqr_b200 0.1 (sm_100a)"; }
#endif

const char *ilqr_strerror(int code)
{
    switch (code) {
    case ILQR_OK: return "ok";
    case ILQR_E_INVALID: return "invalid argument or unsupported model/integrator/dimension combination";
    case ILQR_E_CUDA: return "CUDA runtime error (see ilqr_last_cuda_error)";
    case ILQR_E_WORKSPACE: return "workspace too small (see ilqr_workspace_bytes)";
    }
    return "unknown error";
}

int ilqr_create(const ilqr_problem_t *p, ilqr_handle_t *out)
{
    if (!p || !out) return ILQR_E_INVALID;
    *out = nullptr;
    if (p->integrator < ILQR_EULER || p->integrator > ILQR_BACKWARD_EULER) return ILQR_E_INVALID;
    if (p->dtype != ILQR_F64 && p->dtype != ILQR_F32) return ILQR_E_INVALID;
    int n = 0, m = 0;
    switch (p->model) {
    case ILQR_PENDULUM: n = 2; m = 1; break;
    case ILQR_DOUBLE_PENDULUM: n = 4; m = 2; break;
    case ILQR_UA_DOUBLE_PENDULUM: n = 4; m = 1; break;
    case ILQR_LTV: n = 12; m = 4; if (p->integrator != ILQR_EULER) return ILQR_E_INVALID; break;
#ifdef ILQR_USER_SYS
    case ILQR_USER:
        n = UserSys<user_t>::N; m = UserSys<user_t>::M;
        if (p->integrator != ILQR_USER_INTEG || p->dtype != (ILQR_USER_F32 ? ILQR_F32 : ILQR_F64)) return ILQR_E_INVALID;
        break;
#endif
    default: return ILQR_E_INVALID;
    }
#ifdef ILQR_USER_SYS
    if (p->model != ILQR_USER) return ILQR_E_INVALID;        // this library holds one generated model only
#endif
    if (p->n != n || p->m != m) return ILQR_E_INVALID;
    if (p->N < 1 || p->B < 1 || p->n_alpha < 1 || p->n_alpha > ILQR_MAX_ALPHAS || p->maxiter < 0) return ILQR_E_INVALID;
    if (!(p->dt > 0.0)) return ILQR_E_INVALID;
    if (p->reg_factor > 1.0 && !(p->reg_init >= 0.0 && p->reg_min > 0.0 && p->reg_max >= p->reg_min)) return ILQR_E_INVALID;
    Handle *h = new (std::nothrow) Handle;
    if (!h) return ILQR_E_INVALID;
    std::memset(h, 0, sizeof(Handle));
    h->p = *p;
    h->prof_ev = new std::vector<cudaEvent_t>();
    h->prof_kind = new std::vector<int>();
    // alpha = 1, then *= alpha_factor per failed try; tries stop once alpha < min_alpha (:279-302)
    double a = 1.0;
    int cnt = 0;
    for (int j = 0; j < p->n_alpha; ++j) {
        h->alphas.a[cnt++] = a;
        a *= p->alpha_factor;
        if (a < p->min_alpha) break;
    }
    h->n_alpha_eff = cnt;
    h->n_first = first_wave_size(p->B, cnt);
    h->spec_cap = spec_capacity(p->B, h->n_first, cnt);
    default_waves(h);
    if (cudaMallocHost((void **)&h->h_flag, 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev[1], cudaEventDisableTiming) != cudaSuccess) {
        h->last_cuda = (int)cudaGetLastError();
        delete h;
        return ILQR_E_CUDA;
    }
    *out = (ilqr_handle_t)h;
    return ILQR_OK;
}

int ilqr_destroy(ilqr_handle_t hh)
{
    Handle *h = (Handle *)hh;
    if (!h) return ILQR_E_INVALID;
    cudaEventDestroy(h->ev[0]);
    cudaEventDestroy(h->ev[1]);
    cudaFreeHost(h->h_flag);
    for (cudaEvent_t e : *h->prof_ev) cudaEventDestroy(e);
    delete h->prof_ev;
    delete h->prof_kind;
    delete h;
    return ILQR_OK;
}

size_t ilqr_workspace_bytes(ilqr_handle_t hh)
{
    Handle *h = (Handle *)hh;
    if (!h) return 0;
    return ws_layout(h->p, h->n_alpha_eff).total;
}

int64_t ilqr_launch_count(ilqr_handle_t hh) { return hh ? ((Handle *)hh)->launches : 0; }
int ilqr_last_cuda_error(ilqr_handle_t hh) { return hh ? ((Handle *)hh)->last_cuda : 0; }

int ilqr_step(ilqr_handle_t hh, int t, const void *phi, const void *x, const void *u, void *xn, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !x || !u || !xn) return ILQR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto integ) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        constexpr int I = decltype(integ)::value;
        const int bs = block_for(h->p.B);
        step_kernel<Sys, I, T><<<grid_for(h->p.B, bs), bs, 0, st>>>(sys, qc.dt, h->p.B, t, (const T *)phi, (const T *)x, (const T *)u, (T *)xn);
        ILQR_CHECK_LAUNCH(h);
        return ILQR_OK;
    });
}

int ilqr_linearize(ilqr_handle_t hh, const void *phi, const void *X, const void *U, void *A, void *Bd, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !X || !U || !A || !Bd) return ILQR_E_INVALID;
    return launch_commit_linearize(h, phi, (void *)X, (void *)U, A, Bd, nullptr, nullptr, nullptr, nullptr, nullptr, 1,
                                   nullptr, nullptr, (cudaStream_t)stream);
}

int ilqr_cost_expansion(ilqr_handle_t hh, const void *X, const void *U, void *l, void *lx, void *lu, void *lxx,
                        void *luu, void *lux, void *lf, void *lfx, void *lfxx, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !X || !U) return ILQR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        const size_t threads = (size_t)(h->p.N + 1) * h->p.B;
        const int bs = 128;
        cost_expansion_kernel<decltype(qc), T, Sys::N, Sys::M><<<grid_for(threads, bs), bs, 0, st>>>(
            qc, h->p.N, h->p.B, (const T *)X, (const T *)U, (T *)l, (T *)lx, (T *)lu, (T *)lxx, (T *)luu, (T *)lux,
            (T *)lf, (T *)lfx, (T *)lfxx);
        ILQR_CHECK_LAUNCH(h);
        return ILQR_OK;
    });
}

int ilqr_set_trace(ilqr_handle_t hh, int32_t *alpha_idx, void *cost_trace)
{
    Handle *h = (Handle *)hh;
    if (!h) return ILQR_E_INVALID;
    h->tr_alpha = alpha_idx;
    h->tr_cost = cost_trace;
    return ILQR_OK;
}

int ilqr_backward(ilqr_handle_t hh, const void *X, const void *U, const void *A, const void *Bd, void *K, void *k,
                  void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !X || !U || !A || !Bd || !K || !k) return ILQR_E_INVALID;
    return launch_backward(h, X, U, A, Bd, K, k, nullptr, nullptr, (cudaStream_t)stream);
}

int ilqr_backward_pass(ilqr_handle_t hh, const void *phi, const void *X, const void *U, void *K, void *k, void *ws,
                       size_t ws_bytes, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !ws) return ILQR_E_INVALID;
    const WsLayout L = ws_layout(h->p, h->n_alpha_eff);
    if (ws_bytes < L.total) return ILQR_E_WORKSPACE;
    char *w = (char *)ws;
    if (!X || !U || !K || !k) return ILQR_E_INVALID;
    if (h->p.model == ILQR_LTV) return launch_backward_ltv(h, phi, X, U, K, k, nullptr, nullptr, (cudaStream_t)stream);
    int rc = ilqr_linearize(hh, phi, X, U, w + L.A, w + L.Bd, stream);
    if (rc) return rc;
    return ilqr_backward(hh, X, U, w + L.A, w + L.Bd, K, k, stream);
}

int ilqr_rollout(ilqr_handle_t hh, const void *phi, const void *x0, double alpha, const void *X_old,
                 const void *U_old, const void *k, const void *K, void *X_new, void *U_new, void *cost, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !x0 || !X_old || !U_old || !k || !K || !X_new || !U_new || !cost) return ILQR_E_INVALID;
    AlphaList al;
    std::memset(&al, 0, sizeof al);
    al.a[0] = alpha;
    return launch_rollout(h, 1, al, phi, x0, X_old, U_old, k, K, X_new, U_new, cost, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int ilqr_forward_linesearch(ilqr_handle_t hh, const void *phi, const void *x0, const void *X, const void *U,
                            const void *k, const void *K, const void *cost, void *Xc, void *Uc, void *cost_alpha,
                            int32_t *winner, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !x0 || !X || !U || !k || !K || !cost || !Xc || !Uc || !cost_alpha || !winner) return ILQR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_rollout(h, h->n_alpha_eff, h->alphas, phi, x0, X, U, k, K, Xc, Uc, cost_alpha, nullptr, nullptr, cost, st);
    if (rc) return rc;
    const int bs = 128;
    if (h->p.dtype == ILQR_F64)
        winner_kernel<double><<<grid_for(h->p.B, bs), bs, 0, st>>>(h->p.B, h->n_alpha_eff, (const double *)cost_alpha,
                                                                     (const double *)cost, winner);
    else
        winner_kernel<float><<<grid_for(h->p.B, bs), bs, 0, st>>>(h->p.B, h->n_alpha_eff, (const float *)cost_alpha,
                                                                    (const float *)cost, winner);
    ILQR_CHECK_LAUNCH(h);
    return ILQR_OK;
}

int ilqr_solve(ilqr_handle_t hh, const void *phi, const void *x0, void *X, void *U, void *K, void *k, void *cost,
               int32_t *iters, int32_t *status, void *ws, size_t ws_bytes, void *stream, int64_t *total_iters)
{
    Handle *h = (Handle *)hh;
    if (!h || !x0 || !X || !U || !K || !k || !cost || !iters || !status || !ws) return ILQR_E_INVALID;
    const ilqr_problem_t &p = h->p;
    const WsLayout L = ws_layout(p, h->n_alpha_eff);
    if (ws_bytes < L.total) return ILQR_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)ws;
    Control *ctl = (Control *)(w + L.ctl);
    void *A = w + L.A, *Bd = w + L.Bd, *Xc = w + L.Xc, *Uc = w + L.Uc, *ca = w + L.cost_alpha;
    int *winner = (int *)(w + L.winner), *active = (int *)(w + L.active), *defer = (int *)(w + L.defer);
    int *mark = (int *)(w + L.mark), *lists = (int *)(w + L.lists);
    int *wslot = h->lazy ? (int *)(w + L.wslot) : nullptr;      // only the lazy schedule stores candidates by list position
    RegArgs rg;
    std::memset(&rg, 0, sizeof rg);
    if (p.reg_factor > 1.0) {
        rg.mu = h->mu_user ? h->mu_user : (void *)(w + L.mu);
        rg.factor = p.reg_factor;
        rg.mu_min = p.reg_min;
        rg.mu_max = p.reg_max;
    }
    const int B = p.B, bsB = 128;
    int rc;
    // two-wave line search (see select_kernel): n1 eager step sizes, n2 deferred ones
    const int n1 = h->n_first, n2 = h->n_alpha_eff - h->n_first;
    const size_t wbytes = p.dtype == ILQR_F64 ? 8 : 4;
    const size_t xc_slab = wbytes * (size_t)(p.N + 1) * p.n * B, uc_slab = wbytes * (size_t)p.N * p.m * B;
    AlphaList al2;
    std::memset(&al2, 0, sizeof al2);
    for (int i = 0; i < n2; ++i) al2.a[i] = h->alphas.a[n1 + i];
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { h->last_cuda = (int)e_; return ILQR_E_CUDA; } } while (0)
    CU(cudaMemsetAsync(ctl, 0, ctl_bytes(p.maxiter), st));
    CU(cudaMemsetAsync(defer, 0, sizeof(int) * (size_t)B, st));
    CU(cudaMemsetAsync(mark, 0, sizeof(int) * (size_t)B, st));
    // initial rollout, alpha = 0, with the incoming X,K,k (iLQR_class.py:257-259) into candidate slab 0
    AlphaList a0;
    std::memset(&a0, 0, sizeof a0);
    if (h->profiling) { cudaStreamSynchronize(st); prof_collect(h); }
    prof_mark(h, ILQR_KC_OTHER, st);
    if ((rc = launch_rollout(h, 1, a0, phi, x0, X, U, k, K, Xc, Uc, ca, nullptr, nullptr, nullptr, st))) return rc;
    prof_mark(h, ILQR_KC_INIT_ROLLOUT, st);
    if (p.dtype == ILQR_F64)
        init_kernel<double><<<grid_for(B, bsB), bsB, 0, st>>>(B, (const double *)ca, (double *)cost, winner, active,
                                                                iters, status, p.maxiter, ctl, (double *)h->tr_cost,
                                                                (double *)rg.mu, p.reg_init);
    else
        init_kernel<float><<<grid_for(B, bsB), bsB, 0, st>>>(B, (const float *)ca, (float *)cost, winner, active, iters,
                                                               status, p.maxiter, ctl, (float *)h->tr_cost, (float *)rg.mu,
                                                               (float)p.reg_init);
    ILQR_CHECK_LAUNCH(h);
    // iterations are enqueued in blocks of CHK; the active count after each block is copied to pinned
    // memory and inspected one block later, so the device never idles waiting for the host.
    const char *chk_env = getenv("ILQR_CHECK_EVERY");
    const int CHK = chk_env ? (atoi(chk_env) > 0 ? atoi(chk_env) : 8) : 8;
    int pending = -1;   // event slot holding the count after the previous block
    int it = 0;
    bool stop = false;
    while (it < p.maxiter && !stop) {
        const int end = (it + CHK < p.maxiter) ? it + CHK : p.maxiter;
        for (; it < end; ++it) {
            const unsigned int *g = &ctl->n_active[it];
            const unsigned int *gprev = it > 0 ? &ctl->n_active[it - 1] : g;
            prof_mark(h, ILQR_KC_OTHER, st);
            const bool ltv = p.model == ILQR_LTV;      // commit only: A_t, B_t are generated inside the LTV kernels
            if ((rc = launch_commit_linearize(h, phi, X, U, A, Bd, Xc, Uc, winner, it > 0 ? wslot : nullptr, active,
                                              ltv ? 0 : 1, g, gprev, st))) return rc;
            prof_mark(h, ILQR_KC_LINEARIZE, st);
            if ((rc = ltv ? launch_backward_ltv(h, phi, X, U, K, k, active, g, st, rg.mu)
                          : launch_backward(h, X, U, A, Bd, K, k, active, g, st, rg.mu))) return rc;
            prof_mark(h, ILQR_KC_BACKWARD, st);
            if (h->lazy) {
                // lazy line search: wave v rolls out step sizes [wave_lo[v], wave_lo[v+1]) for the trajectories
                // that accepted none so far (wave 0: every active one); later waves return at once while
                // their list is empty
                unsigned int *wcnt = &ctl->n_active[3 * (p.maxiter + 2)] + (size_t)it * ILQR_MAX_WAVES;
                for (int v = 0; v < h->n_waves; ++v) {
                    const int lo = h->wave_lo[v], hi = h->wave_lo[v + 1];
                    AlphaList alv;
                    std::memset(&alv, 0, sizeof alv);
                    for (int i = lo; i < hi; ++i) alv.a[i - lo] = h->alphas.a[i];
                    const int *lin = v ? lists + (size_t)((v - 1) & 1) * B : nullptr;
                    int *lout = lists + (size_t)(v & 1) * B;
                    const unsigned int *cin = v ? wcnt + v - 1 : nullptr;
                    if ((rc = launch_rollout(h, hi - lo, alv, phi, x0, X, U, k, K, (char *)Xc + xc_slab * lo,
                                             (char *)Uc + uc_slab * lo, (char *)ca + wbytes * (size_t)lo * B,
                                             v ? nullptr : active, v ? cin : g, nullptr, st, nullptr, lin, cin))) return rc;
                    prof_mark(h, ILQR_KC_ROLLOUT, st);
                    if ((rc = launch_select_lazy(h, lo, hi, v, v == h->n_waves - 1, ca, cost, winner, active, iters, status,
                                                 it, ctl, lin, cin, lout, wcnt + v, wslot, rg, st))) return rc;
                }
                continue;
            }
            // line search, wave 1: the first n1 step sizes for every active trajectory (+ the deferred ones of
            // the trajectories on this iteration's speculation list)
            SpecArgs sp;
            std::memset(&sp, 0, sizeof sp);
            if (n2 > 0 && h->spec_cap > 0) {
                unsigned int *n_spec = &ctl->n_active[2 * (p.maxiter + 2)];
                sp.cap = h->spec_cap;
                sp.n2 = n2;
                sp.threshold = n1 - 4 > 1 ? n1 - 4 : 1;
                sp.list_cur = lists + (size_t)(it & 1) * B;
                sp.list_next = lists + (size_t)((it + 1) & 1) * B;
                sp.count_cur = n_spec + it;
                sp.count_next = n_spec + it + 1;
                sp.mark = mark;
            }
            if ((rc = launch_rollout(h, n1, h->alphas, phi, x0, X, U, k, K, Xc, Uc, ca, active, g, nullptr, st, &sp))) return rc;
            prof_mark(h, ILQR_KC_ROLLOUT, st);
            unsigned int *n2c = &ctl->n_active[p.maxiter + 2 + it];
            if ((rc = launch_select(h, 0, n1, 0, ca, cost, winner, active, n2 > 0 ? defer : nullptr, iters, status, it, ctl, n2c, sp, rg, st))) return rc;
            if (n2 > 0) {
                // wave 2: the remaining step sizes, only for trajectories that accepted none so far; both
                // launches return at once while the deferred counter of this iteration is zero
                if ((rc = launch_rollout(h, n2, al2, phi, x0, X, U, k, K, (char *)Xc + xc_slab * n1, (char *)Uc + uc_slab * n1,
                                         (char *)ca + wbytes * (size_t)n1 * B, defer, n2c, nullptr, st))) return rc;
                if ((rc = launch_select(h, n1, n1 + n2, 1, ca, cost, winner, active, defer, iters, status, it, ctl, n2c, sp, rg, st))) return rc;
            }
        }
        if (pending >= 0) {
            CU(cudaEventSynchronize(h->ev[pending]));
            if (h->h_flag[pending] == 0u) stop = true;
        }
        if (!stop && it < p.maxiter) {
            const int slot = pending < 0 ? 0 : 1 - pending;
            CU(cudaMemcpyAsync(&h->h_flag[slot], &ctl->n_active[it], sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(h->ev[slot], st));
            pending = slot;
        }
    }
    // commit the candidates accepted in the last executed iteration (no linearization)
    if ((rc = launch_commit_linearize(h, phi, X, U, A, Bd, Xc, Uc, winner, it > 0 ? wslot : nullptr, nullptr, 0, nullptr,
                                      nullptr, st))) return rc;
    if (total_iters) {
        unsigned long long tot = 0;
        CU(cudaMemcpyAsync(&tot, &ctl->total_iters, sizeof tot, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        *total_iters = (int64_t)tot;
        prof_collect(h);
    }
#undef CU
    return ILQR_OK;
}

int ilqr_set_linesearch_waves(ilqr_handle_t hh, int n_waves, const int32_t *sizes)
{
    Handle *h = (Handle *)hh;
    if (!h || n_waves < 0 || n_waves > ILQR_MAX_WAVES || (n_waves > 0 && !sizes)) return ILQR_E_INVALID;
    set_waves(h, n_waves, sizes);
    return ILQR_OK;
}

int ilqr_set_mu_buffer(ilqr_handle_t hh, void *mu)
{
    Handle *h = (Handle *)hh;
    if (!h) return ILQR_E_INVALID;
    h->mu_user = mu;
    return ILQR_OK;
}

int ilqr_set_profiling(ilqr_handle_t hh, int enable)
{
    Handle *h = (Handle *)hh;
    if (!h) return ILQR_E_INVALID;
    h->profiling = enable ? 1 : 0;
    h->prof_kind->clear();
    for (int i = 0; i < ILQR_N_KERNEL_CLASSES; ++i) { h->prof_ms[i] = 0.0; h->prof_cnt[i] = 0; }
    return ILQR_OK;
}

int ilqr_get_kernel_times(ilqr_handle_t hh, double *ms, int64_t *launches)
{
    Handle *h = (Handle *)hh;
    if (!h || !ms || !launches) return ILQR_E_INVALID;
    for (int i = 0; i < ILQR_N_KERNEL_CLASSES; ++i) { ms[i] = h->prof_ms[i]; launches[i] = h->prof_cnt[i]; }
    return ILQR_OK;
}

int ilqr_mpc_shift(ilqr_handle_t hh, void *U, void *u0, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !U) return ILQR_E_INVALID;
    const int threads = h->p.m * h->p.B, bs = 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->p.dtype == ILQR_F64)
        mpc_shift_kernel<double><<<grid_for(threads, bs), bs, 0, st>>>(h->p.N, h->p.m, h->p.B, (double *)U, (double *)u0);
    else
        mpc_shift_kernel<float><<<grid_for(threads, bs), bs, 0, st>>>(h->p.N, h->p.m, h->p.B, (float *)U, (float *)u0);
    ILQR_CHECK_LAUNCH(h);
    return ILQR_OK;
}

}  // extern "C"
