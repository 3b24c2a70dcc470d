// ilqr_kernels_rollout.cuh -- K3 forward rollout of every (step size, trajectory) pair and K4 selection / convergence /
// regularisation bookkeeping (eager two-wave and lazy multi-wave forms)
// Part of libilqr_b200.so; included by ilqr_b200.cu only (see the file map at its top).
#pragma once
#include "ilqr_systems.cuh"
#include "ilqr_kernels_common.cuh"

namespace ilqr {

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
                           T *__restrict__ Uw, bool valid)
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
    if (valid) {
#pragma unroll
        for (int i = 0; i < n; ++i) Xw[((size_t)t * n + i) * B + bw] = x[i];
#pragma unroll
        for (int j = 0; j < m; ++j) Uw[((size_t)t * m + j) * B + bw] = u[j];
    }
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
                               const unsigned int *__restrict__ list_count, const __grid_constant__ SparseArgs sa)
{
    constexpr int n = Sys::N, m = Sys::M;
    if (gate && *gate == 0u) return;
#if ILQR_TRIG_TABLE
    if constexpr (Sys::TRIG_TABLE) trig_table_init();
#endif
    // Warp w of the grid handles step size (w % n_alpha) of trajectory group (w / n_alpha): the warps that
    // re-read the same nominal trajectory and gains run next to each other, so at large batches those
    // reads come from L1/L2 instead of once per step size from HBM.
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t wg = gid >> 5, ngrp = ((size_t)B + 31) >> 5;
    // sparse iteration (SparseArgs; first wave of the lazy schedule only): the wave walks the compacted active
    // list, and once very few trajectories are left it tries every step size at once
    if (sparse_now(sa)) {
        list = sa.cur;
        list_count = sa.n_cur;
        if (sparse_all(sa)) n_alpha = sa.n_alpha_all;
    }
    int ai, b, bw;        // bw: column of the candidate slabs / cost_alpha this thread writes
    bool valid = true;
    if constexpr (n > 4) {
        // The n = 12 LTV model multiplies by 336 matrix constants per step.  Its instantiation keeps per-thread exits:
        // in convergent form ptxas hoists part of the matrices into registers and feeds them to the FP64 pipe through
        // R2UR moves (29 instead of 17 ms per rollout wave at B=32768, N=1000).
        if (wg < ngrp * n_alpha) {
            ai = (int)(wg % n_alpha);
            const unsigned int idx = (unsigned int)(wg / n_alpha) * 32u + (threadIdx.x & 31u);
            if (list) {                                                  // lazy wave: compacted trajectory list;
                if (idx >= min(*list_count, (unsigned int)B)) return;    // results stored at the list position
                b = list[idx];
            } else {
                if (idx >= (unsigned int)B) return;
                b = (int)idx;
            }
            bw = (int)idx;
        } else {                                                         // speculative extra threads
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
    } else {
    // No per-thread early return: a lane without work follows its warp with its stores masked, so the time loop
    // below is convergent code (whole warps without work leave together); see commit_linearize_point.
    if (wg < ngrp * n_alpha) {
        ai = (int)(wg % n_alpha);
        const unsigned int idx = (unsigned int)(wg / n_alpha) * 32u + (threadIdx.x & 31u);
        if (list) {                                                      // lazy wave: compacted trajectory list;
            valid = idx < min(*list_count, (unsigned int)B);             // results stored at the list position
            b = valid ? list[idx] : 0;
        } else {
            valid = idx < (unsigned int)B;
            b = valid ? (int)idx : 0;
        }
        bw = (int)idx;
    } else {                                                             // speculative extra threads
        const size_t e = gid - ngrp * n_alpha * 32;
        valid = !(list || sp.cap == 0 || e >= (size_t)sp.cap * sp.n2);
        const int q = valid ? (int)(e % sp.cap) : 0;
        if (valid) valid = (unsigned int)q < min(*sp.count_cur, (unsigned int)sp.cap);
        b = valid ? sp.list_cur[q] : 0;
        bw = b;
        ai = valid ? n_alpha + (int)(e / sp.cap) : 0;
    }
    // (an idle lane keeps reading its own trajectory: redirecting it to trajectory 0 would save its loads but made
    // ptxas schedule the whole time loop 17 % slower -- measured)
    if (valid && active && !active[b]) valid = false;
    if (!__any_sync(0xffffffffu, valid)) return;
    }
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
    // the next step's nominal/gains are loaded into a second buffer while the current step computes.  (Unrolling
    // the loop by two over a ping-pong pair of buffers, ILQR_UNROLL2=1, saved the register copies and was the faster
    // form of the 218-register kernel; at 126 registers the single-step body is 6 % faster at B=4096.)
#ifndef ILQR_UNROLL2
#define ILQR_UNROLL2 0
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
        rollout_step<INTEG>(sys, qc, in0, alpha, t, bw, B, ph, x, cost, Xw, Uw, valid);
        if (t + 1 >= N) break;
        if (t + 2 < N) fwd_load(in0, t + 2, b, B, X_old, U_old, k, K);
        rollout_step<INTEG>(sys, qc, in1, alpha, t + 1, bw, B, ph, x, cost, Xw, Uw, valid);
#if ILQR_REJECT
        if (can_reject && !(cost <= c_ref)) {
            if (valid) cost_alpha[(size_t)ai * B + bw] = cost;            // already > cost to beat (or NaN): rejected
            return;
        }
#endif
    }
#else
    FwdIn<T, n, m> in0, in1;
    fwd_load(in0, 0, b, B, X_old, U_old, k, K);
    for (int t = 0; t < N; ++t) {
        if (t + 1 < N) fwd_load(in1, t + 1, b, B, X_old, U_old, k, K);
        rollout_step<INTEG>(sys, qc, in0, alpha, t, bw, B, ph, x, cost, Xw, Uw, valid);
        in0 = in1;
#if ILQR_REJECT
        if (can_reject && !(cost <= c_ref)) {
            if (valid) cost_alpha[(size_t)ai * B + bw] = cost;
            return;
        }
#endif
    }
#endif
    }
    if (!valid) return;
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
                    if (sp.cap > 0) {
                        const int sum = w + sp.hist[b];
                        sp.hist[b] = w;
                        if (still && (w >= sp.threshold || sum >= sp.threshold + 3)) {   // small steps: list it for next time
                            const unsigned int pos = atomicAdd(sp.count_next, 1u);
                            if (pos < (unsigned int)sp.cap) { sp.list_next[pos] = b; sp.mark[b] = it + 2; }
                        } else if (still && sum >= sp.threshold + 2) {                   // tier 2: if room is left
                            const unsigned int pos = atomicAdd(sp.count2_next, 1u);
                            if (pos < (unsigned int)sp.cap) sp.list2_next[pos] = b;
                        }
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
    // the last block of the first wave's select appends tier 2 behind tier 1 while capacity lasts (a second wave, if
    // one runs, lists only trajectories that accepted a deferred step size: tier 1, appended behind these)
    if (wave == 0 && sp.cap > 0) {
        __shared__ bool last_block;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last_block = atomicAdd(sp.ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (last_block) {
            __threadfence();
            const unsigned int c1 = min(*(volatile unsigned int *)sp.count_next, (unsigned int)sp.cap);
            const unsigned int c2 = min(*(volatile unsigned int *)sp.count2_next, (unsigned int)sp.cap);
            const unsigned int nc = min(c2, (unsigned int)sp.cap - c1);
            for (unsigned int j = threadIdx.x; j < nc; j += blockDim.x) {
                const int bj = ((volatile int *)sp.list2_next)[j];
                sp.list_next[c1 + j] = bj;
                sp.mark[bj] = it + 2;
            }
            __syncthreads();
            if (threadIdx.x == 0 && nc) atomicAdd(sp.count_next, nc);
        }
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
                                   int *__restrict__ tr_alpha, T *__restrict__ tr_cost, const __grid_constant__ RegArgs rg,
                                   const __grid_constant__ SparseArgs sa)
{
    if (ctl->n_active[it] == 0u) return;
    if (wave > 0 && *cnt_in == 0u) return;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int b = gid;
    bool mine = false;
    if (wave == 0 && sparse_now(sa)) {
        // sparse iteration: the first wave ran over the active list (and, with very few left, tried every step size)
        if ((unsigned int)gid < *sa.n_cur) {
            b = sa.cur[gid];
            mine = active[b] != 0;
        }
        if (sparse_all(sa)) {
            a_hi = sa.n_alpha_all;
            last = 1;
        }
    } else if (wave == 0) {
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
    // trajectories entering the next iteration: count them and, for the sparse mode, list them (the counter's
    // old value is the warp's position in the list)
    const unsigned ms = __ballot_sync(full, still);
    if (ms) {
        const int leader = __ffs(ms) - 1;
        unsigned int base = 0;
        if ((int)lane == leader) base = atomicAdd(&ctl->n_active[it + 1], (unsigned int)__popc(ms));
        base = __shfl_sync(full, base, leader);
        if (still && sa.next) {
            const unsigned int at = base + __popc(ms & ((1u << lane) - 1u));
            sa.next[at] = b;
            sa.pos[b] = (int)at;
        }
    }
    const unsigned nr = __popc(__ballot_sync(full, mine && wave == 0));
    if (lane == 0 && nr) atomicAdd(&ctl->total_iters, (unsigned long long)nr);
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

}  // namespace ilqr
