// ilqr_kernels_linearize.cuh -- System.f_fcn for a batch (step_kernel), K1 commit + linearization, materialised cost
// expansion and the MPC warm-start shift: the kernels with one thread per (t,b) or per b
// Part of libilqr_b200.so; included by ilqr_b200.cu only (see the file map at its top).
#pragma once
#include "ilqr_systems.cuh"
#include "ilqr_kernels_common.cuh"

namespace ilqr {

template <class Sys, int INTEG, typename T>
__global__ void step_kernel(const __grid_constant__ Sys sys, T dt, int B, int t, const T *__restrict__ phi,
                            const T *__restrict__ x, const T *__restrict__ u, T *__restrict__ xn)
{
    constexpr int n = Sys::N, m = Sys::M;
#if ILQR_TRIG_TABLE
    if constexpr (Sys::TRIG_TABLE) trig_table_init();
#endif
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

// K1.  One thread per (t,b), t in [0,N].  If the trajectory ran the previous iteration (iters[b] == it) and
// accepted a candidate (winner[b] >= 0) the thread first copies it (Xc/Uc slab winner[b]) into the nominal X/U; if
// the trajectory is active it then writes the discrete Jacobians about that nominal point.
// Written as CONVERGENT code: a lane without work (out of range, inactive trajectory, t = N) follows its warp with
// its loads redirected and its stores masked instead of returning early, and warps leave only as a whole.  ptxas
// keeps kernel and system constants in uniform registers only in code it can prove convergent; after a per-thread
// return it falls back to per-thread registers for them (40-70 registers more in these kernels, and DFMAs with
// three register operands, which issue at 3 instead of 2 cycles on sm_100a).
template <class Sys, int INTEG, typename T>
ILQR_DEV void commit_linearize_point(const Sys &sys, T dt, int N, int B, bool valid, int t, int b, const T *__restrict__ phi,
                                     T *__restrict__ X, T *__restrict__ U, T *__restrict__ A, T *__restrict__ Bd,
                                     const T *__restrict__ Xc, const T *__restrict__ Uc, const int *__restrict__ winner,
                                     const int *__restrict__ wslot, const int *__restrict__ active,
                                     const int *__restrict__ iters, int it, int do_linearize,
                                     const int *__restrict__ pos, int ab_blocked)
{
    constexpr int n = Sys::N, m = Sys::M;
    const unsigned full = 0xffffffffu;
    if (!valid) { t = 0; b = 0; }
    int w = (valid && winner) ? winner[b] : -1;
    if (iters && iters[b] != it) w = -1;                 // nothing pending: committed earlier, or never ran
    const bool act = valid && do_linearize && (active ? active[b] != 0 : true) && t < N;
    if (!__any_sync(full, w >= 0 || act)) return;
    T x[n], u[m];
    // lazy line search: candidates of the later waves are stored at the trajectory's list position
    const int col = (w >= 0 && wslot) ? wslot[b] : b;
    const T *xs = w >= 0 ? Xc + (size_t)w * (N + 1) * n * B : X, *us = w >= 0 ? Uc + (size_t)w * N * m * B : U;
    const int tu = t < N ? t : N - 1;                    // t = N has no control: read a valid one, never used
    const bool need = w >= 0 || act;                     // a lane with nothing to commit or linearize reads nothing
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = need ? xs[((size_t)t * n + i) * B + col] : T(0);
#pragma unroll
    for (int j = 0; j < m; ++j) u[j] = need ? us[((size_t)tu * m + j) * B + col] : T(0);
    if (w >= 0) {
#pragma unroll
        for (int i = 0; i < n; ++i) X[((size_t)t * n + i) * B + b] = x[i];
        if (t < N) {
#pragma unroll
            for (int j = 0; j < m; ++j) U[((size_t)t * m + j) * B + b] = u[j];
        }
    }
    if (!__any_sync(full, act)) return;
    T Aj[n][n], Bj[n][m];
    step_jac<INTEG>(sys, dt, x, u, Aj, Bj, sys.time_scalar(t, phi ? phi[b] : T(0)));
    if (!act) return;
    const int c = pos ? pos[b] : b;          // sparse iteration: compact column = position in the active list
    if (ab_blocked) {                        // workspace layout of ilqr_solve (ab_off)
        constexpr int R = n * n + n * m;
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) A[ab_off(R, t, i * n + j, c, B)] = Aj[i][j];
#pragma unroll
            for (int j = 0; j < m; ++j) A[ab_off(R, t, n * n + i * m + j, c, B)] = Bj[i][j];
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < n; ++j) A[(((size_t)t * n + i) * n + j) * B + c] = Aj[i][j];
#pragma unroll
        for (int j = 0; j < m; ++j) Bd[(((size_t)t * n + i) * m + j) * B + c] = Bj[i][j];
    }
}

template <class Sys, int INTEG, typename T>
__global__ void __launch_bounds__(128, Sys::N <= 4 ? 4 : 1) commit_linearize_kernel(const __grid_constant__ Sys sys, T dt, int N, int B,
                                        const T *__restrict__ phi, T *__restrict__ X, T *__restrict__ U,
                                        T *__restrict__ A, T *__restrict__ Bd, const T *__restrict__ Xc,
                                        const T *__restrict__ Uc, const int *__restrict__ winner,
                                        const int *__restrict__ wslot, const int *__restrict__ active,
                                        const int *__restrict__ iters, int it, int do_linearize,
                                        const unsigned int *__restrict__ gate0, const unsigned int *__restrict__ gate1,
                                        const __grid_constant__ SparseArgs sa, int ab_blocked, int sparse_only)
{
    if (gate0 && *gate0 == 0u && *gate1 == 0u) return;   // nobody active now or in the previous iteration
    if (sparse_only && !sparse_now(sa)) return;          // dense iterations: the fused kernel commits and linearizes
#if ILQR_TRIG_TABLE
    if constexpr (Sys::TRIG_TABLE) { if (do_linearize) trig_table_init(); }
#endif
    const int *pos = sparse_now(sa) ? sa.pos : nullptr;  // where K2 will look for A_t, B_t in this iteration
    if (sparse_prev(sa)) {
        // few trajectories ran the previous iteration: the first blocks stride over (t, list entry) pairs, the
        // rest of the grid (sized for the whole batch) leaves at once
        const unsigned int cnt = *sa.n_prev;
        const unsigned int nblk = min(gridDim.x, 148u * 8u);
        if (blockIdx.x >= nblk) return;
        const size_t total = (size_t)(N + 1) * cnt, stride = (size_t)nblk * blockDim.x;
        // the loop bound is the warp's first item, so a warp runs its last round together
        for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total; base += stride) {
            const size_t item = base + (threadIdx.x & 31u);
            const bool valid = item < total;
            commit_linearize_point<Sys, INTEG, T>(sys, dt, N, B, valid, valid ? (int)(item / cnt) : 0,
                                                  valid ? sa.prev[item % cnt] : 0, phi, X, U, A, Bd, Xc, Uc, winner,
                                                  wslot, active, iters, it, do_linearize, pos, ab_blocked);
        }
        return;
    }
    // one item per thread; a grid smaller than the item count (sparse_only launches) strides over them, warps staying whole
    const size_t total = (size_t)(N + 1) * B, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total; base += stride) {
        const size_t gid = base + (threadIdx.x & 31u);
        const bool valid = gid < total;
        commit_linearize_point<Sys, INTEG, T>(sys, dt, N, B, valid, valid ? (int)(gid / B) : 0, valid ? (int)(gid % B) : 0, phi,
                                              X, U, A, Bd, Xc, Uc, winner, wslot, active, iters, it, do_linearize, pos,
                                              ab_blocked);
    }
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

}  // namespace ilqr
