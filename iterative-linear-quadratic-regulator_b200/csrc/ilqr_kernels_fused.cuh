// ilqr_kernels_fused.cuh -- K1 and K2 as ONE warp-specialised kernel: producer warps commit the accepted line-search
// candidate and compute the analytic A_t, B_t of the steps ahead (K1's work, iLQR_class.py:318-331), the consumer warp
// runs the Riccati recursion behind them (K2's work, iLQR_class.py:79-161).  A_t, B_t travel through a shared-memory
// ring guarded by mbarriers and never touch HBM: of the 600 algorithmic bytes per trajectory-timestep of the three-kernel
// iteration, the 320 that existed only because linearization and recursion were two kernels disappear, together with
// one launch per iteration.
// Part of libilqr_b200.so; included by ilqr_b200.cu only (see the file map at its top).
#pragma once
#include "ilqr_systems.cuh"
#include "ilqr_kernels_common.cuh"
#include "ilqr_kernels_backward.cuh"

namespace ilqr {

// (mbarrier helpers: ilqr_kernels_common.cuh)

// Thread-per-trajectory form.  One block = one group of 32 trajectories
// (lane = trajectory) = 1 consumer warp + NP producer warps.  Producer p owns the scan steps i = N-1-t with
// i mod NP == p; ring stage i mod S (S a multiple of NP, so a producer always meets the same stages).  A lane only ever
// reads what the same lane of a producer wrote, so a stage is [rows][32 lanes]: conflict-free in both directions.
//   producer:  x_t,u_t from the accepted candidate slab (or the nominal; cp.async, one step ahead) -> commit into X,U ->
//              wait empty[s] -> step_jac -> ring -> arrive full[s]
//   consumer:  wait full[s] -> ring -> registers -> arrive empty[s] -> riccati_step -> coalesced K_t, k_t stores
// MINB = resident blocks per SM the register allocation is capped for: 1 (no cap: small batches, one block per SM or
// fewer, latency bound) or 5 (large batches: 15 warps per SM at 128 registers keep the FP64 pipe busier than 12 at 164;
// measured 5.37 vs 5.50 vs 6.16 ms per pass at B=131072 for caps of 5 / 4 / none).
template <class Sys, class Cost, int INTEG, typename T, int NP, int S, int MINB>
__global__ void __launch_bounds__(32 * (NP + 1), MINB)
fused_backward_kernel(const __grid_constant__ Sys sys, const __grid_constant__ Cost qc, int N, int B, const T *__restrict__ phi,
                      T *__restrict__ X, T *__restrict__ U, const T *__restrict__ Xc, const T *__restrict__ Uc,
                      const int *__restrict__ winner, const int *__restrict__ wslot, const int *__restrict__ active,
                      const int *__restrict__ iters, int it, const unsigned int *__restrict__ gate0,
                      const unsigned int *__restrict__ gate1, T *__restrict__ K, T *__restrict__ k,
                      const T *__restrict__ mu, const __grid_constant__ SparseArgs sa)
{
    constexpr int n = Sys::N, m = Sys::M, L = n * n + n * m + n + m;
    static_assert(S % NP == 0, "ring stages must be a multiple of the producer count");
    __shared__ __align__(16) T ring[S][L][32];
    __shared__ __align__(16) T pre[NP][2][n + m][32];       // producers' input staging (cp.async)
    __shared__ __align__(8) unsigned long long full[S], empty[S];
    if (gate0 && *gate0 == 0u && *gate1 == 0u) return;       // nobody active now or in the previous iteration
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // Sparse iteration (SparseArgs; lazy schedule, few trajectories left): the groups walk the compacted active list.
    // The pending commits of such an iteration -- which include trajectories that have just finished and are on no list
    // any more -- were made by K1 in its commit-only form just before this launch.
    const bool sparse = sparse_now(sa);
    const int n_items = sparse ? (int)*sa.n_cur : B;
    if (blockIdx.x * 32 >= n_items) return;
    const int item = blockIdx.x * 32 + lane;
    const bool valid = item < n_items;
    const int b = sparse ? sa.cur[valid ? item : n_items - 1] : (valid ? item : B - 1);   // spare lanes: a copy, never stored
    int w = (valid && winner && !sparse) ? winner[b] : -1;
    if (iters && iters[b] != it) w = -1;                     // nothing pending: committed earlier, or never ran
    const bool act = valid && (active ? active[b] != 0 : true);
    const unsigned full_mask = 0xffffffffu;
    const bool any_act = __any_sync(full_mask, act), any_commit = __any_sync(full_mask, w >= 0);
    if (!any_act && !any_commit) return;                     // every warp of the block sees the same 32 trajectories
    // source of x_t, u_t: the accepted candidate (lazy schedule: stored at the trajectory's list position) or the nominal
    const int col = (w >= 0 && wslot) ? wslot[b] : b;
    const T *xs = w >= 0 ? Xc + (size_t)w * (N + 1) * n * B + col : X + col;
    const T *us = w >= 0 ? Uc + (size_t)w * N * m * B + col : U + col;
    const size_t sB = (size_t)B;
    if (!any_act) {
        // the group only has candidates to commit (its trajectories finished in the previous iteration): plain copy
        for (int t = wid; t <= N; t += NP + 1) {
            if (w >= 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) X[((size_t)t * n + i) * sB + b] = xs[((size_t)t * n + i) * sB];
                if (t < N) {
#pragma unroll
                    for (int j = 0; j < m; ++j) U[((size_t)t * m + j) * sB + b] = us[((size_t)t * m + j) * sB];
                }
            }
        }
        return;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 32); mbar_init(&empty[s], 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#if ILQR_TRIG_TABLE
    if constexpr (Sys::TRIG_TABLE) trig_table_init();
#endif
    __syncthreads();

    if (wid == 0) {
        // ---------------- consumer: the reverse scan, value function in registers ----------------
        const T mu_b = mu ? mu[b] : T(0);                    // regularisation (RegArgs), 0 in the reference
        T Vx[n], Vxx[n][n];
        {
            T xN[n];
#pragma unroll
            for (int i = 0; i < n; ++i) xN[i] = xs[((size_t)N * n + i) * sB];
            if (w >= 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) X[((size_t)N * n + i) * sB + b] = xN[i];
            }
            if constexpr (Cost::QUADRATIC) {
                qc.terminal_grad(xN, Vx);                    // iLQR_class.py:136-138
#pragma unroll
                for (int i = 0; i < n; ++i)
#pragma unroll
                    for (int j = 0; j < n; ++j) Vxx[i][j] = qc.Qfs[i][j];
            } else {
                qc.terminal_expand(xN, Vx, Vxx);
            }
        }
        T *Kp = K + (size_t)(N - 1) * m * n * sB + b, *kp = k + (size_t)(N - 1) * m * sB + b;
        for (int i = 0; i < N; ++i) {
            const int s = i % S;
            mbar_wait(&full[s], (unsigned)(i / S) & 1u);
            BwdIn<T, n, m> cur;
            {
                const T *st = &ring[s][0][lane];
                int row = 0;
#pragma unroll
                for (int r = 0; r < n; ++r)
#pragma unroll
                    for (int c = 0; c < n; ++c, ++row) cur.A[r][c] = st[row * 32];
#pragma unroll
                for (int r = 0; r < n; ++r)
#pragma unroll
                    for (int c = 0; c < m; ++c, ++row) cur.Bd[r][c] = st[row * 32];
#pragma unroll
                for (int r = 0; r < n; ++r, ++row) cur.x[r] = st[row * 32];
#pragma unroll
                for (int c = 0; c < m; ++c, ++row) cur.u[c] = st[row * 32];
            }
            mbar_arrive(&empty[s]);                          // the stage is in registers: hand it back at once
            T Kt[m][n], kt[m];
            riccati_step<Cost, T, n, m>(qc, cur, mu_b, Vx, Vxx, Kt, kt);
            if (act) {
#pragma unroll
                for (int j = 0; j < m; ++j) {
#pragma unroll
                    for (int c = 0; c < n; ++c) Kp[((size_t)j * n + c) * sB] = Kt[j][c];
                    kp[(size_t)j * sB] = kt[j];
                }
            }
            Kp -= (size_t)m * n * sB;
            kp -= (size_t)m * sB;
        }
        return;
    }

    // ---------------- producers: commit + linearization, NP steps apart ----------------
    // The inputs of a producer's NEXT step travel by cp.async into a two-deep staging row of its own (no registers held
    // across step_jac); the ring stage is claimed BEFORE the arithmetic, so that x_t, u_t go into it at once and A_t, B_t
    // as they are finished (with S stages a claim only waits when the producers are a full ring ahead anyway).
    const int p = wid - 1;
    const T ph = phi ? phi[b] : T(0);
    const T *px = xs + (size_t)(N - 1 - p) * n * sB, *pu = us + (size_t)(N - 1 - p) * m * sB;
    T *qx = X + (size_t)(N - 1 - p) * n * sB + b, *qu = U + (size_t)(N - 1 - p) * m * sB + b;
    const size_t dx = (size_t)NP * n * sB, du = (size_t)NP * m * sB;
    auto prefetch = [&](int buf) {
        T *dst = &pre[p][buf][0][lane];
#pragma unroll
        for (int r = 0; r < n; ++r) cp_async<sizeof(T)>(dst + r * 32, px + (size_t)r * sB);
#pragma unroll
        for (int c = 0; c < m; ++c) cp_async<sizeof(T)>(dst + (n + c) * 32, pu + (size_t)c * sB);
        px -= dx;
        pu -= du;
    };
    if (p < N) prefetch(0);
    cp_async_commit();
    int buf = 0;
    for (int i = p; i < N; i += NP) {
        const int t = N - 1 - i, s = i % S;
        if (i + NP < N) prefetch(buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();                                  // this step's inputs have landed (own copies only)
        T x[n], u[m];
        {
            const T *src = &pre[p][buf][0][lane];
#pragma unroll
            for (int r = 0; r < n; ++r) x[r] = src[r * 32];
#pragma unroll
            for (int c = 0; c < m; ++c) u[c] = src[(n + c) * 32];
        }
        buf ^= 1;
        if (w >= 0) {                                        // commit the accepted candidate into the nominal
#pragma unroll
            for (int r = 0; r < n; ++r) qx[(size_t)r * sB] = x[r];
#pragma unroll
            for (int c = 0; c < m; ++c) qu[(size_t)c * sB] = u[c];
        }
        qx -= dx;
        qu -= du;
        mbar_wait(&empty[s], ((unsigned)(i / S) & 1u) ^ 1u);
        T *st = &ring[s][0][lane];
#pragma unroll
        for (int r = 0; r < n; ++r) st[(n * n + n * m + r) * 32] = x[r];
#pragma unroll
        for (int c = 0; c < m; ++c) st[(n * n + n * m + n + c) * 32] = u[c];
        T Aj[n][n], Bj[n][m];
        step_jac<INTEG>(sys, qc.dt, x, u, Aj, Bj, sys.time_scalar(t, ph));
#pragma unroll
        for (int r = 0; r < n; ++r) {
#pragma unroll
            for (int c = 0; c < n; ++c) st[(r * n + c) * 32] = Aj[r][c];
#pragma unroll
            for (int c = 0; c < m; ++c) st[(n * n + r * m + c) * 32] = Bj[r][c];
        }
        mbar_arrive(&full[s]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Small batches (one block per SM or fewer): the kernel above runs at the speed of its single consumer warp, which is
// ISSUE bound -- 282 FP64 instructions per step from one warp, >= 2 issue cycles each, ~1 070 cycles per step while the
// other three sub-partitions of the SM only linearize.  Here the recursion is split over TWO consumer warps: warp H owns
// the columns [H n/2, (H+1) n/2) of V_xx, Q_xx and K.  Per step each half computes its columns of T1 = f_x' V_xx and
// T2 = f_u' V_xx, the halves swap them (plus their entries of the new V_x) through shared memory at one named barrier,
// and each finishes its columns of Q_xx, K and V_xx; Q_ux, Q_uu, Q_u (m rows) are computed by both.  Every output element
// is produced by the same operation sequence as in riccati_step, so K, k stay BIT-IDENTICAL to the one-consumer kernel
// and to backward_kernel (test_fused_linearize_backward_is_bit_identical).  NP = 4 producer warps feed the ring.  Warp
// roles follow the hardware's warp -> sub-partition map (warp w runs on sub-partition w mod 4, scripts/micro/
// warp_smsp_map.cu): warps 2 and 3 are the consumers, each alone on its sub-partition, warps 0, 4 and 1, 5 the producers.
// Measured at B=4096, N=500 (profiles/r02_exp_fused_split.log): one consumer + 2 producers 0.273 ms; split + 2 producers
// 0.262 (producer bound: 2 060 cycles per linearization from a lone warp); split + 4 producers 0.226 (with Euler steps,
// whose linearization is almost free, 0.208: what remains is the halves' per-step latency -- ring hand-off, exchange
// barrier, reciprocal chain); six producers, two of them on the consumers' sub-partitions: 0.273-
// 0.281 (a consumer half that shares its sub-partition is as slow as the unsplit consumer); both halves on one
// sub-partition and six producers on the other three: 0.29; three producers per sub-partition: 0.224.
// ------------------------------------------------------------------------------------------------------------------
ILQR_DEV void bar_sync_consumers() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

// One step of the reverse scan for the columns [H h, (H+1) h), h = n/2 (see riccati_step, whose expressions these are).
// Vx: full V_x on entry (own entries valid, the partner's are refreshed from the exchange); Vc: own columns of V_xx.
// xw / xr: this half's / the partner's exchange rows of the step, [row][32 lanes].
template <class Cost, typename T, int n, int m, int H>
ILQR_DEV void riccati_half(const Cost &qc, const BwdIn<T, n, m> &cur, T mu_b, T *Vx, T (*Vc)[n / 2], T (*Kc)[n / 2], T *kt,
                           T *xw, const T *xr, int lane)
{
    static_assert(Cost::QUADRATIC && n % 2 == 0, "split consumer: quadratic cost, even state dimension");
    constexpr int h = n / 2, off = H * h, offo = (1 - H) * h;
    // T1 = f_x' V_xx, T2 = f_u' V_xx: own columns                                                     (:102-104)
    T T1[n][n], T2[m][n];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int jo = 0; jo < h; ++jo) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += cur.A[l][i] * Vc[l][jo];
            T1[i][off + jo] = s;
        }
#pragma unroll
    for (int i = 0; i < m; ++i)
#pragma unroll
        for (int jo = 0; jo < h; ++jo) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += cur.Bd[l][i] * Vc[l][jo];
            T2[i][off + jo] = s;
        }
    // swap: own columns of T1, T2 and own entries of V_x out, the partner's in
    {
        int row = 0;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int jo = 0; jo < h; ++jo, ++row) xw[row * 32 + lane] = T1[i][off + jo];
#pragma unroll
        for (int i = 0; i < m; ++i)
#pragma unroll
            for (int jo = 0; jo < h; ++jo, ++row) xw[row * 32 + lane] = T2[i][off + jo];
#pragma unroll
        for (int io = 0; io < h; ++io, ++row) xw[row * 32 + lane] = Vx[off + io];
    }
    bar_sync_consumers();
    {
        int row = 0;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int jo = 0; jo < h; ++jo, ++row) T1[i][offo + jo] = xr[row * 32 + lane];
#pragma unroll
        for (int i = 0; i < m; ++i)
#pragma unroll
            for (int jo = 0; jo < h; ++jo, ++row) T2[i][offo + jo] = xr[row * 32 + lane];
#pragma unroll
        for (int io = 0; io < h; ++io, ++row) Vx[offo + io] = xr[row * 32 + lane];
    }
    // Q_x (own entries), Q_u                                                                            (:100-101)
    T lx[n], lu[m];
    qc.grad(cur.x, cur.u, lx, lu);
    T Qx[h], Qu[m];
#pragma unroll
    for (int io = 0; io < h; ++io) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < n; ++l) s += cur.A[l][off + io] * Vx[l];
        Qx[io] = lx[off + io] + s;
    }
#pragma unroll
    for (int j = 0; j < m; ++j) {
        T s = T(0);
#pragma unroll
        for (int l = 0; l < n; ++l) s += cur.Bd[l][j] * Vx[l];
        Qu[j] = lu[j] + s;
    }
    // Q_xx (own columns), Q_ux, Q_uu
    T Qxx[n][h], Qux[m][n], Quu[m][m];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int jo = 0; jo < h; ++jo) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += T1[i][l] * cur.A[l][off + jo];
            Qxx[i][jo] = qc.Qs[i][off + jo] * qc.dt + s;
        }
#pragma unroll
    for (int i = 0; i < m; ++i) {
#pragma unroll
        for (int j = 0; j < n; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += T2[i][l] * cur.A[l][j];
            Qux[i][j] = s;                                                // l_ux = 0 for the quadratic cost
        }
#pragma unroll
        for (int j = 0; j < m; ++j) {
            T s = T(0);
#pragma unroll
            for (int l = 0; l < n; ++l) s += T2[i][l] * cur.Bd[l][j];
            Quu[i][j] = qc.Rs[i][j] * qc.dt + s;
            if (i == j) Quu[i][j] += mu_b;
        }
    }
    // K (own columns) and k                                                                             (:109-110)
    if (m == 1) {
        const T r = -rcp_t(Quu[0][0]);
#pragma unroll
        for (int jo = 0; jo < h; ++jo) Kc[0][jo] = Qux[0][off + jo] * r;
        kt[0] = Qu[0] * r;
    } else {
        T rhs[m][h + 1];
#pragma unroll
        for (int i = 0; i < m; ++i) {
#pragma unroll
            for (int jo = 0; jo < h; ++jo) rhs[i][jo] = Qux[i][off + jo];
            rhs[i][h] = Qu[i];
        }
        T Lm[m][m];
#pragma unroll
        for (int i = 0; i < m; ++i)
#pragma unroll
            for (int j = 0; j < m; ++j) Lm[i][j] = Quu[i][j];
        lu_solve_inplace<m, h + 1>(Lm, rhs);
#pragma unroll
        for (int i = 0; i < m; ++i) {
#pragma unroll
            for (int jo = 0; jo < h; ++jo) Kc[i][jo] = -rhs[i][jo];
            kt[i] = -rhs[i][h];
        }
    }
    // V_x (own entries), V_xx (own columns)                                                             (:113-114)
#pragma unroll
    for (int io = 0; io < h; ++io) {
        T s = T(0);
#pragma unroll
        for (int j = 0; j < m; ++j) s += Kc[j][io] * Qu[j];
        Vx[off + io] = Qx[io] + s;
    }
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int co = 0; co < h; ++co) {
            T s2 = T(0);
#pragma unroll
            for (int j = 0; j < m; ++j) s2 += Qux[j][i] * Kc[j][co];
            Vc[i][co] = Qxx[i][co] + s2;
        }
}

template <class Sys, class Cost, int NP, int S, typename T>
constexpr size_t fused_split_smem_bytes()
{
    constexpr int n = Sys::N, m = Sys::M, L = n * n + n * m + n + m, XR = n * (n / 2) + m * (n / 2) + n / 2;
    return sizeof(T) * 32 * (size_t)(S * L + NP * 2 * (n + m) + 2 * 2 * XR) + 2 * S * sizeof(unsigned long long);
}

template <class Sys, class Cost, int INTEG, typename T, int NP, int S>
__global__ void __launch_bounds__(32 * (NP + 2), 1)
fused_backward_split_kernel(const __grid_constant__ Sys sys, const __grid_constant__ Cost qc, int N, int B,
                            const T *__restrict__ phi, T *__restrict__ X, T *__restrict__ U, const T *__restrict__ Xc,
                            const T *__restrict__ Uc, const int *__restrict__ winner, const int *__restrict__ wslot,
                            const int *__restrict__ active, const int *__restrict__ iters, int it,
                            const unsigned int *__restrict__ gate0, const unsigned int *__restrict__ gate1,
                            T *__restrict__ K, T *__restrict__ k, const T *__restrict__ mu,
                            const __grid_constant__ SparseArgs sa)
{
    constexpr int n = Sys::N, m = Sys::M, h = n / 2, L = n * n + n * m + n + m, XR = n * h + m * h + h, NW = NP + 2;
    static_assert(S % NP == 0, "ring stages must be a multiple of the producer count");
    extern __shared__ __align__(16) unsigned char fsplit_raw[];
    T *ring = reinterpret_cast<T *>(fsplit_raw);                       // [S][L][32]
    T *pre = ring + (size_t)S * L * 32;                                // [NP][2][n + m][32]
    T *xch = pre + (size_t)NP * 2 * (n + m) * 32;                      // [2 buffers][2 halves][XR][32]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(xch + (size_t)2 * 2 * XR * 32), *empty = full + S;
    if (gate0 && *gate0 == 0u && *gate1 == 0u) return;       // nobody active now or in the previous iteration
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool sparse = sparse_now(sa);
    const int n_items = sparse ? (int)*sa.n_cur : B;
    if (blockIdx.x * 32 >= n_items) return;
    const int item = blockIdx.x * 32 + lane;
    const bool valid = item < n_items;
    const int b = sparse ? sa.cur[valid ? item : n_items - 1] : (valid ? item : B - 1);   // spare lanes: a copy, never stored
    int w = (valid && winner && !sparse) ? winner[b] : -1;
    if (iters && iters[b] != it) w = -1;                     // nothing pending: committed earlier, or never ran
    const bool act = valid && (active ? active[b] != 0 : true);
    const unsigned full_mask = 0xffffffffu;
    const bool any_act = __any_sync(full_mask, act), any_commit = __any_sync(full_mask, w >= 0);
    if (!any_act && !any_commit) return;                     // every warp of the block sees the same 32 trajectories
    const int col = (w >= 0 && wslot) ? wslot[b] : b;
    const T *xs = w >= 0 ? Xc + (size_t)w * (N + 1) * n * B + col : X + col;
    const T *us = w >= 0 ? Uc + (size_t)w * N * m * B + col : U + col;
    const size_t sB = (size_t)B;
    if (!any_act) {
        // the group only has candidates to commit (its trajectories finished in the previous iteration): plain copy
        for (int t = wid; t <= N; t += NW) {
            if (w >= 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) X[((size_t)t * n + i) * sB + b] = xs[((size_t)t * n + i) * sB];
                if (t < N) {
#pragma unroll
                    for (int j = 0; j < m; ++j) U[((size_t)t * m + j) * sB + b] = us[((size_t)t * m + j) * sB];
                }
            }
        }
        return;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 32); mbar_init(&empty[s], 64); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#if ILQR_TRIG_TABLE
    if constexpr (Sys::TRIG_TABLE) trig_table_init();
#endif
    __syncthreads();

    if (wid == 2 || wid == 3) {
        // ---------------- consumers: the reverse scan, each half's columns of the value function in registers ----------------
        const T mu_b = mu ? mu[b] : T(0);                    // regularisation (RegArgs), 0 in the reference
        const int H = wid == 2 ? 0 : 1;
        T Vx[n], Vc[n][h];
        {
            T xN[n];
#pragma unroll
            for (int i = 0; i < n; ++i) xN[i] = xs[((size_t)N * n + i) * sB];
            if (w >= 0 && H == 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) X[((size_t)N * n + i) * sB + b] = xN[i];
            }
            qc.terminal_grad(xN, Vx);                        // iLQR_class.py:136-138 (both halves: the full V_x)
#pragma unroll
            for (int i = 0; i < n; ++i)
#pragma unroll
                for (int jo = 0; jo < h; ++jo) Vc[i][jo] = H == 0 ? qc.Qfs[i][jo] : qc.Qfs[i][h + jo];
        }
        T *Kp = K + (size_t)(N - 1) * m * n * sB + b, *kp = k + (size_t)(N - 1) * m * sB + b;
        for (int i = 0; i < N; ++i) {
            const int s = i % S;
            mbar_wait(&full[s], (unsigned)(i / S) & 1u);
            BwdIn<T, n, m> cur;
            {
                const T *st = ring + (size_t)s * L * 32 + lane;
                int row = 0;
#pragma unroll
                for (int r = 0; r < n; ++r)
#pragma unroll
                    for (int c = 0; c < n; ++c, ++row) cur.A[r][c] = st[row * 32];
#pragma unroll
                for (int r = 0; r < n; ++r)
#pragma unroll
                    for (int c = 0; c < m; ++c, ++row) cur.Bd[r][c] = st[row * 32];
#pragma unroll
                for (int r = 0; r < n; ++r, ++row) cur.x[r] = st[row * 32];
#pragma unroll
                for (int c = 0; c < m; ++c, ++row) cur.u[c] = st[row * 32];
            }
            mbar_arrive(&empty[s]);                          // the stage is in registers: hand it back at once
            T Kc[m][h], kt[m];
            T *xb = xch + (size_t)(i & 1) * 2 * XR * 32;     // the step's exchange buffer: [half][XR][32]
            if (H == 0) riccati_half<Cost, T, n, m, 0>(qc, cur, mu_b, Vx, Vc, Kc, kt, xb, xb + XR * 32, lane);
            else riccati_half<Cost, T, n, m, 1>(qc, cur, mu_b, Vx, Vc, Kc, kt, xb + XR * 32, xb, lane);
            if (act) {
#pragma unroll
                for (int j = 0; j < m; ++j) {
#pragma unroll
                    for (int co = 0; co < h; ++co) Kp[((size_t)j * n + H * h + co) * sB] = Kc[j][co];
                    if (H == 0) kp[(size_t)j * sB] = kt[j];
                }
            }
            Kp -= (size_t)m * n * sB;
            kp -= (size_t)m * sB;
        }
        return;
    }

    // ---------------- producers: commit + linearization, NP steps apart (as in fused_backward_kernel) ----------------
    const int p = wid < 2 ? wid : wid - 2;
    const T ph = phi ? phi[b] : T(0);
    const T *px = xs + (size_t)(N - 1 - p) * n * sB, *pu = us + (size_t)(N - 1 - p) * m * sB;
    T *qx = X + (size_t)(N - 1 - p) * n * sB + b, *qu = U + (size_t)(N - 1 - p) * m * sB + b;
    const size_t dx = (size_t)NP * n * sB, du = (size_t)NP * m * sB;
    T *mypre = pre + (size_t)p * 2 * (n + m) * 32;
    auto prefetch = [&](int buf) {
        T *dst = mypre + (size_t)buf * (n + m) * 32 + lane;
#pragma unroll
        for (int r = 0; r < n; ++r) cp_async<sizeof(T)>(dst + r * 32, px + (size_t)r * sB);
#pragma unroll
        for (int c = 0; c < m; ++c) cp_async<sizeof(T)>(dst + (n + c) * 32, pu + (size_t)c * sB);
        px -= dx;
        pu -= du;
    };
    if (p < N) prefetch(0);
    cp_async_commit();
    int buf = 0;
    for (int i = p; i < N; i += NP) {
        const int t = N - 1 - i, s = i % S;
        if (i + NP < N) prefetch(buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();                                  // this step's inputs have landed (own copies only)
        T x[n], u[m];
        {
            const T *src = mypre + (size_t)buf * (n + m) * 32 + lane;
#pragma unroll
            for (int r = 0; r < n; ++r) x[r] = src[r * 32];
#pragma unroll
            for (int c = 0; c < m; ++c) u[c] = src[(n + c) * 32];
        }
        buf ^= 1;
        if (w >= 0) {                                        // commit the accepted candidate into the nominal
#pragma unroll
            for (int r = 0; r < n; ++r) qx[(size_t)r * sB] = x[r];
#pragma unroll
            for (int c = 0; c < m; ++c) qu[(size_t)c * sB] = u[c];
        }
        qx -= dx;
        qu -= du;
        mbar_wait(&empty[s], ((unsigned)(i / S) & 1u) ^ 1u);
        T *st = ring + (size_t)s * L * 32 + lane;
#pragma unroll
        for (int r = 0; r < n; ++r) st[(n * n + n * m + r) * 32] = x[r];
#pragma unroll
        for (int c = 0; c < m; ++c) st[(n * n + n * m + n + c) * 32] = u[c];
        T Aj[n][n], Bj[n][m];
        step_jac<INTEG>(sys, qc.dt, x, u, Aj, Bj, sys.time_scalar(t, ph));
#pragma unroll
        for (int r = 0; r < n; ++r) {
#pragma unroll
            for (int c = 0; c < n; ++c) st[(r * n + c) * 32] = Aj[r][c];
#pragma unroll
            for (int c = 0; c < m; ++c) st[(n * n + r * m + c) * 32] = Bj[r][c];
        }
        mbar_arrive(&full[s]);
    }
}

}  // namespace ilqr
