// ilqr_kernels_backward.cuh -- K2, the reverse Riccati scan: thread-per-trajectory kernel with a shared-memory ring (filled by
// bulk copies onto mbarriers at large batches, by per-thread cp.async otherwise), four-lane
// kernel for small batches of n=4/m=1, sixteen-lane kernel for the n=12/m=4 LTV model
// Part of libilqr_b200.so; included by ilqr_b200.cu only (see the file map at its top).
#pragma once
#include "ilqr_systems.cuh"
#include "ilqr_kernels_common.cuh"

namespace ilqr {

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

// rows of one ring stage: A (n*n), Bd (n*m), x (n), u (m); `stage` points at this thread's element of row 0, rows are
// `bd` elements apart (filled by backward_kernel's issue())
template <int bd, typename T, int n, int m>
ILQR_DEV void bwd_read(BwdIn<T, n, m> &d, const T *stage)
{
    int row = 0;
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j, ++row) d.A[i][j] = stage[row * bd];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < m; ++j, ++row) d.Bd[i][j] = stage[row * bd];
#pragma unroll
    for (int i = 0; i < n; ++i, ++row) d.x[i] = stage[row * bd];
#pragma unroll
    for (int j = 0; j < m; ++j, ++row) d.u[j] = stage[row * bd];
}

// One step of the reverse scan (iLQR_class.py:92-119) on the inputs `cur` = (A_t, B_t, x_t, u_t): updates the value
// function (V_x, V_xx) in place and returns the gains K_t, k_t.  Shared by backward_kernel and the fused
// linearize+backward kernel so that both execute the same operation sequence.
template <class Cost, typename T, int n, int m>
ILQR_DEV void riccati_step(const Cost &qc, const BwdIn<T, n, m> &cur, T mu_b, T *Vx, T (*Vxx)[n], T (*Kt)[n], T *kt)
{
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
}

template <class Cost, typename T, int n, int m, int DEPTH, int BS>
__global__ void __launch_bounds__(BS) backward_kernel(const __grid_constant__ Cost qc, int N, int B, const T *__restrict__ X,
                                const T *__restrict__ U, const T *__restrict__ A, const T *__restrict__ Bd,
                                T *__restrict__ K, T *__restrict__ k, const int *__restrict__ active,
                                const unsigned int *__restrict__ gate, const T *__restrict__ mu,
                                const __grid_constant__ SparseArgs sa, int ab_blocked)
{
    constexpr int L = n * n + n * m + n + m;
    extern __shared__ __align__(16) unsigned char ring_raw[];
    T *ring = reinterpret_cast<T *>(ring_raw);
    if (gate && *gate == 0u) return;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int cAB = b;                                                         // column of A, Bd
    if (sparse_now(sa)) {                                                // few active trajectories: walk their list
        if (sa.only == 1) return;                                        // ... which the four-lane kernel does better
        if ((unsigned int)b >= *sa.n_cur) return;
        b = sa.cur[b];                                                   // A_t, B_t stay at the list position (K1)
    }
    if (b >= B) return;
    constexpr int stage_elems = L * BS, R = n * n + n * m, W = BS / 32;
    // A ring stage holds one chunk per warp, [warp][row][lane]: the thread's element of row i is wbase[i * 32].
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wbase = warp * L * 32 + lane;
    // BULK form (ab_blocked == 2: blocked A_t, B_t, B a multiple of 32, chosen by the host; dense iterations only).  In
    // the blocked layout the R rows of [A_t | B_t] of a warp's 32 trajectories are ONE contiguous chunk (R * 32 elements)
    // and each row of x_t, u_t is 32 contiguous elements: an elected lane fetches the warp's step with 1 + n + m bulk
    // copies (cp.async.bulk -- the TMA unit -- SASS UBLKCP) that complete on a per-(stage, warp) mbarrier, instead of
    // L eight-byte LDGSTS per thread.  Every lane of the warp stays in the loop (stores masked by `valid`).
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + DEPTH * stage_elems) + warp;   // [stage][W]
    const bool bulk = ab_blocked == 2 && !sparse_now(sa);
    bool valid = true;
    if (bulk) {
        valid = !active || active[b] != 0;
        if (!__any_sync(0xffffffffu, valid)) return;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) mbar_init(bars + s * W, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    } else if (active && !active[b]) {
        return;
    }
    // Sources of the ring: timesteps are requested in strictly decreasing order, so each is a pointer that steps
    // back by a fixed stride.  In the solver's blocked layout (ab_off) the R rows of [A_t | B_t] of one trajectory are
    // 32 elements apart: ONE pointer with compile-time offsets serves all of them (the index arithmetic of 25 copies
    // per step used to be 60 % of the kernel's instructions).
    const int c0 = bulk ? (cAB & ~31) : cAB, b0 = bulk ? (b & ~31) : b;            // bulk: the warp's first column
    const T *pab = ab_blocked ? A + ab_off(R, N - 1, 0, c0, B) : nullptr;
    const size_t ab_dec = (((size_t)B + 31) >> 5) * R * 32;
    const T *px = X + (size_t)(N - 1) * n * B + b0, *pu = U + (size_t)(N - 1) * m * B + b0;
    const size_t sB = (size_t)B;
    auto issue = [&](T *stage, int t) {
        T *dst = stage + wbase;
        if (ab_blocked) {
#pragma unroll
            for (int i = 0; i < R; ++i) cp_async<sizeof(T)>(dst + i * 32, pab + i * 32);
            pab -= ab_dec;
        } else {
#pragma unroll
            for (int i = 0; i < n * n; ++i) cp_async<sizeof(T)>(dst + i * 32, A + ((size_t)t * n * n + i) * B + cAB);
#pragma unroll
            for (int i = 0; i < n * m; ++i) cp_async<sizeof(T)>(dst + (n * n + i) * 32, Bd + ((size_t)t * n * m + i) * B + cAB);
        }
#pragma unroll
        for (int i = 0; i < n; ++i) cp_async<sizeof(T)>(dst + (R + i) * 32, px + i * sB);
#pragma unroll
        for (int i = 0; i < m; ++i) cp_async<sizeof(T)>(dst + (R + n + i) * 32, pu + i * sB);
        px -= n * sB;
        pu -= m * sB;
    };
    auto issue_bulk = [&](int s) {                                       // lane 0 of each warp
        T *dst = ring + s * stage_elems + warp * L * 32;
        unsigned long long *bar = bars + s * W;
        mbar_arrive_expect_tx(bar, (unsigned)(L * 32 * sizeof(T)));
        bulk_load(dst, pab, (unsigned)(R * 32 * sizeof(T)), bar);
#pragma unroll
        for (int i = 0; i < n; ++i) bulk_load(dst + (R + i) * 32, px + i * sB, (unsigned)(32 * sizeof(T)), bar);
#pragma unroll
        for (int i = 0; i < m; ++i) bulk_load(dst + (R + n + i) * 32, pu + i * sB, (unsigned)(32 * sizeof(T)), bar);
        pab -= ab_dec;
        px -= n * sB;
        pu -= m * sB;
    };
    if (bulk) {
        if (lane == 0)
            for (int s = 0; s < DEPTH && N - 1 - s >= 0; ++s) issue_bulk(s);
    } else {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) {
            if (N - 1 - s >= 0) issue(ring + s * stage_elems, N - 1 - s);
            cp_async_commit();
        }
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
    unsigned phase = 0;
    for (int t = N - 1; t >= 0; --t) {
        if (bulk) mbar_wait(bars + stage * W, phase);                     // the warp's chunk of step t has landed
        else cp_async_wait<DEPTH - 1>();                                  // the group holding step t has landed
        bwd_read<32>(cur, ring + stage * stage_elems + wbase);
        T Kt[m][n], kt[m];
        riccati_step<Cost, T, n, m>(qc, cur, mu_b, Vx, Vxx, Kt, kt);
        if (bulk) {
            // refill the stage: every lane's shared-memory reads of it have been CONSUMED by the arithmetic above (so they
            // have completed, not merely been issued, before the copy engine may overwrite the stage)
            __syncwarp();
            if (lane == 0 && t - DEPTH >= 0) issue_bulk(stage);
        }
        if (valid) {
#pragma unroll
            for (int j = 0; j < m; ++j) {
#pragma unroll
                for (int i = 0; i < n; ++i) K[(((size_t)t * m + j) * n + i) * B + b] = Kt[j][i];
                k[((size_t)t * m + j) * B + b] = kt[j];
            }
        }
        if (!bulk) {
            if (t - DEPTH >= 0) issue(ring + stage * stage_elems, t - DEPTH);
            cp_async_commit();
        }
        if (++stage == DEPTH) { stage = 0; phase ^= 1u; }
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
                           const unsigned int *__restrict__ gate, const T *__restrict__ mu,
                           const __grid_constant__ SparseArgs sa, int ab_blocked)
{
    constexpr int n = 4, LP = 26, SLOTS = 8;       // LP: padded slot stride (bank-conflict free LDS.128)
    extern __shared__ __align__(16) unsigned char lanes_raw[];
    T *ring = reinterpret_cast<T *>(lanes_raw);             // [DEPTH][SLOTS][LP]
    T *exQ = ring + DEPTH * SLOTS * LP;                     // [SLOTS][4]   Q_ux exchange
    T *exV = exQ + SLOTS * 4;                               // [SLOTS][20]  V_xx (row-major 16) + V_x (4)
    if (gate && *gate == 0u) return;
    const int lane = threadIdx.x, s = lane >> 2, j = lane & 3;
    // slot -> trajectory: the batch index itself, or an entry of the active list when few trajectories are left
    const bool sparse = sparse_now(sa);
    if (sa.only == 2 && !sparse) return;                    // large batch: dense iterations belong to backward_kernel
    const int n_items = sparse ? (int)*sa.n_cur : B;
    auto item = [&](int slot) -> int {
        const int i = min(blockIdx.x * SLOTS + slot, n_items - 1);      // clamped: spare slots compute on a copy
        return sparse ? sa.cur[i] : i;
    };
    const int b = item(s);
    const bool valid = blockIdx.x * SLOTS + s < n_items && (!active || active[b] != 0);
    if (__ballot_sync(0xffffffffu, valid) == 0u) return;

    // cooperative fill of one ring stage: element e = row * 8 + slot, 32 elements per LDGSTS.  Lane (f_row0, f_slot)
    // copies rows f_row0 + 4 i of slot f_slot: i < 4 are rows of A, i = 4 of B, i = 5 of x, i = 6 is u (f_row0 == 0
    // only) -- so every source is a base pointer fixed before the scan plus t times a fixed stride.
    const int f_slot = lane & 7, f_row0 = lane >> 3;
    const int f_b = item(f_slot);
    // column of A, Bd: the trajectory, or in a sparse iteration its list position (where K1 wrote them, coalesced)
    const int f_c = sparse ? min(blockIdx.x * SLOTS + f_slot, n_items - 1) : f_b;
    const T *pA = ab_blocked ? A + ab_off(20, 0, f_row0, f_c, B) : A + (size_t)f_row0 * B + f_c;
    const T *pB = ab_blocked ? A + ab_off(20, 0, 16 + f_row0, f_c, B) : Bd + (size_t)f_row0 * B + f_c;
    const T *pX = X + (size_t)f_row0 * B + f_b, *pU = U + f_b;
    const size_t ab_step = (((size_t)B + 31) >> 5) * 20 * 32;           // blocked: elements per timestep
    const size_t tsA = ab_blocked ? ab_step : (size_t)16 * B, tsB = ab_blocked ? ab_step : (size_t)4 * B;
    const size_t rsA = ab_blocked ? (size_t)4 * 32 : (size_t)4 * B;     // four rows further
    T *const f_dst = ring + f_slot * LP + f_row0;
    auto issue = [&](int stage, int t) {
        T *dst = f_dst + stage * SLOTS * LP;
        const T *a = pA + (size_t)t * tsA;
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async<sizeof(T)>(dst + 4 * i, a + i * rsA);
        cp_async<sizeof(T)>(dst + 16, pB + (size_t)t * tsB);
        cp_async<sizeof(T)>(dst + 20, pX + (size_t)t * 4 * B);
        if (f_row0 == 0) cp_async<sizeof(T)>(dst + 24, pU + (size_t)t * B);
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
#ifndef ILQR_LANES_PREFETCH
#define ILQR_LANES_PREFETCH 1
#endif
    // Software pipeline: the inputs of step t-1 are read from the ring into registers at the top of step t, so their
    // shared-memory latency (and the cp.async wait) hides behind step t's arithmetic instead of heading the
    // dependent chain of step t-1; the ring slot of step t is refilled as soon as step t starts.
    struct StepIn { T Am[n][n], Bv[n], x[n], Acol[n], u; };
    auto read_in = [&](StepIn &d, int st) {
        const T *in = ring + (st * SLOTS + s) * LP;
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int c = 0; c < n; ++c) d.Am[i][c] = in[i * 4 + c];
            d.Bv[i] = in[16 + i];
            d.x[i] = in[20 + i];
            d.Acol[i] = in[i * 4 + j];
        }
        d.u = in[24];
    };
    int stage = 0;
    // one scan step on the inputs `cur`; with the prefetch, `nxt` receives the inputs of step t-1 meanwhile
    auto scan_step = [&](const StepIn &cur, StepIn &nxt, int t) {
#if ILQR_LANES_PREFETCH
        __syncwarp();                                                    // every lane holds its copy of this slot
        if (t - DEPTH >= 0) issue(stage, t - DEPTH);
        cp_async_commit();
        stage = (stage + 1 == DEPTH) ? 0 : stage + 1;
        if (t > 0) {
            cp_async_wait<DEPTH - 1>();
            __syncwarp();                                                // other lanes' copies are visible
            read_in(nxt, stage);
        }
#endif
        const T (&Am)[n][n] = cur.Am;
        const T (&Bv)[n] = cur.Bv, (&x)[n] = cur.x, (&Acol)[n] = cur.Acol;
        const T u = cur.u;
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
#if !ILQR_LANES_PREFETCH
        if (t - DEPTH >= 0) issue(stage, t - DEPTH);                     // every lane is past its reads of this stage
        cp_async_commit();
        stage = (stage + 1 == DEPTH) ? 0 : stage + 1;
#endif
    };
#if ILQR_LANES_PREFETCH
    StepIn in_a, in_b;                                                   // ping-pong: no register copies between steps
    cp_async_wait<DEPTH - 1>();
    __syncwarp();
    read_in(in_a, 0);
    for (int t = N - 1; t >= 0; t -= 2) {
        scan_step(in_a, in_b, t);
        if (t - 1 >= 0) scan_step(in_b, in_a, t - 1);
    }
#else
    for (int t = N - 1; t >= 0; --t) {
        cp_async_wait<DEPTH - 1>();
        __syncwarp();                                                    // other lanes' copies are visible
        StepIn cur;
        read_in(cur, stage);
        scan_step(cur, cur, t);
    }
#endif
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

// K2 for the LTV model, register-tiled (round 2).  The sixteen-lane kernel above is bound by shared-memory bandwidth:
// every lane re-reads all of V_xx and all of [A_t|B_t]' for ONE column (ncu: the LSU data pipe, not the FP64 pipe).
// Here FOUR lanes share a trajectory and lane c owns FOUR columns of [A_t | B_t]: 4c .. 4c+3 (lanes 0-2: columns of
// A_t, lane 3: the four columns of B_t), so every value read from shared memory feeds four FMAs instead of one:
//   A  build the own columns of A_t = I + dt (Ac + w_t E) pairwise from the block's constant tables, publish them as
//      rows of MT = [A_t|B_t]' ; q = own columns' dot with V_x                             (iLQR_class.py:100-101)
//   B  W[:, own] = V_xx M[:, own]            two passes over V_xx, 2 columns x 12 rows of accumulators each
//   C  G[r][own] = MT[r][:] W[:, own]  r = 0..15: rows < 12 give Q_xx[:, own] (written over V_xx[:, own], which nobody
//      reads any more), rows >= 12 give Q_ux[:, own] -- or, in lane 3, Q_uu                          (:102-104)
//   D  exchange Q_ux, Q_uu, Q_u through shared memory; every lane factors the 4x4 Q_uu (LU with partial pivoting, as
//      the reference's solve) and solves for its own four right-hand sides: K[:, own] (lane 3: k)    (:109-110)
//   E  V_xx[:, own] += Q_ux' K[:, own],  V_x[own] = Q_x[own] + K[:, own]' Q_u                       (:113-114)
// Lanes of a trajectory sit in one warp (__syncwarp between the phases); one block barrier per step for the staged,
// coalesced gain stores and the x_t, u_t prefetch, as in the sixteen-lane kernel.
#ifndef ILQR_LTV4_MINBLOCKS
#define ILQR_LTV4_MINBLOCKS 1
#endif
// Shared-memory layout of the register-tiled LTV kernel.  A 128-bit shared load is served eight lanes at a time, i.e. two
// trajectories (four lanes each) per phase: the per-trajectory strides are == 16 bytes (mod 128) so that the two never
// meet in a bank, and the tables whose rows are owned by different lanes (MT, AcdT, EdT: rows 4c..4c+3 belong to lane c)
// skew each group of four rows by 16 bytes for the same reason.
struct Ltv4Layout {
    static constexpr int VXX = 146;        // [12][12] + 2
    static constexpr int MT = 162;         // 12 rows of 12, group of four rows skewed by 2: row r at r*12 + 2*(r/4); 150 + 12
    static constexpr int VX = 14, QUX = 66 /* [12][4] transposed, each lane's four rows skewed by 2 */, QUU = 18, QU = 6, XS = 18;
    static constexpr int KROW = 17;        // gain stage [52][KROW >= TPB]: rows 4 apart (lanes c, c+1) land 32 bytes apart
    static constexpr int PER_TRAJ = VXX + MT + VX + QUX + QUU + QU + 2 * XS;
    static constexpr int KSTAGE = 2 * 52 * KROW;
    static constexpr int TAB = 150;        // AcdT / EdT: 12 skewed rows
    static constexpr int CONST = KSTAGE + 48 + 2 * TAB + 144 + 16;
    __host__ __device__ static constexpr int row(int r) { return r * 12 + 2 * (r >> 2); }
};

template <typename T, int TPB>
__global__ void __launch_bounds__(TPB * 4, ILQR_LTV4_MINBLOCKS)
backward_ltv4_kernel(const __grid_constant__ LtvSys<T> sys, const __grid_constant__ QuadCost<T, 12, 4> qc, int N, int B,
                     const T *__restrict__ phi, const T *__restrict__ X, const T *__restrict__ U, T *__restrict__ K,
                     T *__restrict__ k, const int *__restrict__ active, const unsigned int *__restrict__ gate,
                     const T *__restrict__ mu)
{
    constexpr int n = 12, m = 4, NT = TPB * 4, ROWS = n * m + m;     // 52 gain rows per step
    using V2 = typename Vec2<T>::type;
    using LY = Ltv4Layout;
    static_assert(TPB <= LY::KROW, "gain stage rows hold TPB trajectories");
    extern __shared__ __align__(16) unsigned char ltv4_raw[];
    T *sm = reinterpret_cast<T *>(ltv4_raw);
    T *VxxS = sm;                               // [TPB] row-major V_xx (Q_xx during a step)
    T *MTS = VxxS + TPB * LY::VXX;              // [TPB] MT[r][l] = A_t[l][r], r < 12 (skewed rows)
    T *VxS = MTS + TPB * LY::MT;                // [TPB][12]
    T *QuxS = VxS + TPB * LY::VX;               // [TPB][12][4]   Q_ux transposed: QuxT[i][u]
    T *QuuS = QuxS + TPB * LY::QUX;             // [TPB][4][4]
    T *QuS = QuuS + TPB * LY::QUU;              // [TPB][4]
    T *xsS = QuS + TPB * LY::QU;                // [2][TPB][16]   x_t (12), u_t (4), double buffered
    T *KS = xsS + 2 * TPB * LY::XS;             // [2][52][KROW]  gain stage, double buffered
    T *BdT = KS + LY::KSTAGE;               // [4][12]        rows 12..15 of MT: BdT[j][l] = dt Bc[l][j]   (constant)
    T *AcdT = BdT + 48;                         // skewed rows    AcdT[c][l] = delta(l,c) + dt Ac[l][c]
    T *EdT = AcdT + LY::TAB;                    // skewed rows    EdT[c][l]  = dt E[l][c]
    T *QsS = EdT + LY::TAB;                     // [12][12]       symmetrised Q times dt
    T *RsS = QsS + 144;                         // [4][4]         symmetrised R times dt
    __shared__ int vflag[TPB];
    if (gate && *gate == 0u) return;
    const int tid = threadIdx.x, s = tid >> 2, c = tid & 3;
    const int b0 = blockIdx.x * TPB;
    const int b_raw = b0 + s;
    const bool valid = b_raw < B && (!active || active[b_raw] != 0);
    if (__syncthreads_or(valid) == 0) return;
    const int b = b_raw < B ? b_raw : B - 1;        // out-of-range / inactive slots compute on a copy, never store
    if (c == 0) vflag[s] = valid;
    for (int e = tid; e < 48; e += NT) BdT[e] = qc.dt * sys.Bc[e % 12][e / 12];
    for (int e = tid; e < 144; e += NT) {
        const int cc = e / 12, l = e % 12;
        QsS[e] = qc.Qs[cc][l] * qc.dt;
        AcdT[LY::row(cc) + l] = (cc == l ? T(1) : T(0)) + qc.dt * sys.Ac[l][cc];
        EdT[LY::row(cc) + l] = qc.dt * sys.E[l][cc];
    }
    if (tid < 16) RsS[tid] = qc.Rs[tid >> 2][tid & 3] * qc.dt;
    // staged loads: element e = row * TPB + bb -> row (x_0..x_11, u_0..u_3) of trajectory b0 + bb
    auto fetch = [&](int t, int e) -> T {
        const int row = e / TPB, bb = e % TPB;
        const int lb = min(b0 + bb, B - 1);
        if (row < n) return X[((size_t)t * n + row) * B + lb];
        return t < N ? U[((size_t)t * m + (row - n)) * B + lb] : T(0);
    };
    // 16 TPB values per step over 4 TPB threads: four per thread, fetched at the top of a step into registers and put
    // into the other staging buffer just before the step's block barrier (the global latency hides behind the step)
    T pre[4];
    auto fetch_pre = [&](int t) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pre[i] = fetch(t, tid + i * NT);
    };
    auto store_pre = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * NT;
            xsS[(buf * TPB + e % TPB) * LY::XS + e / TPB] = pre[i];
        }
    };
    const T ph = phi ? phi[b] : T(0);
    const T mu_b = mu ? mu[b] : T(0);                                    // regularisation (RegArgs), 0 in the reference
    T *Vxx = VxxS + s * LY::VXX, *MT = MTS + s * LY::MT, *Vx = VxS + s * LY::VX, *QuxT = QuxS + s * LY::QUX,
      *Quu = QuuS + s * LY::QUU, *Qu = QuS + s * LY::QU;
    // terminal condition (iLQR_class.py:136-138): V_x = Q_f (x_N - x_target), V_xx = Q_f ; lane c fills rows 3c..3c+2
    fetch_pre(N);
    store_pre(0);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = 3 * c + r;
        T g = T(0);
#pragma unroll
        for (int j = 0; j < n; ++j) {
            g += qc.Qfs[i][j] * (xsS[s * LY::XS + j] - qc.xt[j]);
            Vxx[i * 12 + j] = qc.Qfs[i][j];
        }
        Vx[i] = g;
    }
    __syncthreads();
    fetch_pre(N - 1);
    store_pre(1);                        // buffer index of step t = (N - t) & 1
    __syncthreads();
    for (int t = N - 1; t >= 0; --t) {
        const int buf = (N - t) & 1;
        const T *xs = xsS + (buf * TPB + s) * LY::XS;
        const T w = sys.time_scalar(t, ph);
        if (t > 0) fetch_pre(t - 1);
        // ---- A + B: own columns of [A_t|B_t] two at a time; W[:, own] = V_xx M[:, own]; q = M[:, own]' V_x ----
        T Wc[4][n], q[4];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            T Mc[2][n];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * pr + jj;
                if (c < 3) {
                    const int ro = LY::row(4 * c + j);
                    const T *ar = AcdT + ro, *er = EdT + ro;
#pragma unroll
                    for (int l = 0; l < n; l += 2) {
                        const V2 a = *reinterpret_cast<const V2 *>(ar + l), e = *reinterpret_cast<const V2 *>(er + l);
                        Mc[jj][l] = fma_t(w, e.x, a.x);
                        Mc[jj][l + 1] = fma_t(w, e.y, a.y);
                        *reinterpret_cast<V2 *>(MT + ro + l) = V2{Mc[jj][l], Mc[jj][l + 1]};
                    }
                } else {
#pragma unroll
                    for (int l = 0; l < n; l += 2) {
                        const V2 v = *reinterpret_cast<const V2 *>(BdT + j * 12 + l);
                        Mc[jj][l] = v.x;
                        Mc[jj][l + 1] = v.y;
                    }
                }
            }
            {
                T q0 = T(0), q1 = T(0), q2 = T(0), q3 = T(0);     // two chains per column
#pragma unroll
                for (int l = 0; l < n; l += 2) {
                    const V2 v = *reinterpret_cast<const V2 *>(Vx + l);
                    q0 += Mc[0][l] * v.x;
                    q1 += Mc[1][l] * v.x;
                    q2 += Mc[0][l + 1] * v.y;
                    q3 += Mc[1][l + 1] * v.y;
                }
                q[2 * pr] = q0 + q2;
                q[2 * pr + 1] = q1 + q3;
            }
            // four rows of V_xx at a time: eight independent accumulation chains keep the FP64 pipe fed from one or two
            // warps per sub-partition
#pragma unroll
            for (int i0 = 0; i0 < n; i0 += 4) {
                T acc[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = T(0);
#pragma unroll
                for (int l = 0; l < n; l += 2) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const V2 v = *reinterpret_cast<const V2 *>(Vxx + (i0 + r) * 12 + l);
                        acc[r][0] += v.x * Mc[0][l];
                        acc[r][1] += v.x * Mc[1][l];
                        acc[r][0] += v.y * Mc[0][l + 1];
                        acc[r][1] += v.y * Mc[1][l + 1];
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    Wc[2 * pr][i0 + r] = acc[r][0];
                    Wc[2 * pr + 1][i0 + r] = acc[r][1];
                }
                asm volatile("" ::: "memory");       // keep ptxas from hoisting every row's loads (it spills W otherwise)
            }
        }
        __syncwarp();                    // every lane's MT rows are published; nobody reads V_xx any more
        // ---- C: G[r][own] = MT[r][:] W[:, own] ----
        // rows < 12: Q_xx[r][own] = l_xx[r][own] + G (lane 3's block, f_x' V_xx f_u, is not used by the reference).  Two
        // rows per trip (eight chains), not unrolled further: the 48 values of W stay in registers, MT passes through
#pragma unroll 1
        for (int r = 0; r < n; r += 2) {
            const T *row = MT + LY::row(r);
            T g[2][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) g[0][j] = g[1][j] = T(0);
#pragma unroll
            for (int l = 0; l < n; l += 2) {
                const V2 v0 = *reinterpret_cast<const V2 *>(row + l), v1 = *reinterpret_cast<const V2 *>(row + 12 + l);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[0][j] += v0.x * Wc[j][l];
                    g[1][j] += v1.x * Wc[j][l];
                    g[0][j] += v0.y * Wc[j][l + 1];
                    g[1][j] += v1.y * Wc[j][l + 1];
                }
            }
            if (c < 3) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        const V2 lq = *reinterpret_cast<const V2 *>(QsS + (r + rr) * 12 + 4 * c + j);
                        *reinterpret_cast<V2 *>(Vxx + (r + rr) * 12 + 4 * c + j) = V2{lq.x + g[rr][j], lq.y + g[rr][j + 1]};
                    }
            }
        }
        T Gu[m][4];                      // rows 12..15: Q_ux[:, own] (lanes 0-2) or Q_uu (lane 3), without l_uu
#pragma unroll
        for (int u = 0; u < m; ++u) {
            const T *row = BdT + u * 12;
#pragma unroll
            for (int j = 0; j < 4; ++j) Gu[u][j] = T(0);
#pragma unroll
            for (int l = 0; l < n; l += 2) {
                const V2 v = *reinterpret_cast<const V2 *>(row + l);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    Gu[u][j] += v.x * Wc[j][l];
                    Gu[u][j] += v.y * Wc[j][l + 1];
                }
            }
        }
        // cost gradient of the own entries and Q_x / Q_u  (QsS, RsS already carry dt)      (:100-101)
        T Qc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            T g = T(0);
            if (c < 3) {
                const int i = 4 * c + j;
                if (qc.diag) g = QsS[i * 12 + i] * (xs[i] - qc.xt[i]);
                else {
#pragma unroll
                    for (int jj = 0; jj < n; ++jj) g += QsS[i * 12 + jj] * (xs[jj] - qc.xt[jj]);
                }
            } else {
                if (qc.diag) g = RsS[j * 4 + j] * xs[n + j];
                else {
#pragma unroll
                    for (int jj = 0; jj < m; ++jj) g += RsS[j * 4 + jj] * xs[n + jj];
                }
            }
            Qc[j] = g + q[j];
        }
        // ---- D: exchange, 4x4 solve ----
        if (c < 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {               // QuxT[i][u], i = 4c + j
                *reinterpret_cast<V2 *>(QuxT + (4 * c + j) * 4 + 2 * c) = V2{Gu[0][j], Gu[1][j]};
                *reinterpret_cast<V2 *>(QuxT + (4 * c + j) * 4 + 2 * c + 2) = V2{Gu[2][j], Gu[3][j]};
            }
        } else {
#pragma unroll
            for (int u = 0; u < m; ++u) {
                const V2 r0 = *reinterpret_cast<const V2 *>(RsS + u * 4), r1 = *reinterpret_cast<const V2 *>(RsS + u * 4 + 2);
                *reinterpret_cast<V2 *>(Quu + u * 4) = V2{(r0.x + Gu[u][0]) + (u == 0 ? mu_b : T(0)), (r0.y + Gu[u][1]) + (u == 1 ? mu_b : T(0))};
                *reinterpret_cast<V2 *>(Quu + u * 4 + 2) = V2{(r1.x + Gu[u][2]) + (u == 2 ? mu_b : T(0)), (r1.y + Gu[u][3]) + (u == 3 ? mu_b : T(0))};
            }
            *reinterpret_cast<V2 *>(Qu) = V2{Qc[0], Qc[1]};
            *reinterpret_cast<V2 *>(Qu + 2) = V2{Qc[2], Qc[3]};
        }
        __syncwarp();
        T Lm[m][m], rhs[m][4], Quv[m];
        {
            const V2 a = *reinterpret_cast<const V2 *>(Qu), bq = *reinterpret_cast<const V2 *>(Qu + 2);
            Quv[0] = a.x; Quv[1] = a.y; Quv[2] = bq.x; Quv[3] = bq.y;
        }
#pragma unroll
        for (int u = 0; u < m; ++u) {
            const V2 a = *reinterpret_cast<const V2 *>(Quu + u * 4), bq = *reinterpret_cast<const V2 *>(Quu + u * 4 + 2);
            Lm[u][0] = a.x; Lm[u][1] = a.y; Lm[u][2] = bq.x; Lm[u][3] = bq.y;
#pragma unroll
            for (int j = 0; j < 4; ++j) rhs[u][j] = c < 3 ? Gu[u][j] : (j == 0 ? Quv[u] : T(0));
        }
        lu_solve_inplace<m, 4, T, true>(Lm, rhs);
        T *ks = KS + ((N - t) & 1) * ROWS * LY::KROW;
        if (c < 3) {
            // ---- E: V_xx[:, own] = Q_xx[:, own] + Q_ux' K[:, own] ; V_x[own] = Q_x[own] + K[:, own]' Q_u ----
            T Kc[m][4];
#pragma unroll
            for (int u = 0; u < m; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) Kc[u][j] = -rhs[u][j];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const V2 qa = *reinterpret_cast<const V2 *>(QuxT + i * 4 + 2 * (i >> 2)), qb = *reinterpret_cast<const V2 *>(QuxT + i * 4 + 2 * (i >> 2) + 2);
                const T qx[m] = { qa.x, qa.y, qb.x, qb.y };
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    V2 v = *reinterpret_cast<const V2 *>(Vxx + i * 12 + 4 * c + j);
                    T a0 = T(0), a1 = T(0);
#pragma unroll
                    for (int u = 0; u < m; ++u) { a0 += qx[u] * Kc[u][j]; a1 += qx[u] * Kc[u][j + 1]; }
                    v.x += a0;
                    v.y += a1;
                    *reinterpret_cast<V2 *>(Vxx + i * 12 + 4 * c + j) = v;
                }
            }
            T vxn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                T vx = T(0);
#pragma unroll
                for (int u = 0; u < m; ++u) vx += Kc[u][j] * Quv[u];
                vxn[j] = Qc[j] + vx;
#pragma unroll
                for (int u = 0; u < m; ++u) ks[(u * n + 4 * c + j) * LY::KROW + s] = Kc[u][j];
            }
            *reinterpret_cast<V2 *>(Vx + 4 * c) = V2{vxn[0], vxn[1]};
            *reinterpret_cast<V2 *>(Vx + 4 * c + 2) = V2{vxn[2], vxn[3]};
        } else {
#pragma unroll
            for (int u = 0; u < m; ++u) ks[(n * m + u) * LY::KROW + s] = -rhs[u][0];
        }
        if (t > 0) store_pre(buf ^ 1);
        __syncthreads();
        // coalesced store of the step's gains: rows of TPB consecutive trajectories
        for (int e = tid; e < ROWS * TPB; e += NT) {
            const int row = e / TPB, bb = e % TPB;
            if (vflag[bb]) {
                if (row < n * m) K[((size_t)t * n * m + row) * B + b0 + bb] = ks[row * LY::KROW + bb];
                else k[((size_t)t * m + (row - n * m)) * B + b0 + bb] = ks[row * LY::KROW + bb];
            }
        }
    }
}

}  // namespace ilqr
