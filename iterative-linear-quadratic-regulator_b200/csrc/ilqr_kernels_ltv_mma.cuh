// ilqr_kernels_ltv_mma.cuh -- K2 of the LTV model (n = 12, m = 4; BASELINE.json config 4) on the FP64 tensor cores.
//
// The Riccati step of iLQR_class.py:92-119 is two dense products per trajectory and timestep,
//     W = V_xx M,   G = M' W,     M = [A_t | B_t]  (12 x 16),
// G carrying Q_xx - l_xx (rows, columns < 12), Q_ux (rows >= 12, columns < 12) and Q_uu - l_uu (rows, columns >= 12).
// The lane-tiled kernels of ilqr_kernels_backward.cuh feed every FMA of these products from shared memory and are bound
// by the LSU data pipe (profiles/r02_ncu_full_ltv_B32768.txt: 59 % wavefronts against 35 % FP64).  On sm_100a
// `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4) runs at exactly the DFMA rate -- 16 cycles per sub-partition for 256 FMAs,
// scripts/micro/fp64_dmma.cu -- but takes its operands from registers once per 256 FMAs, so the products stop being
// operand-delivery bound.  ONE WARP owns ONE trajectory; everything stays in tensor-core fragments between steps:
//
//   fragment of an 8x8 tile, thread (g = lane / 4, q = lane % 4):   A-operand a = A[g][q]   B-operand b = B[q][g]
//                                                                     accumulator c0, c1 = C[g][2q], C[g][2q + 1]
//   state  V  = 16 x 16 tiles V[rt][ct][e] = Vaug[g + 8 rt][8 ct + 2q + e],  Vaug = [[V_xx, 0], [V_x', 0], [0, 0]]
//              (V_x rides along as row 12, so M' V_x -- Q_x - l_x and Q_u - l_u -- falls out of the first product)
//   MA        = M' as A-operand fragments, k index PERMUTED: in k-step (ct, e) thread q contracts index 8 ct + 2q + e.
//              That is exactly the column an accumulator register (ct, e) of thread q holds, so the accumulators of one
//              product are the B-operands of the next WITHOUT any shuffle:
//   phase 1   Wt[mt][nt] += MA[mt][ct][e] x V[nt][ct][e]        Wt = (Vaug M)' (16 x 16), 16 DMMAs
//   phase 2   G[mr][mt]  += MA[mr][nt][e] x Wt[mt][nt][e]       G  = M' V_xx M (16 x 16), 16 DMMAs
//   phase 3   Q_uu, Q_u, Q_ux, Q_x through the warp's shared-memory patch; every lane factors the 4 x 4 Q_uu (the
//             reference's LU: swap-free fast path, the pivoting routine when a pivot is not its column's largest) and
//             lane c solves for column c of K (lane 12: k)                                                  (:109-110)
//   phase 4   Vaug' = [Q_xx ; Q_x'] + [Q_ux' ; Q_u'] K          4 DMMAs straight into the next step's V fragments (:113-114)
// 36 DMMAs = 9 216 FMA slots for the step's ~6 000 useful FMAs (padding 12 -> 16).  Gains are staged through shared memory
// and stored as rows of WPB consecutive trajectories, one block barrier per step.  FP64 only (the FP32 mode keeps the
// lane-tiled kernels).  Part of libilqr_b200.so; included by ilqr_b200.cu only.
#pragma once
#include "ilqr_systems.cuh"
#include "ilqr_kernels_common.cuh"

namespace ilqr {

ILQR_DEV void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// lu_solve_inplace<4, 1> (ilqr_systems.cuh) WITHOUT the row swaps: the same operations in the same order as long as no
// swap is due, and the return value says whether one was (some |a[i][j]| > |a[j][j]| below a pivot: the caller then
// repeats the solve with the pivoting routine).  Q_uu = R dt + B' V_xx B is symmetric positive definite up to rounding,
// so swaps are rare, and ptxas turns the swaps of the general routine -- selects or, written as branches, predicated
// moves -- into ~120 instructions that issue every step whether or not a swap happens.
ILQR_DEV bool lu4_solve_nopivot(double (*a)[4], double *b)
{
    bool swap_due = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = j + 1; i < 4; ++i) swap_due |= fabs(a[i][j]) > fabs(a[j][j]);
        const double r = rcp_t(a[j][j]);
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            const double l = a[i][j] * r;
#pragma unroll
            for (int c = j + 1; c < 4; ++c) a[i][c] -= l * a[j][c];
            b[i] -= l * b[j];
        }
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        const double r = rcp_t(a[i][i]);
        double s = b[i];
#pragma unroll
        for (int k = i + 1; k < 4; ++k) s -= a[i][k] * b[k];
        b[i] = s * r;
    }
    return swap_due;
}

// register cap: ILQR_LTV_MMA_WARPS warps per SM whatever the block size (20: 96 registers and ~120 bytes of spills, still
// the fastest: 36.1 / 33.0 / 34.9 ms per pass at B=32768, N=1000 for 16 (128 registers, no spills) / 20 / 24)
#ifndef ILQR_LTV_MMA_WARPS
#define ILQR_LTV_MMA_WARPS 20
#endif
template <int WPB>
__global__ void __launch_bounds__(WPB * 32, ILQR_LTV_MMA_WARPS / WPB)
backward_ltv_mma_kernel(const __grid_constant__ LtvSys<double> sys, const __grid_constant__ QuadCost<double, 12, 4> qc,
                        int N, int B, const double *__restrict__ phi, const double *__restrict__ X,
                        const double *__restrict__ U, double *__restrict__ K, double *__restrict__ k,
                        const int *__restrict__ active, const unsigned int *__restrict__ gate,
                        const double *__restrict__ mu)
{
    constexpr int n = 12, m = 4, NT = WPB * 32, ROWS = n * m + m, KROW = WPB + 1;
    // per-warp patch:  TAB [16][4]: rows c < 12 = Q_ux[:, c], rows 12 + v = Q_uu[:, v]  |  QV [16] = Q_x (12), Q_u (4)  |
    //                  Kw [4][KWS] (columns 12..15 stay zero)  |  xs [2][16] = x_t - x_target (12), u_t (4), double buffered
    // KWS = 20: the B-operand read Kw[q][g] of a half warp (g 0..3, q 0..3) lands in sixteen different 8-byte banks
    constexpr int KWS = 20, LXS = 24;
    constexpr int P_TAB = 0, P_QV = 64, P_KW = 80, P_XS = 160, PATCH = 192;
    __shared__ __align__(16) double patchS[WPB][PATCH];
    __shared__ __align__(16) double ksS[2][ROWS][KROW];       // gain stage, double buffered
    // Q dt padded with zeros; row stride 24 doubles = 64 bytes (mod 128): the 16-byte reads of rows g, g + 1 in one
    // quarter warp do not meet in a bank
    __shared__ __align__(16) double lxxS[16][LXS];
    __shared__ __align__(16) double2 maS[8][32];              // (base, ecf) of every lane's eight fragment entries
    __shared__ __align__(16) double RsS[16];                  // R dt
    __shared__ __align__(16) double dgS[16];                  // diagonals of Q dt (12) and R dt (4)
    __shared__ int vflag[WPB];
    if (gate && *gate == 0u) return;
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int b0 = blockIdx.x * WPB, b_raw = b0 + wp;
    const bool valid = b_raw < B && (!active || active[b_raw] != 0);
    if (__syncthreads_or(valid) == 0) return;
    const int b = b_raw < B ? b_raw : B - 1;          // out-of-range / inactive warps compute on a copy, never store
    if (lane == 0) vflag[wp] = valid;
    for (int e = tid; e < 256; e += NT) {
        const int i = e >> 4, j = e & 15;
        lxxS[i][j] = (i < n && j < n) ? qc.Qs[i][j] * qc.dt : 0.0;
        if (j < LXS - 16) lxxS[i][16 + j] = 0.0;
    }
    if (tid < 16) {
        RsS[tid] = qc.Rs[tid >> 2][tid & 3] * qc.dt;
        dgS[tid] = tid < n ? qc.Qs[tid][tid] * qc.dt : qc.Rs[tid - n][tid - n] * qc.dt;
    }
    double *pt = patchS[wp];
    for (int e = lane; e < PATCH; e += 32) pt[e] = 0.0;
    const double ph = phi ? phi[b] : 0.0;
    const double mu_b = mu ? mu[b] : 0.0;             // regularisation (RegArgs), 0 in the reference

    // M' as A-operand fragments: MA[mt][ct][e] = Mpad[kk][r], kk = 8 ct + 2q + e (row of [A_t|B_t], zero for kk >= 12),
    // r = g + 8 mt (column: A_t for r < 12, B_t beyond).  A_t = I + dt (Ac + w_t E), B_t = dt Bc: base + w_t * ecf,
    // the pairs kept in shared memory (32 registers otherwise)
    if (wp == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kk = 8 * ct + 2 * q + e, r = g + 8 * mt;
                    double bs = 0.0, ec = 0.0;
                    if (kk < n) {
                        if (r < n) {
                            bs = (kk == r ? 1.0 : 0.0) + qc.dt * sys.Ac[kk][r];
                            ec = qc.dt * sys.E[kk][r];
                        } else {
                            bs = qc.dt * sys.Bc[kk][r - n];
                        }
                    }
                    maS[(mt * 2 + ct) * 2 + e][lane] = make_double2(bs, ec);
                }
    }
    // loop invariants of the lane.  Everything a lane does differently from its neighbours is an address or an addend
    // fixed here, so that the step below is straight-line code (scripts/micro/fp64_dmma.cu: DMMAs and integer
    // instructions add up rather than overlap; in this kernel 653 -> 415 instructions per step bought 6 %)
    //   x_t, u_t of the own trajectory: lanes 0..15 fetch one value each, one step ahead, and store x - x_target
    const double *fsrc = lane < n ? X + (size_t)lane * B + b : U + (size_t)((lane - n) & 3) * B + b;
    const size_t fstride = (size_t)(lane < n ? n : m) * B;
    const double xt_l = lane < n ? qc.xt[lane] : 0.0;
    //   Q_uu = R dt + G (+ mu on the diagonal) sits in the q >= 2 columns of the tile holding rows and columns 8..15
    //   (zero for the Q_ux entries of the same tile, so that the step needs no case distinction)
    double radd[2], mudd[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int u = (g - 4) & 3, v = (2 * (q - 2) + e) & 3;
        radd[e] = q >= 2 ? qc.Rs[u][v] * qc.dt : 0.0;
        mudd[e] = (q >= 2 && u == v) ? mu_b : 0.0;
    }
    //   right-hand side of the lane's solve: column `lane` of Q_ux, or Q_u (lane 12 and the idle lanes)
    const double *rhs_src = lane < n ? pt + P_TAB + lane * 4 : pt + P_QV + n;
    //   phase 4, A operand of rows 8..15: Q_ux[q][8 + g] for g < 4, Q_u[q] for row 12, anything finite beyond
    const double *ap1_src = g == 4 ? pt + P_QV + n + q : pt + P_TAB + (g + 8) * 4 + q;
    //   gain stage rows of the lane (lanes 0..11: K[u][lane], lane 12: k[u]) and the two stage entries the thread stores
    const int krow0 = lane < n ? lane : n * m, kstep = lane < n ? n : 1;
    double *dst[2];                      // null: nothing to store
    size_t dstride[2];
    int srow[2], sbb[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int e = tid + i * NT, row = e / WPB, bb = e % WPB;
        const bool ok = e < ROWS * WPB && b0 + bb < B && (!active || active[b0 + bb] != 0);
        srow[i] = ok ? row : 0;
        sbb[i] = bb;
        dst[i] = !ok ? nullptr : row < n * m ? K + (size_t)row * B + b0 + bb : k + (size_t)(row - n * m) * B + b0 + bb;
        dstride[i] = (size_t)(row < n * m ? n * m : m) * B;
    }

    // terminal condition (iLQR_class.py:136-138): V_xx = Q_f, V_x = Q_f (x_N - x_target)
    __syncwarp();
    if (lane < n) pt[P_XS + lane] = X[((size_t)N * n + lane) * B + b] - xt_l;
    __syncthreads();                                  // lxxS, RsS, dgS, maS, vflag, the patch
    double V[2][2][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = g + 8 * rt, j = 8 * ct + 2 * q + e;
                double v = 0.0;
                if (j < n) {
                    if (i < n) v = qc.Qfs[i][j];
                    else if (i == n) {
                        for (int l = 0; l < n; ++l) v += qc.Qfs[j][l] * pt[P_XS + l];
                    }
                }
                V[rt][ct][e] = v;
            }
    __syncwarp();
    if (lane < 16) pt[P_XS + 16 + lane] = fsrc[(size_t)(N - 1) * fstride] - xt_l;      // buffer of step t: (N - t) & 1
    __syncwarp();

    double wv = 0.0;
    for (int t = N - 1; t >= 0; --t) {
        const int buf = (N - t) & 1;
        const double *xs = pt + P_XS + buf * 16;
        double pre = 0.0;
        if (t > 0 && lane < 16) pre = fsrc[(size_t)(t - 1) * fstride] - xt_l;
        // w_t = amp sin(2 pi t / N + phi): lane L evaluates step t - L once per 32 steps, every step takes its value
        // from the lane that holds it (the same bits as evaluating it per step, 1/32 of the FP64 instructions)
        const int ti = (N - 1 - t) & 31;
        if (ti == 0) wv = sys.time_scalar(t - lane, ph);
        const double w = __shfl_sync(0xffffffffu, wv, ti);
        double MA[2][2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double2 be = maS[(mt * 2 + ct) * 2 + e][lane];
                    MA[mt][ct][e] = fma(w, be.y, be.x);
                }
        // ---- phase 1: Wt = (Vaug M)' ; Wt[mt][nt][e] = (Vaug M)[8 nt + 2q + e][g + 8 mt] ----
        double Wt[2][2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) Wt[mt][nt][0] = Wt[mt][nt][1] = 0.0;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) dmma884(Wt[mt][nt][0], Wt[mt][nt][1], MA[mt][ct][e], V[nt][ct][e]);
        // row 12 of Vaug M = (M' V_x)': Q_x = l_x + A_t' V_x, Q_u = l_u + B_t' V_x (:100-101), held by the q == 2 threads
        // (entry r = g + 8 mt).  It stays in the fragments: as contraction index 12 of the second product it meets the
        // zero padding of M'.  The same holds for whatever the padding rows and columns of Vaug collect from here on.
        {
            double l0, l1;
            if (qc.diag) {
                l0 = dgS[g] * xs[g];
                l1 = dgS[g + 8] * xs[g + 8];
            } else {
                l0 = l1 = 0.0;
                for (int j = 0; j < n; ++j) l0 += lxxS[g][j] * xs[j];
                if (g < 4) {
                    for (int j = 0; j < n; ++j) l1 += lxxS[g + 8][j] * xs[j];
                } else {
                    for (int j = 0; j < m; ++j) l1 += RsS[(g - 4) * 4 + j] * xs[n + j];
                }
            }
            if (q == 2) {
                pt[P_QV + g] = l0 + Wt[0][1][0];
                pt[P_QV + g + 8] = l1 + Wt[1][1][0];
            }
        }
        // ---- phase 2: G = M' W ; G[mr][mt][e] = (M' V_xx M)[g + 8 mr][8 mt + 2q + e] ----
        double G[2][2][2];
#pragma unroll
        for (int mr = 0; mr < 2; ++mr)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) G[mr][mt][0] = G[mr][mt][1] = 0.0;
        // the tiles holding rows 8..15 first: their rows 12..15 feed the solve, whose shared-memory round trip then runs
        // under the DMMAs of the other two tiles
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma884(G[1][mt][0], G[1][mt][1], MA[1][nt][e], Wt[mt][nt][e]);
        // ---- phase 3: rows 12..15 of G (threads g >= 4, u = g - 4) are Q_ux[u][:] and Q_uu[u][:] - l_uu (:103-104) ----
        if (g >= 4) {
            double *col = pt + P_TAB + (g - 4);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                col[(2 * q + e) * 4] = G[1][0][e];                                   // columns 0..7 (l_ux = 0)
                col[(8 + 2 * q + e) * 4] = (G[1][1][e] + radd[e]) + mudd[e];         // columns 8..11, Q_uu beyond
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma884(G[0][mt][0], G[0][mt][1], MA[0][nt][e], Wt[mt][nt][e]);
        __syncwarp();
        double Lm[m][m], rhs[m];
#pragma unroll
        for (int v = 0; v < m; ++v) {
            const double2 r0 = *reinterpret_cast<const double2 *>(pt + P_TAB + (n + v) * 4);
            const double2 r1 = *reinterpret_cast<const double2 *>(pt + P_TAB + (n + v) * 4 + 2);
            Lm[0][v] = r0.x; Lm[1][v] = r0.y; Lm[2][v] = r1.x; Lm[3][v] = r1.y;
        }
        {
            const double2 r0 = *reinterpret_cast<const double2 *>(rhs_src), r1 = *reinterpret_cast<const double2 *>(rhs_src + 2);
            rhs[0] = r0.x; rhs[1] = r0.y; rhs[2] = r1.x; rhs[3] = r1.y;
        }
        if (__any_sync(0xffffffffu, lu4_solve_nopivot(Lm, rhs))) {      // rare: a pivot was not the largest of its column
            double Lp[m][m], rp[m][1];
#pragma unroll
            for (int u = 0; u < m; ++u) {
#pragma unroll
                for (int v = 0; v < m; ++v) Lp[u][v] = pt[P_TAB + (n + v) * 4 + u];
                rp[u][0] = rhs_src[u];
            }
            lu_solve_inplace<m, 1, double, true>(Lp, rp);
#pragma unroll
            for (int u = 0; u < m; ++u) rhs[u] = rp[u][0];
        }
        double (*ks)[KROW] = ksS[buf];
#pragma unroll
        for (int u = 0; u < m; ++u) {
            const double kv = -rhs[u];
            if (lane < n) pt[P_KW + u * KWS + lane] = kv;
            if (lane <= n) ks[krow0 + u * kstep][wp] = kv;
        }
        __syncwarp();
        // ---- phase 4: Vaug' = [Q_xx ; Q_x'] + [Q_ux' ; Q_u'] K (:113-114), accumulated on top of G ----
        {
            const double ap0 = pt[P_TAB + g * 4 + q], ap1 = *ap1_src;                // Q_ux[q][g], Q_ux[q][8 + g] / Q_u[q]
            const double bp0 = pt[P_KW + q * KWS + g], bp1 = pt[P_KW + q * KWS + g + 8];
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const double2 lq0 = *reinterpret_cast<const double2 *>(&lxxS[g][8 * ct + 2 * q]);
                const double2 lq1 = *reinterpret_cast<const double2 *>(&lxxS[g + 8][8 * ct + 2 * q]);
                const double2 qx = *reinterpret_cast<const double2 *>(pt + P_QV + 8 * ct + 2 * q);
                double c0 = G[0][ct][0] + lq0.x, c1 = G[0][ct][1] + lq0.y;
                dmma884(c0, c1, ap0, ct == 0 ? bp0 : bp1);
                V[0][ct][0] = c0;
                V[0][ct][1] = c1;
                c0 = g == 4 ? qx.x : G[1][ct][0] + lq1.x;                            // row 12: V_x' = Q_x' + Q_u' K
                c1 = g == 4 ? qx.y : G[1][ct][1] + lq1.y;
                dmma884(c0, c1, ap1, ct == 0 ? bp0 : bp1);
                V[1][ct][0] = c0;
                V[1][ct][1] = c1;
            }
        }
        if (t > 0 && lane < 16) pt[P_XS + (buf ^ 1) * 16 + lane] = pre;
        __syncthreads();
        // coalesced store of the step's gains: rows of WPB consecutive trajectories
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (dst[i]) dst[i][(size_t)t * dstride[i]] = ks[srow[i]][sbb[i]];
    }
}

}  // namespace ilqr
