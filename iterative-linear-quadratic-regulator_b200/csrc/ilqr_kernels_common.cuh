// ilqr_kernels_common.cuh -- structures shared by the kernels: step-size list, iteration control block, speculation and
// regularisation arguments
// Part of libilqr_b200.so; included by ilqr_b200.cu only (see the file map at its top).
#pragma once
#include "ilqr_systems.cuh"
#include "ilqr_b200.h"

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
// Two tiers (round 2): a trajectory is listed at once when its accepted try index w, or w plus the index it accepted
// the iteration before, is high (tier 1); the milder cases (tier 2) are collected apart and fill whatever capacity
// tier 1 leaves, merged by the last block of select_kernel to finish.  On config 2 (offline, oracle traces of eleven
// shards of 4096): single threshold 5 missed second waves per 11 solves, two tiers 1.
struct SpecArgs {
    int cap;                              // list capacity; 0 switches speculation off
    int n2;                               // deferred step sizes per trajectory
    int threshold;                        // tier 1: w >= threshold or w + previous w >= threshold + 3; tier 2: sum >= threshold + 2
    int *list_cur, *list_next;            // [cap]
    unsigned int *count_cur, *count_next; // entries appended this / next iteration (may exceed cap)
    int *mark;                            // [B]: mark[b] == it + 1 <=> b is on the list of iteration it
    int *hist;                            // [B]: try index accepted in the previous iteration
    int *list2_next;                      // [cap] tier-2 candidates for the next iteration
    unsigned int *count2_next;            // tier-2 candidates appended (may exceed cap)
    unsigned int *ticket;                 // blocks of this iteration's first select that have finished
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

// Blocked layout of the linearization inside ilqr_solve's workspace: AB[t][group of 32 columns][row][lane], rows =
// the n*n entries of A_t followed by the n*m entries of B_t.  Everything a warp reads or writes for one timestep
// (n(n+m) rows of 32 values) is ONE contiguous chunk (5 KB for n=4, m=1) instead of n(n+m) pieces a whole batch
// row apart -- DRAM page locality for K1's stores and K2's streaming loads.  The public entry points
// (ilqr_linearize, ilqr_backward) keep the documented [N][n][n][B] / [N][n][m][B] arrays.
ILQR_DEV size_t ab_off(int rows, int t, int row, int col, int B)
{
    const size_t groups = ((size_t)B + 31) >> 5;
    return ((((size_t)t * groups + ((size_t)col >> 5)) * rows + row) << 5) + (col & 31);
}

// Solver mode (tol > 0): trajectories converge at different iterations, and the late iterations of a solve
// serve a small, scattered fraction of the batch.  The select kernels therefore keep a compacted list of the
// trajectories entering the next iteration (its length is the n_active counter they maintain anyway), and
// when that count drops to `thresh` or below, K1, K2 and the rollouts index the batch through the list instead
// of scanning it: the cost of an iteration follows the number of active trajectories.  Dense or sparse is
// decided on the device, per iteration, from the same counter in every kernel.  cur == nullptr disables.
struct SparseArgs {
    const int *cur;               // list of the trajectories active in this iteration     [*n_cur entries]
    const unsigned int *n_cur;    // = &ctl->n_active[it]
    const int *prev;              // list of the previous iteration (K1: pending commits)    [*n_prev entries]
    const unsigned int *n_prev;   // = &ctl->n_active[it - 1]; nullptr in iteration 0
    int *next;                    // list being built for the next iteration (select kernels)
    int *pos;                     // [B] position of a trajectory in the list it was last put on: in a sparse iteration
                                  // K1 writes A_t, B_t at that column and K2 reads them back coalesced
    int only;                     // K2 comes in two kernels at large batches: 1 = run in dense iterations only,
                                  // 2 = in sparse iterations only, 0 = always
    unsigned int thresh;          // at or below this many active trajectories the kernels walk the list
    unsigned int thresh_all;      // at or below this many the first rollout wave tries EVERY step size (the wave
                                  // is latency bound by then: extra rollouts are free, extra waves are not)
    int n_alpha_all;
};

ILQR_DEV bool sparse_now(const SparseArgs &sa) { return sa.cur != nullptr && *sa.n_cur <= sa.thresh; }
ILQR_DEV bool sparse_all(const SparseArgs &sa) { return sa.cur != nullptr && *sa.n_cur <= sa.thresh_all; }
ILQR_DEV bool sparse_prev(const SparseArgs &sa) { return sa.prev != nullptr && sa.n_prev != nullptr && *sa.n_prev <= sa.thresh; }

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

// ---- mbarrier (shared::cta) -------------------------------------------------------------------------------------------
ILQR_DEV void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// release: the caller's earlier shared-memory writes are visible to whoever observes the phase complete
ILQR_DEV void mbar_arrive(unsigned long long *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// acquire: spin until the phase of the given parity has completed (a fresh barrier passes parity 1 at once)
ILQR_DEV void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(a),
                 "r"(parity)
                 : "memory");
}

// ---- bulk asynchronous copies (the TMA unit's linear form) ---------------------------------------------------------
// the arriving thread announces how many bytes the copies it is about to issue will deliver to the barrier's phase
ILQR_DEV void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
// one contiguous chunk global -> shared (16-byte aligned on both sides, size a multiple of 16); completion is counted in
// bytes on `bar`
ILQR_DEV void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                 "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

}  // namespace ilqr
