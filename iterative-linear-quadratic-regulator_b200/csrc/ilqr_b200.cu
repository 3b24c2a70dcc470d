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
//
// File map (csrc/):
//   ilqr_systems.cuh            device models, integrators, analytic step Jacobians, quadratic cost
//   ilqr_kernels_common.cuh     structures shared by the kernels (control block, speculation / regularisation args)
//   ilqr_kernels_linearize.cuh  step_kernel, K1, materialised cost expansion, MPC shift
//   ilqr_kernels_backward.cuh   K2 in its three forms
//   ilqr_kernels_ltv_mma.cuh    K2 of the LTV model on the FP64 tensor cores (DMMA), one warp per trajectory
//   ilqr_kernels_fused.cuh      K1 + K2 as one warp-specialised kernel (producers linearize, consumer scans)
//   ilqr_kernels_rollout.cuh    K3 and K4 (eager and lazy line-search schedules)
//   ilqr_b200.cu (this file)    handle, workspace layout, launch configuration, ilqr_solve, the C ABI
#include "ilqr_b200.h"
#include "ilqr_systems.cuh"
// A user-defined System subclass (model ILQR_USER) brings its own device code: class_files/codegen.py generates
// ilqr::UserSys<T> / ilqr::UserCost<T>, NVRTC compiles the generic kernel templates of csrc/*.cuh against them in
// process, and this library loads the cubin (ilqr_module_load) and launches those kernels by handle (UserModule).

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <type_traits>
#include <vector>

#include "ilqr_kernels_common.cuh"
#include "ilqr_kernels_linearize.cuh"
#include "ilqr_kernels_backward.cuh"
#include "ilqr_kernels_ltv_mma.cuh"
#include "ilqr_kernels_fused.cuh"
#include "ilqr_kernels_rollout.cuh"

namespace ilqr {

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

template <typename T, int M, bool TAB = false> DoublePendulumSys<T, M, TAB> make_double(const ilqr_problem_t &p)
{
    const double g = p.phys[0], m1 = p.phys[1], m2 = p.phys[2], l1 = p.phys[3], l2 = p.phys[4];
    const double d1 = p.phys[5], d2 = p.phys[6], th1 = p.phys[7], th2 = p.phys[8];
    DoublePendulumSys<T, M, TAB> s;
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

// Kernels of a user-defined system: a cubin built by NVRTC (class_files/codegen.py) from the generic templates of
// csrc/*.cuh and the generated ilqr::UserSys / ilqr::UserCost, loaded through the runtime's library API.  The kernels
// are the SAME templates this file instantiates for the shipped models, so their parameter lists are marshalled here
// exactly as the <<< >>> launches below pass them.
struct UserModule {
    cudaLibrary_t lib;
    cudaKernel_t k[ILQR_N_USER_KERNELS];
    int n, m, integrator, dtype;
};
// host mirrors of the generated parameter structs (UserSys<T> is empty; UserCost<T> = { T dt; int diag, monotone; })
struct UserSysArg { char unused; };
template <typename T> struct UserCostArg { T dt; int diag, monotone; };

struct Handle {
    ilqr_problem_t p;
    UserModule *umod;         // model ILQR_USER: where the kernels live (not owned)
    int n_alpha_eff;          // tries actually made: stops once alpha < min_alpha (iLQR_class.py:300-302)
    int n_first;              // step sizes rolled out eagerly (first wave); the rest only where needed
    int spec_cap;             // trajectories whose deferred step sizes ride along speculatively (SpecArgs)
    // tuning overrides read once from the environment in ilqr_create (exploration; defaults in brackets):
    int env_lanes;            // ILQR_BACKWARD_LANES: -1 [auto: lanes kernel for B <= 32768], 0, 1
    int env_rollout_bs;       // ILQR_ROLLOUT_BS: 0 [auto] or a block size
    int env_fused;            // ILQR_FUSED: -1 [auto: fused K1+K2 inside ilqr_solve where it is the faster form], 0 never,
                              // 1 wherever the model allows (also ilqr_backward_pass)
    int env_fused_minb;       // ILQR_FUSED_MINB: 0 [auto: by batch size], 1 = uncapped registers, 4 / 5 = capped for 4 / 5 blocks per SM
    int env_ltv_lanes;        // ILQR_LTV_LANES: 0 [auto: by batch size and precision], 4 (register-tiled kernel), 16 (one column
                              // per lane) or 32 (FP64 tensor-core kernel, one warp per trajectory)
    int env_ltv_wpb;          // ILQR_LTV_MMA_WPB: trajectories per block of the tensor-core kernel (4 [default] or 8)
    int env_check_every;      // ILQR_CHECK_EVERY: iterations enqueued between host polls of the active count [8]
    long env_sparse_thresh;   // ILQR_SPARSE_THRESH / ILQR_SPARSE_ALL: -1 [auto] or the thresholds of SparseArgs
    long env_sparse_all;
    // opt-in to > 48 KB of dynamic shared memory is a per-device, per-kernel attribute: tracked per handle (a handle
    // lives on one device), never in function statics
    size_t smem_backward;     // largest size configured for this handle's backward_kernel instantiation
    int smem_ltv;             // backward_ltv_kernel configured
    int smem_fsplit;          // fused_backward_split_kernel configured (dynamic shared memory above 48 KB)
    int trig_table;           // the double pendulums' FP64 kernels take sin/cos from the shared-memory table: batches of
                              // ILQR_TRIG_TABLE_MIN [8192] trajectories or more (pipe-bound kernels)
    int env_fused_split;      // ILQR_FUSED_SPLIT: -1 [auto: at most two blocks per SM, n = 4], 0 never, 1 whenever n = 4
    void *mu_user;            // optional caller buffer for the per-trajectory regularisation (ilqr_set_mu_buffer)
    int env_bulk;             // ILQR_BACKWARD_BULK: bulk-copy ring of the thread-per-trajectory K2 (-1: by batch size)
    int ab_blocked;           // ilqr_solve stores the linearization blocked by groups of 32 trajectories (ab_off)
    int sparse;               // lazy schedule: late iterations index the batch through the active list (SparseArgs)
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

// does this handle's ilqr_solve / ilqr_backward_pass materialise A_t, B_t in the workspace?  (see ws_layout)
static bool stores_linearization(const Handle *h)
{
    return h->p.model != ILQR_LTV && (h->p.model == ILQR_USER || h->env_fused == 0);
}

static inline int grid_for(size_t threads, int bs) { return (int)((threads + bs - 1) / bs); }

// pick a block size that still spreads small batches over all 148 SMs
static inline int block_for(size_t threads)
{
    int bs = 256;
    while (bs > 32 && (threads + bs - 1) / bs < 16 * 148) bs >>= 1;   // small batches: one warp per block balances best
    return bs;
}

struct WsLayout {
    size_t ctl, A, Bd, Xc, Uc, cost_alpha, winner, wslot, active, defer, mark, hist, lists, lists2, alists, apos, mu, total;
};

static size_t ctl_bytes(int maxiter)
{
    return sizeof(Control) + sizeof(unsigned int) * (5 + ILQR_MAX_WAVES) * (size_t)(maxiter + 2);
}

static WsLayout ws_layout(const ilqr_problem_t &p, int n_alpha, bool store_linearization)
{
    const size_t w = p.dtype == ILQR_F64 ? 8 : 4;
    const size_t B = p.B, N = p.N, n = p.n, m = p.m;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    WsLayout L;
    size_t off = 0;
    // n_active[maxiter + 2], then the per-iteration deferred (second-wave) counters [maxiter + 2], then the
    // per-iteration speculation-list counters [maxiter + 2], then the lazy-wave list counters
    // [maxiter + 2][ILQR_MAX_WAVES], then the tier-2 speculation counters and the select tickets [maxiter + 2] each
    L.ctl = off; off = al(off + ctl_bytes(p.maxiter));
    // A_t, B_t are stored only where K1 and K2 run as two kernels (user-defined models, ILQR_FUSED=0): the fused kernel
    // hands them over in shared memory and the LTV model generates them inside its kernels.  Config 5's shard
    // (B=131072, N=500): 36.8 -> 26.3 GB of workspace
    const size_t lin = store_linearization ? 1 : 0;
    // sized for the batch padded to whole groups of 32 columns: ilqr_solve keeps A and Bd as ONE blocked array
    // starting at L.A (ab_off); the two sizes are multiples of 256 bytes, so the regions are contiguous
    const size_t Bpad = (B + 31) / 32 * 32;
    L.A = off; off = al(off + lin * w * N * n * n * Bpad);
    L.Bd = off; off = al(off + lin * w * N * n * m * Bpad);
    L.Xc = off; off = al(off + w * (size_t)n_alpha * (N + 1) * n * B);
    L.Uc = off; off = al(off + w * (size_t)n_alpha * N * m * B);
    L.cost_alpha = off; off = al(off + w * (size_t)n_alpha * B);
    L.winner = off; off = al(off + 4 * B);
    L.wslot = off; off = al(off + 4 * B);
    L.active = off; off = al(off + 4 * B);
    L.defer = off; off = al(off + 4 * B);
    L.mark = off; off = al(off + 4 * B);
    L.hist = off; off = al(off + 4 * B);          // try index accepted in the previous iteration (SpecArgs)
    L.lists = off; off = al(off + 4 * 2 * B);     // two speculation lists (capacity <= B each)
    L.lists2 = off; off = al(off + 4 * B);        // tier-2 candidates of the next iteration
    L.alists = off; off = al(off + 4 * 2 * B);    // active lists of the current / next iteration (SparseArgs)
    L.apos = off; off = al(off + 4 * B);          // position of each trajectory in its active list
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

// launch kernel `which` of the handle's user module; args = addresses of the kernel's parameters in order
static int launch_user(Handle *h, int which, int grid, int block, size_t smem, cudaStream_t st, void **args)
{
    cudaError_t e = cudaLaunchKernel((const void *)h->umod->k[which], dim3((unsigned)grid), dim3((unsigned)block), args, smem, st);
    h->launches++;
    if (e != cudaSuccess) { h->last_cuda = (int)e; cudaGetLastError(); return ILQR_E_CUDA; }
    return ILQR_OK;
}

// the generated cost's parameter block in the handle's element type (UserCost<T> = { T dt; int diag, monotone; })
struct UserCostAny {
    UserCostArg<double> d;
    UserCostArg<float> f;
    void *ptr(const Handle *h) { return h->p.dtype == ILQR_F64 ? (void *)&d : (void *)&f; }
    explicit UserCostAny(const Handle *h) : d{h->p.dt, 0, 0}, f{(float)h->p.dt, 0, 0} {}
};
struct ScalarAny {           // a `T` kernel parameter
    double d;
    float f;
    void *ptr(const Handle *h) { return h->p.dtype == ILQR_F64 ? (void *)&d : (void *)&f; }
    explicit ScalarAny(double v) : d(v), f((float)v) {}
};

// dispatch on (dtype, model, integrator): calls f(T{}, sys, qc, integral_constant<int,INTEG>{})
template <typename T, class Sys, class F> static int dispatch_integ(const Handle *h, const Sys &sys, F &&f)
{
    auto qc = make_cost<T, Sys::N, Sys::M>(h->p);
#ifdef ILQR_FAST_BUILD      // kernel experiments (scripts/): UA double pendulum, rk4, FP64 only -- compiles in seconds
    if (h->p.integrator == ILQR_RK4) return f(T(0), sys, qc, std::integral_constant<int, RK4>{});
    return ILQR_E_INVALID;
#else
    switch (h->p.integrator) {
    case ILQR_EULER: return f(T(0), sys, qc, std::integral_constant<int, EULER>{});
    case ILQR_MIDPOINT: return f(T(0), sys, qc, std::integral_constant<int, MIDPOINT>{});
    case ILQR_RK4: return f(T(0), sys, qc, std::integral_constant<int, RK4>{});
    case ILQR_BACKWARD_EULER: return f(T(0), sys, qc, std::integral_constant<int, BACKWARD_EULER>{});
    }
    return ILQR_E_INVALID;
#endif
}

template <typename T, class F> static int dispatch_model(const Handle *h, F &&f)
{
#if defined(ILQR_FAST_BUILD) && ILQR_FAST_BUILD == 2     // LTV model only
    if (h->p.model == ILQR_LTV) {
        auto sys = make_ltv<T>(h->p);
        auto qc = make_cost<T, 12, 4>(h->p);
        return f(T(0), sys, qc, std::integral_constant<int, EULER>{});
    }
    return ILQR_E_INVALID;
#elif defined(ILQR_FAST_BUILD)
    if (h->p.model == ILQR_UA_DOUBLE_PENDULUM) {
        if constexpr (std::is_same_v<T, double>) {
            if (h->trig_table) return dispatch_integ<T>(h, make_double<T, 1, true>(h->p), f);
        }
        return dispatch_integ<T>(h, make_double<T, 1>(h->p), f);
    }
    return ILQR_E_INVALID;
#else
    switch (h->p.model) {
    case ILQR_PENDULUM: return dispatch_integ<T>(h, make_pendulum<T>(h->p), f);
    // large FP64 batches: the double pendulums' sines and cosines through the shared-memory table (a separate model type)
    case ILQR_DOUBLE_PENDULUM:
        if constexpr (std::is_same_v<T, double>) {
            if (h->trig_table) return dispatch_integ<T>(h, make_double<T, 2, true>(h->p), f);
        }
        return dispatch_integ<T>(h, make_double<T, 2>(h->p), f);
    case ILQR_UA_DOUBLE_PENDULUM:
        if constexpr (std::is_same_v<T, double>) {
            if (h->trig_table) return dispatch_integ<T>(h, make_double<T, 1, true>(h->p), f);
        }
        return dispatch_integ<T>(h, make_double<T, 1>(h->p), f);
    case ILQR_LTV: {   // forward Euler only (validated in ilqr_create)
        auto sys = make_ltv<T>(h->p);
        auto qc = make_cost<T, 12, 4>(h->p);
        return f(T(0), sys, qc, std::integral_constant<int, EULER>{});
    }
    }
    return ILQR_E_INVALID;
#endif
}

template <class F> static int dispatch(const Handle *h, F &&f)
{
    if (h->p.dtype == ILQR_F64) return dispatch_model<double>(h, f);
#ifdef ILQR_FAST_BUILD
    return ILQR_E_INVALID;
#else
    return dispatch_model<float>(h, f);
#endif
}

// ---- launch helpers -----------------------------------------------------------------------

static int launch_commit_linearize(Handle *h, const void *phi, void *X, void *U, void *A, void *Bd, const void *Xc, const void *Uc,
                                   const int *winner, const int *wslot, const int *active, int do_lin, const unsigned int *g0,
                                   const unsigned int *g1, cudaStream_t st, const int *iters = nullptr, int it = 0,
                                   const SparseArgs *sparse = nullptr, int ab_blocked = 0, int sparse_only = 0)
{
    SparseArgs sa;
    std::memset(&sa, 0, sizeof sa);
    if (sparse) sa = *sparse;
    if (h->umod) {
        UserSysArg us{};
        ScalarAny dt(h->p.dt);
        int N = h->p.N, B = h->p.B;
        void *args[] = { &us, dt.ptr(h), &N, &B, &phi, &X, &U, &A, &Bd, &Xc, &Uc, &winner, &wslot, &active, &iters, &it, &do_lin,
                         &g0, &g1, &sa, &ab_blocked, &sparse_only };
        int grid = grid_for((size_t)(N + 1) * B, 128);
        if (sparse_only && grid > 148 * 16) grid = 148 * 16;
        return launch_user(h, ILQR_UK_LINEARIZE, grid, 128, 0, st, args);
    }
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto integ) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        constexpr int I = decltype(integ)::value;
        const size_t threads = (size_t)(h->p.N + 1) * h->p.B;
        const int bs = 128;
        int grid = grid_for(threads, bs);
        // next to the fused kernel K1 only serves sparse iterations: a capped grid that strides over its items, so
        // that the launch costs nothing in the dense ones (513k empty blocks took 0.28 ms at B=131072)
        if (sparse_only && grid > 148 * 16) grid = 148 * 16;
        commit_linearize_kernel<Sys, I, T><<<grid, bs, 0, st>>>(
            sys, qc.dt, h->p.N, h->p.B, (const T *)phi, (T *)X, (T *)U, (T *)A, (T *)Bd, (const T *)Xc, (const T *)Uc, winner, wslot,
            active, iters, it, do_lin, g0, g1, sa, ab_blocked, sparse_only);
        ILQR_CHECK_LAUNCH(h);
        return ILQR_OK;
    });
}

template <typename T, int n, int m, int DEPTH, int bs, class Cost>
static int launch_backward_depth(Handle *h, const Cost &qc, const void *X, const void *U,
                                 const void *A, const void *Bd, void *K, void *k, const int *active,
                                 const unsigned int *gate, const void *mu, cudaStream_t st, const SparseArgs &sa,
                                 int ab_blocked)
{
    constexpr int L = n * n + n * m + n + m;
    const size_t smem = (size_t)DEPTH * L * bs * sizeof(T) + (size_t)DEPTH * (bs / 32) * 8;     // ring + its mbarriers
    if (smem > h->smem_backward) {          // a handle uses ONE instantiation (fixed model, dtype, batch)
        cudaError_t e = cudaFuncSetAttribute(backward_kernel<Cost, T, n, m, DEPTH, bs>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
        h->smem_backward = smem;
    }
    backward_kernel<Cost, T, n, m, DEPTH, bs><<<grid_for(h->p.B, bs), bs, smem, st>>>(
        qc, h->p.N, h->p.B, (const T *)X, (const T *)U, (const T *)A, (const T *)Bd, (T *)K, (T *)k, active, gate,
        (const T *)mu, sa, ab_blocked);
    ILQR_CHECK_LAUNCH(h);
    return ILQR_OK;
}

static int launch_backward(Handle *h, const void *X, const void *U, const void *A, const void *Bd, void *K, void *k,
                           const int *active, const unsigned int *gate, cudaStream_t st, const void *mu = nullptr,
                           const SparseArgs *sparse = nullptr, int ab_blocked = 0)
{
    SparseArgs sa;
    std::memset(&sa, 0, sizeof sa);
    if (sparse) sa = *sparse;
    // thread-per-trajectory kernel: with the blocked linearization and whole warps of trajectories a warp's step arrives
    // by bulk copies (ab_blocked = 2, see backward_kernel); from 16384 trajectories up by default (measured on the generic
    // kernel of a user system: 0.188 -> 0.162 ms per pass at B=16384, N=200; 0.137 -> 0.161 at 8192), ILQR_BACKWARD_BULK=0/1
    const int ab_lanes = ab_blocked;
    if (ab_blocked && h->p.B % 32 == 0 && (h->env_bulk >= 0 ? h->env_bulk != 0 : h->p.B >= 16384)) ab_blocked = 2;
    if (h->umod) {
        // the generic thread-per-trajectory scan, ring depth / block size by state dimension and batch as below
        const int n = h->p.n, m = h->p.m, L = n * n + n * m + n + m;
        const bool large = n <= 4 && h->p.B > 32768;
        const int depth = n > 4 ? 2 : (large ? 4 : 8), bs = large ? 64 : 32;
        const int which = large ? ILQR_UK_BACKWARD_LARGE : ILQR_UK_BACKWARD_SMALL;
        const size_t smem = (size_t)depth * L * bs * (h->p.dtype == ILQR_F64 ? 8 : 4) + (size_t)depth * (bs / 32) * 8;
        if (smem > h->smem_backward) {
            cudaError_t e = cudaFuncSetAttribute((const void *)h->umod->k[which], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
            h->smem_backward = smem;
        }
        UserCostAny qc(h);
        int N = h->p.N, B = h->p.B;
        void *args[] = { qc.ptr(h), &N, &B, &X, &U, &A, &Bd, &K, &k, &active, &gate, &mu, &sa, &ab_blocked };
        return launch_user(h, which, grid_for(B, bs), bs, smem, st, args);
    }
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        // small batches: one warp per block and a deep ring (latency bound); large batches: shallower
        // ring so that more warps fit per SM (HBM bound)
        if constexpr (Sys::N == 4 && Sys::M == 1 && decltype(qc)::QUADRATIC) {
            // four lanes per trajectory: small batches (latency bound regime), and the sparse iterations of large
            // ones, where it runs next to the thread-per-trajectory kernel and each returns at once when the
            // iteration is not its kind (SparseArgs::only)
            const bool lanes = h->env_lanes >= 0 ? h->env_lanes != 0 : h->p.B <= 32768;
            const bool both = !lanes && sa.cur != nullptr && h->env_lanes < 0;
            if (lanes || both) {
                constexpr int DEPTH = 8, SLOTS = 8, LP = 26;
                const size_t smem = sizeof(T) * (size_t)(DEPTH * SLOTS * LP + SLOTS * 4 + SLOTS * 20);
                SparseArgs sl = sa;
                sl.only = both ? 2 : 0;
                const int items = both ? (int)sa.thresh : h->p.B;
                backward_n4m1_lanes_kernel<T, DEPTH><<<grid_for(items, SLOTS), 32, smem, st>>>(
                    qc, h->p.N, h->p.B, (const T *)X, (const T *)U, (const T *)A, (const T *)Bd, (T *)K, (T *)k, active,
                    gate, (const T *)mu, sl, ab_lanes);
                ILQR_CHECK_LAUNCH(h);
                if (!both) return ILQR_OK;
                sa.only = 1;
            }
        }
        if constexpr (Sys::N > 4) {
            // n = 12, m = 4: a ring stage is 208 rows; two stages of one warp fit the 227 KB limit
            return launch_backward_depth<T, Sys::N, Sys::M, 2, 32>(h, qc, X, U, A, Bd, K, k, active, gate, mu, st, sa, ab_blocked);
        } else {
            if (h->p.B <= 32768)
                return launch_backward_depth<T, Sys::N, Sys::M, 8, 32>(h, qc, X, U, A, Bd, K, k, active, gate, mu, st, sa, ab_blocked);
            return launch_backward_depth<T, Sys::N, Sys::M, 4, 64>(h, qc, X, U, A, Bd, K, k, active, gate, mu, st, sa, ab_blocked);
        }
    });
}

// The fused K1+K2 kernel exists for the hand-written second-order models with the quadratic cost (n <= 4).
template <class Sys, class Cost> constexpr bool fused_eligible()
{
    return !Sys::GENERIC && !Sys::FIRST_ORDER && Sys::N <= 4 && Cost::QUADRATIC;
}

static bool fused_available(const Handle *h)
{
    return h->p.model != ILQR_LTV && h->p.model != ILQR_USER && h->env_fused != 0;
}

// K1 + K2 in one launch (ilqr_kernels_fused.cuh): commit of the accepted candidates, linearization and reverse scan
static int launch_fused(Handle *h, const void *phi, void *X, void *U, const void *Xc, const void *Uc, const int *winner,
                        const int *wslot, const int *active, const int *iters, int it, const unsigned int *g0,
                        const unsigned int *g1, void *K, void *k, const void *mu, cudaStream_t st,
                        const SparseArgs *sparse = nullptr)
{
    SparseArgs sa;
    std::memset(&sa, 0, sizeof sa);
    if (sparse) sa = *sparse;
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto integ) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        using Cost = decltype(qc);
        constexpr int I = decltype(integ)::value;
        if constexpr (fused_eligible<Sys, Cost>()) {
            const int groups = (h->p.B + 31) / 32;
            auto go = [&](auto np, auto stages, auto minb) {
                constexpr int NP = decltype(np)::value, S = decltype(stages)::value, MB = decltype(minb)::value;
                fused_backward_kernel<Sys, Cost, I, T, NP, S, MB><<<groups, 32 * (NP + 1), 0, st>>>(
                    sys, qc, h->p.N, h->p.B, (const T *)phi, (T *)X, (T *)U, (const T *)Xc, (const T *)Uc, winner, wslot, active,
                    iters, it, g0, g1, (T *)K, (T *)k, (const T *)mu, sa);
            };
            // Small batches (at most two rounds of one block per SM): two consumer warps, each half the columns, four
            // producers (ms per pass, one consumer / split: B=1024 0.273 / 0.226, B=4096 0.272 / 0.226, B=8192 0.474 / 0.448;
            // profiles/r02_exp_fused_split.log)
            if constexpr (Sys::N == 4) {
                if (h->env_fused_split != 0 && (h->env_fused_split == 1 || groups <= 2 * 148)) {
                    constexpr int NP = 4, S = 4;
                    constexpr size_t smem = fused_split_smem_bytes<Sys, Cost, NP, S, T>();
                    auto kern = fused_backward_split_kernel<Sys, Cost, I, T, NP, S>;
                    if (!h->smem_fsplit) {
                        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                        if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
                        h->smem_fsplit = 1;
                    }
                    kern<<<groups, 32 * (NP + 2), smem, st>>>(sys, qc, h->p.N, h->p.B, (const T *)phi, (T *)X, (T *)U,
                                                              (const T *)Xc, (const T *)Uc, winner, wslot, active, iters, it, g0,
                                                              g1, (T *)K, (T *)k, (const T *)mu, sa);
                    ILQR_CHECK_LAUNCH(h);
                    return ILQR_OK;
                }
            }
            using std::integral_constant;
            using I1 = integral_constant<int, 1>;
            using I2 = integral_constant<int, 2>;
            using I4 = integral_constant<int, 4>;
            using I5 = integral_constant<int, 5>;
            // Two producer warps per consumer (three or four gain nothing: at small batches the kernel runs at the speed
            // of the consumer's dependent chain, 0.273 ms at B=4096 = the thread-per-trajectory scan alone, with the
            // linearization hidden completely; 4 producers on single ring stages: 0.352 ms).  The register cap follows
            // the batch so that every block is resident at once (one block = 32 trajectories = 96 threads):
            //   <= 3 blocks per SM: uncapped (194 registers)            B=8192: 0.475 vs 0.490 ms capped for 4
            //   <= 4 blocks per SM: capped for 4 (164, no spills)       B=16384: 0.861 vs 0.925 ms uncapped
            //   more: capped for 5 (128 registers, 15 warps per SM)     B=131072: 5.37 vs 5.50 (4) vs 6.16 ms (uncapped)
            // ILQR_FUSED_MINB overrides (experiments).
            int mb = h->env_fused_minb;
            if (mb == 0) mb = groups <= 148 * 3 ? 1 : (groups <= 148 * 4 ? 4 : 5);
            if (mb >= 5) go(I2{}, I4{}, I5{});
            else if (mb == 4) go(I2{}, I4{}, I4{});
            else go(I2{}, I4{}, I1{});
            ILQR_CHECK_LAUNCH(h);
            return ILQR_OK;
        } else {
            return ILQR_E_INVALID;
        }
    });
}

// K2 of the LTV model: A_t, B_t generated in the kernel (no linearization buffers)
#ifndef ILQR_LTV_MMA_MIN_BATCH
#define ILQR_LTV_MMA_MIN_BATCH 1
#endif
static int launch_backward_ltv(Handle *h, const void *phi, const void *X, const void *U, void *K, void *k,
                               const int *active, const unsigned int *gate, cudaStream_t st, const void *mu = nullptr)
{
#if defined(ILQR_FAST_BUILD) && ILQR_FAST_BUILD != 2
    return ILQR_E_INVALID;
#else
    constexpr int TPB = 16;
    auto go = [&](auto tz) -> int {
        using T = decltype(tz);
        // FP64: the tensor-core kernel (ilqr_kernels_ltv_mma.cuh), one warp per trajectory, at every batch size (N=1000,
        // ms per pass, sixteen-lane / four-lane / tensor-core kernel: B=256 3.94 / 6.90 / 1.54, B=1024 4.00 / 7.16 / 1.92,
        // B=4096 6.85 / 7.60 / 5.15, B=32768 47.1 / 39.6 / 33.0; profiles/r02_exp_ltv_sizes.log)
        if constexpr (std::is_same_v<T, double>) {
            if (h->env_ltv_lanes ? h->env_ltv_lanes == 32 : h->p.B >= ILQR_LTV_MMA_MIN_BATCH) {
                auto run = [&](auto wz) -> int {
                    constexpr int WPB = decltype(wz)::value;
                    backward_ltv_mma_kernel<WPB><<<grid_for(h->p.B, WPB), WPB * 32, 0, st>>>(
                        make_ltv<double>(h->p), make_cost<double, 12, 4>(h->p), h->p.N, h->p.B, (const double *)phi,
                        (const double *)X, (const double *)U, (double *)K, (double *)k, active, gate, (const double *)mu);
                    ILQR_CHECK_LAUNCH(h);
                    return ILQR_OK;
                };
                if (h->env_ltv_wpb == 8) return run(std::integral_constant<int, 8>());
                return run(std::integral_constant<int, 4>());
            }
        }
        // FP32 (and ILQR_LTV_LANES): sixteen lanes per trajectory, one column each: the latency-bound regime of small
        // batches (B=2048, N=1000: 4.0 vs 7.5 ms per pass); four lanes x four columns, register tiled: large batches
        // (B=32768: 39.6 vs 47.1 ms)
        const bool lanes16 = h->env_ltv_lanes ? h->env_ltv_lanes == 16 : h->p.B < 8192;
        if (lanes16) {
            const size_t smem = sizeof(T) * (size_t)(504 * TPB + 48 + 144 + 16 + 288);
            if (!h->smem_ltv) {
                cudaError_t e = cudaFuncSetAttribute(backward_ltv_kernel<T, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)smem);
                if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
                h->smem_ltv = 1;
            }
            backward_ltv_kernel<T, TPB><<<grid_for(h->p.B, TPB), TPB * 16, smem, st>>>(
                make_ltv<T>(h->p), make_cost<T, 12, 4>(h->p), h->p.N, h->p.B, (const T *)phi, (const T *)X, (const T *)U,
                (T *)K, (T *)k, active, gate, (const T *)mu);
            ILQR_CHECK_LAUNCH(h);
            return ILQR_OK;
        }
        const size_t smem = sizeof(T) * (size_t)(Ltv4Layout::PER_TRAJ * TPB + Ltv4Layout::CONST);
        if (!h->smem_ltv) {
            cudaError_t e = cudaFuncSetAttribute(backward_ltv4_kernel<T, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) { h->last_cuda = (int)e; return ILQR_E_CUDA; }
            h->smem_ltv = 1;
        }
        backward_ltv4_kernel<T, TPB><<<grid_for(h->p.B, TPB), TPB * 4, smem, st>>>(
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
                          const int *list = nullptr, const unsigned int *list_count = nullptr,
                          const SparseArgs *sparse = nullptr)
{
    SpecArgs sp;
    std::memset(&sp, 0, sizeof sp);
    if (spec) sp = *spec;
    SparseArgs sa;
    std::memset(&sa, 0, sizeof sa);
    if (sparse) sa = *sparse;
    if (h->umod) {
        const size_t threads = (size_t)n_alpha * (((size_t)h->p.B + 31) / 32 * 32) + (size_t)sp.cap * sp.n2;
        int bs = h->env_rollout_bs > 0 ? h->env_rollout_bs : block_for(threads);
        if (h->env_rollout_bs <= 0 && bs > 128) bs = 128;
        UserSysArg us{};
        UserCostAny qc(h);
        int N = h->p.N, B = h->p.B;
        AlphaList alc = al;
        void *args[] = { &us, qc.ptr(h), &N, &B, &n_alpha, &alc, &phi, &x0, &X, &U, &k, &K, &Xc, &Uc, &cost_alpha, &active, &gate,
                         &cost_ref, &sp, &list, &list_count, &sa };
        return launch_user(h, ILQR_UK_ROLLOUT, grid_for(threads, bs), bs, 0, st, args);
    }
    return dispatch(h, [&](auto tz, auto sys, auto qc, auto integ) -> int {
        using T = decltype(tz);
        using Sys = decltype(sys);
        constexpr int I = decltype(integ)::value;
        const size_t threads = (size_t)n_alpha * (((size_t)h->p.B + 31) / 32 * 32) + (size_t)sp.cap * sp.n2;
        const bool bs_env = h->env_rollout_bs > 0;
        // <= 128 threads per block: more resident blocks per SM at the kernel's register count
        int bs = bs_env ? h->env_rollout_bs : block_for(threads);
        if (!bs_env && bs > 128) bs = 128;
        // lazy waves: the warps of one trajectory group (one per step size) share a block, hence an L1
        if (!bs_env && h->lazy && n_alpha <= 4 && threads >= (size_t)148 * 16 * 32 * n_alpha) bs = 32 * n_alpha;
        auto go = [&](const auto &cost) {
            rollout_kernel<Sys, std::decay_t<decltype(cost)>, I, T><<<grid_for(threads, bs), bs, 0, st>>>(
                sys, cost, h->p.N, h->p.B, n_alpha, al, (const T *)phi, (const T *)x0, (const T *)X, (const T *)U,
                (const T *)k, (const T *)K, (T *)Xc, (T *)Uc, (T *)cost_alpha, active, gate, (const T *)cost_ref, sp, list,
                list_count, sa);
        };
        // diagonal weights (every reference script): the compact cost keeps the kernel's constants in uniform registers
        // (n <= 4 only: with the compact cost ptxas hoists the LTV model's matrix constants into registers and feeds
        // them to the FP64 pipe through R2UR moves -- 1.75x slower than its kernel with the dense cost)
        if (Sys::N <= 4 && qc.diag) go(DiagCost<T, Sys::N, Sys::M>(qc));
        else go(qc);
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
                              const SparseArgs &sa, cudaStream_t st)
{
    const int B = h->p.B, bs = 128;
    if (h->p.dtype == ILQR_F64)
        select_lazy_kernel<double><<<grid_for(B, bs), bs, 0, st>>>(B, a_lo, a_hi, wave, last, (const double *)ca,
                                                                   (double *)cost, winner, active, iters, status, h->p.tol,
                                                                   it, h->p.maxiter, ctl, list_in, cnt_in, list_out, cnt_out,
                                                                   wslot, h->tr_alpha, (double *)h->tr_cost, rg, sa);
    else
        select_lazy_kernel<float><<<grid_for(B, bs), bs, 0, st>>>(B, a_lo, a_hi, wave, last, (const float *)ca,
                                                                  (float *)cost, winner, active, iters, status,
                                                                  (float)h->p.tol, it, h->p.maxiter, ctl, list_in, cnt_in,
                                                                  list_out, cnt_out, wslot, h->tr_alpha, (float *)h->tr_cost, rg, sa);
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

// DFMA throughput probe for the FP64 roofline of bench.py: 8 independent chains per thread, operands in the
// two-register + uniform form that issues at the pipe's full rate (scripts/micro/fp64_operands.cu)
__global__ void __launch_bounds__(128) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;          // never true: keeps the chains alive without a store per thread
}

}  // namespace ilqr

using namespace ilqr;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char *ilqr_version(void) { return "ilqr_b200 0.2 (sm_100a)"; }

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

static int create_handle(const ilqr_problem_t *p, UserModule *umod, ilqr_handle_t *out);

int ilqr_create(const ilqr_problem_t *p, ilqr_handle_t *out) { return create_handle(p, nullptr, out); }

int ilqr_create_user(const ilqr_problem_t *p, ilqr_module_t mod, ilqr_handle_t *out)
{
    if (!mod) return ILQR_E_INVALID;
    return create_handle(p, (UserModule *)mod, out);
}

int ilqr_module_load(const void *image, size_t bytes, const char *const *kernel_names, int n, int m, int integrator,
                     int dtype, ilqr_module_t *out)
{
    if (!image || !bytes || !kernel_names || !out) return ILQR_E_INVALID;
    *out = nullptr;
    if (n < 1 || n > ILQR_NMAX || m < 1 || m > ILQR_MMAX || integrator < ILQR_EULER || integrator > ILQR_BACKWARD_EULER ||
        (dtype != ILQR_F64 && dtype != ILQR_F32)) return ILQR_E_INVALID;
    UserModule *u = new (std::nothrow) UserModule;
    if (!u) return ILQR_E_INVALID;
    std::memset(u, 0, sizeof *u);
    u->n = n; u->m = m; u->integrator = integrator; u->dtype = dtype;
    if (cudaLibraryLoadData(&u->lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
        cudaGetLastError();
        delete u;
        return ILQR_E_CUDA;
    }
    for (int i = 0; i < ILQR_N_USER_KERNELS; ++i) {
        if (!kernel_names[i] || cudaLibraryGetKernel(&u->k[i], u->lib, kernel_names[i]) != cudaSuccess) {
            cudaGetLastError();
            cudaLibraryUnload(u->lib);
            delete u;
            return ILQR_E_INVALID;
        }
    }
    *out = (ilqr_module_t)u;
    return ILQR_OK;
}

int ilqr_module_unload(ilqr_module_t mod)
{
    UserModule *u = (UserModule *)mod;
    if (!u) return ILQR_E_INVALID;
    cudaLibraryUnload(u->lib);
    delete u;
    return ILQR_OK;
}

static int create_handle(const ilqr_problem_t *p, UserModule *umod, ilqr_handle_t *out)
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
    case ILQR_USER:          // kernels come from a loaded module (ilqr_create_user)
        if (!umod || p->integrator != umod->integrator || p->dtype != umod->dtype) return ILQR_E_INVALID;
        n = umod->n; m = umod->m;
        break;
    default: return ILQR_E_INVALID;
    }
    if (p->n != n || p->m != m) return ILQR_E_INVALID;
    if (p->N < 1 || p->B < 1 || p->n_alpha < 1 || p->n_alpha > ILQR_MAX_ALPHAS || p->maxiter < 0) return ILQR_E_INVALID;
    if (!(p->dt > 0.0)) return ILQR_E_INVALID;
    if (p->reg_factor > 1.0 && !(p->reg_init >= 0.0 && p->reg_min > 0.0 && p->reg_max >= p->reg_min)) return ILQR_E_INVALID;
    Handle *h = new (std::nothrow) Handle;
    if (!h) return ILQR_E_INVALID;
    std::memset(h, 0, sizeof(Handle));
    h->p = *p;
    h->umod = umod;
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
    const char *e;
    {
        h->env_lanes = (e = getenv("ILQR_BACKWARD_LANES")) ? (atoi(e) != 0) : -1;
        h->env_rollout_bs = (e = getenv("ILQR_ROLLOUT_BS")) && atoi(e) >= 32 ? atoi(e) / 32 * 32 : 0;
        h->env_check_every = (e = getenv("ILQR_CHECK_EVERY")) && atoi(e) > 0 ? atoi(e) : 8;
        h->env_fused = (e = getenv("ILQR_FUSED")) ? (atoi(e) != 0) : -1;
        h->env_ltv_lanes = (e = getenv("ILQR_LTV_LANES")) && (atoi(e) == 32 || atoi(e) == 16 || atoi(e) == 4) ? atoi(e) : 0;
        h->env_ltv_wpb = (e = getenv("ILQR_LTV_MMA_WPB")) && atoi(e) == 8 ? 8 : 4;
        h->env_fused_minb = (e = getenv("ILQR_FUSED_MINB")) ? atoi(e) : 0;
        h->env_fused_split = (e = getenv("ILQR_FUSED_SPLIT")) ? (atoi(e) != 0) : -1;
        h->trig_table = p->B >= ((e = getenv("ILQR_TRIG_TABLE_MIN")) ? atol(e) : 8192L);
    }
    h->env_sparse_thresh = (e = getenv("ILQR_SPARSE_THRESH")) ? atol(e) : -1;
    h->env_sparse_all = (e = getenv("ILQR_SPARSE_ALL")) ? atol(e) : -1;
    h->sparse = (e = getenv("ILQR_SPARSE")) ? atoi(e) != 0 : 1;
    h->ab_blocked = (e = getenv("ILQR_AB_BLOCKED")) ? atoi(e) != 0 : 1;
    h->env_bulk = (e = getenv("ILQR_BACKWARD_BULK")) ? atoi(e) : -1;
    h->n_first = first_wave_size(p->B, cnt);
    h->spec_cap = spec_capacity(p->B, h->n_first, cnt);
    default_waves(h);
    if (cudaMallocHost((void **)&h->h_flag, 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev[1], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (h->ev[0]) cudaEventDestroy(h->ev[0]);
        if (h->ev[1]) cudaEventDestroy(h->ev[1]);
        if (h->h_flag) cudaFreeHost(h->h_flag);
        delete h->prof_ev;
        delete h->prof_kind;
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
    return ws_layout(h->p, h->n_alpha_eff, stores_linearization(h)).total;
}

int64_t ilqr_launch_count(ilqr_handle_t hh) { return hh ? ((Handle *)hh)->launches : 0; }
int ilqr_last_cuda_error(ilqr_handle_t hh) { return hh ? ((Handle *)hh)->last_cuda : 0; }

int ilqr_step(ilqr_handle_t hh, int t, const void *phi, const void *x, const void *u, void *xn, void *stream)
{
    Handle *h = (Handle *)hh;
    if (!h || !x || !u || !xn) return ILQR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->umod) {
        UserSysArg us{};
        ScalarAny dt(h->p.dt);
        int B = h->p.B;
        const int bs = block_for(B);
        void *args[] = { &us, dt.ptr(h), &B, &t, &phi, &x, &u, &xn };
        return launch_user(h, ILQR_UK_STEP, grid_for(B, bs), bs, 0, st, args);
    }
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
    if (h->umod) {
        UserCostAny qc(h);
        int N = h->p.N, B = h->p.B;
        void *args[] = { qc.ptr(h), &N, &B, &X, &U, &l, &lx, &lu, &lxx, &luu, &lux, &lf, &lfx, &lfxx };
        return launch_user(h, ILQR_UK_COST_EXPANSION, grid_for((size_t)(N + 1) * B, 128), 128, 0, st, args);
    }
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
    const WsLayout L = ws_layout(h->p, h->n_alpha_eff, stores_linearization(h));
    if (ws_bytes < L.total) return ILQR_E_WORKSPACE;
    char *w = (char *)ws;
    if (!X || !U || !K || !k) return ILQR_E_INVALID;
    if (h->p.model == ILQR_LTV) return launch_backward_ltv(h, phi, X, U, K, k, nullptr, nullptr, (cudaStream_t)stream);
    if (fused_available(h))                                  // the fused kernel, nothing to commit
        return launch_fused(h, phi, (void *)X, (void *)U, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr,
                            nullptr, K, k, nullptr, (cudaStream_t)stream);
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
    const WsLayout L = ws_layout(p, h->n_alpha_eff, stores_linearization(h));
    if (ws_bytes < L.total) return ILQR_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)ws;
    Control *ctl = (Control *)(w + L.ctl);
    void *A = w + L.A, *Bd = w + L.Bd, *Xc = w + L.Xc, *Uc = w + L.Uc, *ca = w + L.cost_alpha;
    int *winner = (int *)(w + L.winner), *active = (int *)(w + L.active), *defer = (int *)(w + L.defer);
    int *mark = (int *)(w + L.mark), *lists = (int *)(w + L.lists);
    int *wslot = h->lazy ? (int *)(w + L.wslot) : nullptr;      // only the lazy schedule stores candidates by list position
    // sparse mode (SparseArgs): at or below `sparse_thresh` active trajectories an iteration walks the active list;
    // at or below `sparse_all` its first rollout wave tries every step size -- by then that wave is latency bound
    // (about two warps per SM sub-partition), and n_alpha * sparse_all must fit the threads of the dense first wave
    int *alists = (int *)(w + L.alists);
    unsigned int sparse_thresh = 0, sparse_all = 0;
    if (h->lazy && h->sparse && p.model != ILQR_LTV) {
        const long w0 = h->wave_lo[1];
        long t = p.B / 4, ta = 2L * 148 * 4 * 32 / h->n_alpha_eff;
        if (ta > (long)w0 * p.B / h->n_alpha_eff) ta = (long)w0 * p.B / h->n_alpha_eff;
        if (h->env_sparse_thresh >= 0) t = h->env_sparse_thresh;
        if (h->env_sparse_all >= 0) ta = h->env_sparse_all;
        if (ta > t) ta = t;
        t &= ~31L;
        ta &= ~31L;
        if (t >= 32) { sparse_thresh = (unsigned int)t; sparse_all = (unsigned int)ta; }
    }
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
    // fused K1+K2 (ilqr_kernels_fused.cuh) wherever the model has it: at B=131072 A_t, B_t no longer stream through HBM
    // (7.26 -> 5.4 ms per iteration), at B=4096 the linearization hides behind the scan (0.328 -> 0.272 ms)
    const bool fused = fused_available(h);
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
    CU(cudaMemsetAsync(mark, 0, L.lists - L.mark, st));                 // mark and hist (adjacent regions)
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
    const int CHK = h->env_check_every;
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
            // sparse mode (SparseArgs): active lists of this / the previous / the next iteration
            SparseArgs sa;
            std::memset(&sa, 0, sizeof sa);
            if (sparse_thresh > 0) {
                sa.cur = alists + (size_t)(it & 1) * B;
                sa.n_cur = g;
                sa.prev = it > 0 ? alists + (size_t)((it - 1) & 1) * B : nullptr;
                sa.n_prev = it > 0 ? gprev : nullptr;
                sa.next = alists + (size_t)((it + 1) & 1) * B;
                sa.pos = (int *)(w + L.apos);
                sa.thresh = sparse_thresh;
                sa.thresh_all = sparse_all;
                sa.n_alpha_all = h->n_alpha_eff;
            }
            if (fused) {
                // ONE kernel commits, linearizes and scans; in the sparse iterations of the lazy schedule it walks the
                // active list and K1, in its commit-only form, first commits what the previous iteration accepted
                // (dense or sparse is decided on the device from the same counter in every kernel)
                if (sparse_thresh > 0) {
                    // commit only, sparse iterations only (finished trajectories are on no list the fused kernel walks)
                    if ((rc = launch_commit_linearize(h, phi, X, U, A, Bd, Xc, Uc, winner, it > 0 ? wslot : nullptr, active,
                                                      0, g, gprev, st, iters, it, &sa, h->ab_blocked, 1))) return rc;
                    prof_mark(h, ILQR_KC_LINEARIZE, st);
                }
                if ((rc = launch_fused(h, phi, X, U, Xc, Uc, winner, it > 0 ? wslot : nullptr, active, iters, it, g, gprev, K, k,
                                       rg.mu, st, &sa))) return rc;
                prof_mark(h, ILQR_KC_BACKWARD, st);
            } else {
            if ((rc = launch_commit_linearize(h, phi, X, U, A, Bd, Xc, Uc, winner, it > 0 ? wslot : nullptr, active,
                                              ltv ? 0 : 1, g, gprev, st, iters, it, &sa, h->ab_blocked))) return rc;
            prof_mark(h, ILQR_KC_LINEARIZE, st);
            if ((rc = ltv ? launch_backward_ltv(h, phi, X, U, K, k, active, g, st, rg.mu)
                          : launch_backward(h, X, U, A, Bd, K, k, active, g, st, rg.mu, &sa, h->ab_blocked))) return rc;
            prof_mark(h, ILQR_KC_BACKWARD, st);
            }
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
                    if (v == 0) alv = h->alphas;           // a sparse iteration tries every step size in wave 0
                    const int *lin = v ? lists + (size_t)((v - 1) & 1) * B : nullptr;
                    int *lout = lists + (size_t)(v & 1) * B;
                    const unsigned int *cin = v ? wcnt + v - 1 : nullptr;
                    if ((rc = launch_rollout(h, hi - lo, alv, phi, x0, X, U, k, K, (char *)Xc + xc_slab * lo,
                                             (char *)Uc + uc_slab * lo, (char *)ca + wbytes * (size_t)lo * B,
                                             v ? nullptr : active, v ? cin : g, nullptr, st, nullptr, lin, cin,
                                             v ? nullptr : &sa))) return rc;
                    prof_mark(h, ILQR_KC_ROLLOUT, st);
                    if ((rc = launch_select_lazy(h, lo, hi, v, v == h->n_waves - 1, ca, cost, winner, active, iters, status,
                                                 it, ctl, lin, cin, lout, wcnt + v, wslot, rg, sa, st))) return rc;
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
                // listed for the next iteration (tier 1): accepted try index >= n1 - 5 this time, or this and the previous
                // index adding up to >= n1 - 2; tier 2 (sum >= n1 - 3) fills what capacity is left.  On config 2 (10 tries,
                // n1 = 9) tier 1 is ~20 % of the batch, within the ~1000 spare slots; every miss costs a latency-bound second
                // wave, 0.37 ms at N = 500 (oracle traces of eleven shards: 1 miss per 11 solves, 5 with the single threshold;
                // measured on 2 GPUs: 8.76 -> 8.48 ms per step on the slower shard)
                sp.threshold = n1 - 5 > 1 ? n1 - 5 : 1;
                sp.list_cur = lists + (size_t)(it & 1) * B;
                sp.list_next = lists + (size_t)((it + 1) & 1) * B;
                sp.count_cur = n_spec + it;
                sp.count_next = n_spec + it + 1;
                sp.mark = mark;
                sp.hist = (int *)(w + L.hist);
                sp.list2_next = (int *)(w + L.lists2);
                unsigned int *extra = &ctl->n_active[(3 + ILQR_MAX_WAVES) * (p.maxiter + 2)];
                sp.count2_next = extra + it + 1;
                sp.ticket = extra + (p.maxiter + 2) + it;
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
                                      nullptr, st, iters, it))) return rc;
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

int ilqr_get_linesearch_waves(ilqr_handle_t hh, int32_t *sizes)
{
    Handle *h = (Handle *)hh;
    if (!h) return ILQR_E_INVALID;
    if (!h->lazy) {
        if (sizes) sizes[0] = h->n_first;      // eager: the first (dense) wave; the other tries follow where needed
        return 0;
    }
    if (sizes)
        for (int v = 0; v < h->n_waves; ++v) sizes[v] = h->wave_lo[v + 1] - h->wave_lo[v];
    return h->n_waves;
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

int ilqr_fp64_peak(double *tflops, void *scratch, void *stream)
{
    if (!tflops || !scratch) return ILQR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return ILQR_E_CUDA;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return ILQR_E_CUDA;
    const int blocks = sms * 8, threads = 128, iters = 4096;          // 8 warps per SM sub-partition
    fp64_peak_kernel<<<blocks, threads, 0, st>>>((double *)scratch, 64, 1.0000001, 1e-9);       // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        fp64_peak_kernel<<<blocks, threads, 0, st>>>((double *)scratch, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return ILQR_E_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
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
