"""Receding-horizon MPC loop on the device -- the loop the reference writes out in its run scripts
(run_iLQR_UA_MPC.py:146-174, run_iLQR_MPC.py:116-143, run_MPC_double_pendulum.py:142-170):

    for k in range(N_sim):
        ilqr_solver.x_0 = current_x                    # :148
        ilqr_solver.U   = U_guess                      # :151
        X_bar, U_bar, cost = ilqr_solver.optimize_trajectory()   # :154
        uk = U_bar[:, 0]                               # :157
        xkPlusOne = plant.f_fcn(current_x, uk)         # :161  (the plant may use another integrator)
        U_guess = concatenate([U_bar[:, 1:], U_bar[:, -1:]])     # :168
        current_x = xkPlusOne                          # :171

That script-style loop runs unchanged against this package (attributes and f_fcn are drop-ins).
`run_mpc` is the same loop for a whole batch of MPC instances with every tick's solve, warm-start
shift and plant step staying on the GPU: the solver object is re-used across ticks, so -- exactly as
in the reference -- X, K and U_ff persist and each tick's alpha=0 rollout applies the previous tick's
gains about the previous tick's trajectory.
"""
import numpy as np
import torch

from . import _device as D


def run_mpc(ilqr_solver, plant_system, x_0, N_sim, U_init=None, record_plans=False):
    """Closed-loop simulation of `N_sim` control ticks.

    ilqr_solver : class_files.iLQR_class.iLQR (batched or not); its state is advanced in place.
    plant_system: System used as the "real" plant (same model family/dimensions, any integrator).
    x_0         : (n,) or (B, n) initial plant state(s).
    Returns dict(X_sim ((B,) n, N_sim+1), U_sim ((B,) m, N_sim), costs ((B,) N_sim), iterations ((B,) N_sim)
    [, X_bar, U_bar per tick when record_plans]) as numpy arrays, or CUDA tensors if x_0 is a CUDA tensor.
    """
    sol = ilqr_solver
    n, m, B, N = sol.n_x, sol.n_u, sol.B, sol.N
    if (plant_system.n_x, plant_system.n_u) != (n, m):
        raise ValueError("plant and optimizer systems must have the same dimensions")
    torch_out = D.is_torch(x_0) and x_0.is_cuda
    tdt = sol._tdt
    if getattr(plant_system, "TIME_VARYING", False) or getattr(sol.system, "TIME_VARYING", False):
        # the solver's horizon always starts at time index 0 while the plant's clock runs on: for a time-varying
        # model the two would silently disagree (the reference has no such system; config 4 is open loop)
        raise NotImplementedError("run_mpc does not support time-varying systems (MyLTVSystem)")
    plant = D.Handle(plant_system.make_problem(N=1, B=B), lib=plant_system._library())
    lib = plant.lib
    dev = dict(dtype=tdt, device="cuda")
    X_sim = torch.empty((N_sim + 1, n, B), **dev)
    U_sim = torch.empty((N_sim, m, B), **dev)
    costs = torch.empty((N_sim, B), **dev)
    iters = torch.empty((N_sim, B), dtype=torch.int32, device="cuda")
    plans = ([], []) if record_plans else None

    cur = D.to_device(x_0, tdt).reshape(-1, n)
    if cur.shape[0] != B:
        raise ValueError(f"x_0 must hold {B} initial state(s), got {tuple(cur.shape)}")
    X_sim[0].copy_(cur.t())
    if U_init is not None:
        sol.U = U_init
    for k in range(N_sim):
        sol._x0.copy_(X_sim[k])                                  # :148
        sol.solve_device(sync=False)                             # :154 (U already holds the warm start)
        costs[k].copy_(sol._cost)
        iters[k].copy_(sol._iters)
        if record_plans:
            plans[0].append(sol._X.clone())
            plans[1].append(sol._U.clone())
        u0 = sol.mpc_shift()                                     # :157 and :168 in one pass over U
        U_sim[k].copy_(u0)
        plant.check(lib.ilqr_step(plant.h, k, None, D.ptr(X_sim[k]), D.ptr(u0), D.ptr(X_sim[k + 1]),
                                  D.stream_ptr()))               # :161
    torch.cuda.current_stream().synchronize()

    def fin(t, time_last=True):
        # device [time][dims..][B] -> ((B,) dims.., time)
        perm = [t.ndim - 1] + list(range(1, t.ndim - 1)) + [0] if time_last else [t.ndim - 1] + list(range(t.ndim - 1))
        v = t.permute(*perm)
        if not sol.batched:
            v = v[0]
        return v if torch_out else D.to_host(v)

    out = dict(X_sim=fin(X_sim), U_sim=fin(U_sim), costs=fin(costs), iterations=fin(iters))
    if record_plans:
        out["X_bar"] = [fin(x) for x in plans[0]]
        out["U_bar"] = [fin(u) for u in plans[1]]
    return out
