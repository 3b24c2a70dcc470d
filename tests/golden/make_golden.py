#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference sources.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference).
`jax` is not installable here, so `oracle/jaxshim` provides the third-party `jax`
API surface on torch.func/float64 (see oracle/jaxshim/README.md); everything above
that -- class_files/iLQR_class.py, class_files/systems/*.py -- is imported from
/root/reference/python and executed as shipped.  Float64 == the reference with
JAX_ENABLE_X64=1, which is the parity target named by BASELINE.json.

    python tests/golden/make_golden.py [--only PATTERN] [--jobs 8]

Outputs tests/golden/<case>.npz.  Cases mirror the parameter blocks of the
reference run scripts (cited per case below).
"""
import argparse
import fnmatch
import math
import multiprocessing as mp
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/python"


def _setup():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
    sys.path.insert(0, REF)
    import torch
    torch.set_num_threads(1)


# --------------------------------------------------------------------------- params
PI = math.pi
# run_iLQR_open_loop.py:16-43
P_PEND_OL = dict(kind="pendulum", dt=0.01, g=9.81, l=1.0, d=0.0, Q=[1.0, 1.0], R=[1.0],
                 Q_f=[0.0, 0.0], x_target=[PI, 0.0])
# damped variant (class default d=0.01, pendulum_sys.py:29) with non-trivial terminal weight
P_PEND_D = dict(kind="pendulum", dt=0.01, g=9.81, l=1.0, d=0.01, Q=[1.0, 0.1], R=[0.1],
                Q_f=[100.0, 10.0], x_target=[PI, 0.0])
# run_iLQR_MPC.py:36-51 (single-pendulum MPC: optimizer backward_euler, plant midpoint, maxiter 10)
P_PEND_MPC = dict(kind="pendulum", dt=0.01, g=9.81, l=1.0, d=0.0, Q=[10.0, 1.0], R=[1.0],
                  Q_f=[10.0, 10.0], x_target=[PI, 0.0])
# run_iLQR_OL_UA_Pendulum.py:17-56  (== BASELINE config 2 parameter set)
P_UA_OL = dict(kind="ua", dt=0.01, g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.1, d2=0.1,
               theta1=1.0 / 12, theta2=1.0 / 12, Q=[1.0, 1.0, 0.1, 0.1], R=[1.0],
               Q_f=[1000.0, 1000.0, 100.0, 100.0], x_target=[PI, 0.0, 0.0, 0.0])
# run_iLQR_UA_MPC.py:17-67 (config 3)
P_UA_MPC = dict(P_UA_OL, Q=[5.0, 5.0, 0.1, 0.1], R=[50.0], Q_f=[1000.0, 1000.0, 10.0, 10.0])
# run_double_pendulum_open_loop.py:16-55
P_DP_OL = dict(kind="double", dt=0.01, g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.1, d2=0.1,
               theta1=1.0 / 12, theta2=1.0 / 12, Q=[10.0, 10.0, 0.1, 0.1], R=[0.1, 0.1],
               Q_f=[1000.0, 1000.0, 100.0, 100.0], x_target=[PI, 0.0, 0.0, 0.0])


# DENSE, NON-SYMMETRIC weights on the fully actuated double pendulum (m = 2) and the under-actuated one: no shipped script
# uses them, but the reference accepts any array -- the cost is the written quadratic form and autodiff differentiates it
# (system_base.py:212-219), i.e. the derivatives see the symmetric part.  Pins that behaviour for the oracle and the kernels.
_QD = [[2.0, 0.3, -0.1, 0.05], [-0.2, 1.5, 0.1, 0.0], [0.1, -0.3, 0.4, 0.02], [0.0, 0.1, -0.05, 0.3]]
_QFD = [[300.0, 40.0, -5.0, 2.0], [-20.0, 250.0, 6.0, 1.0], [5.0, -3.0, 40.0, 4.0], [1.0, 2.0, -6.0, 30.0]]
P_DP_DENSE = dict(P_DP_OL, Q=_QD, R=[[0.2, 0.05], [-0.03, 0.15]], Q_f=_QFD)
P_UA_DENSE = dict(P_UA_OL, Q=_QD, R=[[0.8]], Q_f=_QFD)


def make_system(p, integrator):
    import numpy as np
    import jax.numpy as jnp
    diag = lambda v: jnp.diag(jnp.array(v)) if np.ndim(v) == 1 else jnp.array(v)     # weights as diagonals or full matrices
    common = dict(dt=p["dt"], x_target=jnp.array(p["x_target"]), Q=diag(p["Q"]), R=diag(p["R"]),
                  Q_f=diag(p["Q_f"]), integrator=integrator, use_jit=True)
    if p["kind"] == "user_cartpole":
        # a USER-DEFINED subclass of the reference's System (tests/user_systems.py): the reference's own
        # jit + autodiff factory (system_base.py:203-251) supplies every derivative
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from class_files.systems.system_base import System
        from user_systems import make_cartpole_class
        extra = {k: p[k] for k in ("mc", "mp", "l", "g", "b", "p_max", "w_bar")}
        return make_cartpole_class(System, jnp)(**extra, **common)
    if p["kind"] == "user_spring":
        # a user-defined subclass with a lax.while_loop (data-dependent trip count) and a lax.cond in its dynamics: the
        # reference's jacfwd differentiates THROUGH the loop (system_base.py:204-205)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from jax import lax
        from class_files.systems.system_base import System
        from user_systems import make_implicit_spring_class
        return make_implicit_spring_class(System, jnp, lax)(a=p["a"], ks=p["ks"], **common)
    if p["kind"] == "pendulum":
        from class_files.systems.pendulum_sys import MyPendulum
        return MyPendulum(g=p["g"], l=p["l"], d=p["d"], **common)
    phys = {k: p[k] for k in ("g", "m1", "m2", "l1", "l2", "d1", "d2", "theta1", "theta2")}
    if p["kind"] == "double":
        from class_files.systems.double_pendulum_sys import MyDoublePendulum
        return MyDoublePendulum(**phys, **common)
    from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum
    return MyUADoublePendulum(**phys, **common)


def params_arrays(p, integrator):
    import numpy as np
    out = {"p_kind": np.array(p["kind"]), "p_integrator": np.array(integrator)}
    for k, v in p.items():
        if k != "kind":
            out["p_" + k] = np.asarray(v, dtype=np.float64)
    return out


def npy(t):
    import numpy as np
    import torch
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().astype(np.float64)
    return np.asarray(t, dtype=np.float64)


def cfg2_x0(count, seed=0):
    """BASELINE config 2 initial states: default_rng(0); q~U(-pi,pi), qd~U(-2,2)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    x0 = np.empty((count, 4))
    x0[:, :2] = rng.uniform(-PI, PI, size=(count, 2))
    x0[:, 2:] = rng.uniform(-2.0, 2.0, size=(count, 2))
    return x0


# --------------------------------------------------------------------------- cases
def case_derivs(p, integrator, seed):
    """Point evaluations of the 12 System callables (system_base.py:223-251)."""
    import numpy as np
    import jax.numpy as jnp
    s = make_system(p, integrator)
    rng = np.random.default_rng(seed)
    n, m = s.n_x, s.n_u
    P = 8
    xs = rng.uniform(-3.5, 3.5, size=(P, n))
    us = rng.uniform(-4.0, 4.0, size=(P, m))
    names = ["f", "f_x", "f_u", "l", "l_x", "l_u", "l_xx", "l_uu", "l_ux", "l_f", "l_f_x", "l_f_xx"]
    acc = {k: [] for k in names}
    for x, u in zip(xs, us):
        x = jnp.array(x); u = jnp.array(u)
        acc["f"].append(npy(s.f_fcn(x, u)))
        acc["f_x"].append(npy(s.f_x_fcn(x, u)))
        acc["f_u"].append(npy(s.f_u_fcn(x, u)))
        acc["l"].append(npy(s.l_fcn(x, u)))
        acc["l_x"].append(npy(s.l_x_fcn(x, u)))
        acc["l_u"].append(npy(s.l_u_fcn(x, u)))
        acc["l_xx"].append(npy(s.l_xx_fcn(x, u)))
        acc["l_uu"].append(npy(s.l_uu_fcn(x, u)))
        acc["l_ux"].append(npy(s.l_ux_fcn(x, u)))
        acc["l_f"].append(npy(s.l_f_fcn(x)))
        acc["l_f_x"].append(npy(s.l_f_x_fcn(x)))
        acc["l_f_xx"].append(npy(s.l_f_xx_fcn(x)))
    out = params_arrays(p, integrator)
    out.update(xs=xs, us=us, **{k: np.stack(v) for k, v in acc.items()})
    return out


def case_passes(p, integrator, T, seed):
    """backward_pass / forward_pass on a given nominal (iLQR_class.py:122-161,193-247)."""
    import numpy as np
    import jax.numpy as jnp
    from class_files.iLQR_class import iLQR
    s = make_system(p, integrator)
    n, m = s.n_x, s.n_u
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-1.0, 1.0, size=n)
    sol = iLQR(s, T, jnp.array(x0), jnp.zeros((m, int(round(T / p["dt"])))), verbose=False)
    N = sol.N
    U_nom = jnp.array(rng.uniform(-1.0, 1.0, size=(m, N)))
    # a dynamically consistent nominal: rollout with alpha=0, zero gains
    X_nom, U_nom2, cost0 = sol.forward_pass(jnp.array(x0), 0.0, sol.X, U_nom, sol.U_ff, sol.K)
    U_ff, K = sol.backward_pass(X_nom, U_nom2)
    out = params_arrays(p, integrator)
    out.update(T=T, N=N, x0=x0, U_nom=npy(U_nom2), X_nom=npy(X_nom), cost0=npy(cost0),
               U_ff=npy(U_ff), K=npy(K))
    x0b = x0 + 0.01 * rng.standard_normal(n)   # perturbed start exercises the K(x - x_old) term
    out["x0_b"] = x0b
    for a in (1.0, 0.5, 0.125):
        Xn, Un, c = sol.forward_pass(jnp.array(x0b), a, X_nom, U_nom2, U_ff, K)
        tag = str(a).replace(".", "p")
        out.update({f"X_a{tag}": npy(Xn), f"U_a{tag}": npy(Un), f"cost_a{tag}": npy(c)})
    return out


class _Trace:
    def __init__(self, sol):
        self.fw = []       # (alpha, cost) per forward_pass call
        self.n_bw = 0
        self.snap = []     # (X, U, U_ff, K) per backward_pass call
        f0, b0 = sol.forward_pass, sol.backward_pass

        def fwd(x0, alpha, X, U, U_ff, K):
            r = f0(x0, alpha, X, U, U_ff, K)
            self.fw.append((float(alpha), float(r[2])))
            return r

        def bwd(X, U):
            self.n_bw += 1
            r = b0(X, U)
            # per-iteration snapshot: the nominal the iteration starts from and the gains it produces
            self.snap.append((npy(X), npy(U), npy(r[0]), npy(r[1])))
            return r
        sol.forward_pass, sol.backward_pass = fwd, bwd


def case_solve(p, integrator, T, x0, maxiter, tol):
    """Full optimize_trajectory() (iLQR_class.py:250-313) with call trace."""
    import numpy as np
    import jax.numpy as jnp
    from class_files.iLQR_class import iLQR
    s = make_system(p, integrator)
    m = s.n_u
    N = len(np.arange(0, T + p["dt"], p["dt"])) - 1
    sol = iLQR(s, T, jnp.array(x0), jnp.zeros((m, N)), tol=tol, maxiter=maxiter, verbose=False)
    tr = _Trace(sol)
    X, U, cost = sol.optimize_trajectory()
    out = params_arrays(p, integrator)
    out.update(T=T, N=sol.N, x0=np.asarray(x0, dtype=np.float64), maxiter=maxiter, tol=tol,
               X=npy(X), U=npy(U), cost=npy(cost), K=npy(sol.K), U_ff=npy(sol.U_ff),
               trace_alpha=np.array([a for a, _ in tr.fw]), trace_cost=np.array([c for _, c in tr.fw]),
               n_backward=tr.n_bw)
    if tr.snap:
        out.update(it_X=np.stack([s[0] for s in tr.snap]), it_U=np.stack([s[1] for s in tr.snap]),
                   it_U_ff=np.stack([s[2] for s in tr.snap]), it_K=np.stack([s[3] for s in tr.snap]))
    return out


def case_mpc(p, integ_opt, integ_plant, T_h, ticks, x0, maxiter, tol):
    """Receding-horizon loop exactly as run_iLQR_UA_MPC.py:146-174 (one solver object re-used)."""
    import numpy as np
    import jax.numpy as jnp
    from class_files.iLQR_class import iLQR
    s = make_system(p, integ_opt)
    plant = make_system(p, integ_plant)
    n, m = s.n_x, s.n_u
    N = len(np.arange(0, T_h + p["dt"], p["dt"])) - 1
    sol = iLQR(s, T_h, jnp.array(x0), jnp.zeros((m, N)), tol=tol, maxiter=maxiter, verbose=False)
    tr = _Trace(sol)
    X_sim = np.zeros((n, ticks + 1)); U_sim = np.zeros((m, ticks)); costs = np.zeros(ticks)
    nbw = np.zeros(ticks, dtype=np.int64)
    current_x = jnp.array(x0); X_sim[:, 0] = npy(current_x)
    U_guess = jnp.zeros((m, N))
    Xb_all, Ub_all = [], []
    for k in range(ticks):
        sol.x_0 = current_x
        sol.U = U_guess
        b0 = tr.n_bw
        X_bar, U_bar, cost = sol.optimize_trajectory()
        nbw[k] = tr.n_bw - b0
        uk = U_bar[:, 0]
        xkp = plant.f_fcn(current_x, uk)
        U_sim[:, k] = npy(uk); X_sim[:, k + 1] = npy(xkp); costs[k] = float(cost)
        Xb_all.append(npy(X_bar)); Ub_all.append(npy(U_bar))
        U_guess = jnp.concatenate([U_bar[:, 1:], U_bar[:, -1:]], axis=1)
        current_x = xkp
    out = params_arrays(p, integ_opt)
    out.update(p_integrator_plant=np.array(integ_plant), T=T_h, N=N, ticks=ticks,
               x0=np.asarray(x0, dtype=np.float64), maxiter=maxiter, tol=tol, X_sim=X_sim, U_sim=U_sim,
               costs=costs, n_backward=nbw, X_bar=np.stack(Xb_all), U_bar=np.stack(Ub_all),
               K_last=npy(sol.K), U_ff_last=npy(sol.U_ff))
    return out


def build_cases():
    cases = {}
    integs = ["euler", "midpoint", "rk4", "backward_euler"]
    # user-defined System subclass (cart-pole with a non-quadratic barrier cost), tests/user_systems.py
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from user_systems import CARTPOLE
    p_cp = dict(CARTPOLE, kind="user_cartpole")
    for i, integ in enumerate(integs):
        cases[f"user_cartpole_derivs_{integ}"] = (case_derivs, (p_cp, integ, 300 + i))
    cases["user_cartpole_passes_rk4"] = (case_passes, (p_cp, "rk4", 0.6, 310))
    cases["user_cartpole_solve_rk4_T1"] = (case_solve, (p_cp, "rk4", 1.0, [0.0, 0.3, 0.0, 0.0], 30, 1e-5))
    cases["user_cartpole_solve_be_T1"] = (case_solve, (p_cp, "backward_euler", 1.0, [0.2, -0.2, 0.0, 0.5], 20, 1e-5))
    cases["user_cartpole_mpc_T0p5"] = (case_mpc, (p_cp, "rk4", "midpoint", 0.5, 4, [0.1, 2.9, 0.0, 0.0], 20, 1e-5))
    # user-defined subclass with lax.while_loop + lax.cond in the dynamics (tests/user_systems.py)
    from user_systems import SPRING
    p_sp = dict(SPRING, kind="user_spring")
    for i, integ in enumerate(integs):
        cases[f"user_spring_derivs_{integ}"] = (case_derivs, (p_sp, integ, 320 + i))
    cases["user_spring_passes_rk4"] = (case_passes, (p_sp, "rk4", 0.6, 330))
    cases["user_spring_solve_rk4_T1"] = (case_solve, (p_sp, "rk4", 1.0, [0.4, -0.3], 30, 1e-5))
    cases["user_spring_solve_midpoint_T1"] = (case_solve, (p_sp, "midpoint", 1.0, [-1.0, 0.8], 30, 1e-5))
    for tag, p in (("pend", P_PEND_D), ("double", P_DP_OL), ("ua", P_UA_OL)):
        for i, integ in enumerate(integs):
            cases[f"derivs_{tag}_{integ}"] = (case_derivs, (p, integ, 100 + i))
            cases[f"passes_{tag}_{integ}"] = (case_passes, (p, integ, 0.6, 200 + i))
    # --- full solves -------------------------------------------------------------
    # config 1: run_iLQR_open_loop.py as shipped (N=400, backward_euler, maxiter=100)
    cases["solve_cfg1_pend_be"] = (case_solve, (P_PEND_OL, "backward_euler", 4.0, [1.0, 0.0], 100, 1e-5))
    cases["solve_pend_rk4_T1"] = (case_solve, (P_PEND_D, "rk4", 1.0, [0.0, 0.0], 60, 1e-5))
    cases["solve_pend_euler_T2"] = (case_solve, (P_PEND_OL, "euler", 2.0, [1.0, 0.0], 60, 1e-5))
    # config 2 parameter set, the first initial states of the seeded batch, N=500 rk4
    x0s = cfg2_x0(8)
    for b in range(3):
        cases[f"solve_cfg2_ua_rk4_b{b}"] = (case_solve, (P_UA_OL, "rk4", 5.0, x0s[b].tolist(), 25, 1e-5))
    # the same systems at a short horizon so the CPU suite has many fast full solves
    for b in range(8):
        cases[f"solve_ua_rk4_T1_b{b}"] = (case_solve, (P_UA_OL, "rk4", 1.0, x0s[b].tolist(), 40, 1e-5))
    for b in range(3):
        cases[f"solve_ua_euler_T1_b{b}"] = (case_solve, (P_UA_OL, "euler", 1.0, x0s[b].tolist(), 40, 1e-5))
        cases[f"solve_ua_midpoint_T1_b{b}"] = (case_solve, (P_UA_OL, "midpoint", 1.0, x0s[b].tolist(), 40, 1e-5))
        cases[f"solve_double_rk4_T1_b{b}"] = (case_solve, (P_DP_OL, "rk4", 1.0, x0s[b].tolist(), 40, 1e-6))
    cases["solve_ua_be_T1_b0"] = (case_solve, (P_UA_OL, "backward_euler", 1.0, x0s[0].tolist(), 25, 1e-5))
    # run_double_pendulum_open_loop.py as shipped except the iteration cap (N=500, euler, m=2)
    cases["solve_double_euler_T5"] = (case_solve, (P_DP_OL, "euler", 5.0, [0.0, 0.0, 0.0, 0.0], 30, 1e-6))
    # run_iLQR_OL_UA_Pendulum.py down-down start at a shorter horizon (shipped: T=8, backward_euler)
    cases["solve_ua_rk4_T2_down"] = (case_solve, (P_UA_OL, "rk4", 2.0, [0.0, 0.0, 0.0, 0.0], 40, 1e-5))
    # dense non-symmetric weights (P_DP_DENSE, P_UA_DENSE above)
    cases["derivs_double_dense_rk4"] = (case_derivs, (P_DP_DENSE, "rk4", 120))
    cases["derivs_ua_dense_euler"] = (case_derivs, (P_UA_DENSE, "euler", 121))
    cases["passes_double_dense_rk4"] = (case_passes, (P_DP_DENSE, "rk4", 0.6, 220))
    cases["passes_ua_dense_midpoint"] = (case_passes, (P_UA_DENSE, "midpoint", 0.6, 221))
    cases["solve_double_dense_rk4_T1"] = (case_solve, (P_DP_DENSE, "rk4", 1.0, x0s[1].tolist(), 15, 1e-5))
    cases["solve_ua_dense_rk4_T1"] = (case_solve, (P_UA_DENSE, "rk4", 1.0, x0s[2].tolist(), 15, 1e-5))
    # --- MPC ----------------------------------------------------------------------
    cases["mpc_ua_T0p5_ticks6"] = (case_mpc, (P_UA_MPC, "rk4", "backward_euler", 0.5, 6,
                                             [0.1, -0.1, 0.3, -0.2], 50, 1e-5))
    cases["mpc_cfg3_ua_T2_ticks2"] = (case_mpc, (P_UA_MPC, "rk4", "backward_euler", 2.0, 2,
                                                [0.05, -0.08, 0.4, -0.3], 50, 1e-5))
    cases["mpc_pend_be_T1_ticks5"] = (case_mpc, (P_PEND_MPC, "backward_euler", "midpoint", 1.0, 5,
                                                [0.0, 0.0], 10, 1e-5))
    return cases


def _run(item):
    name, (fn, args) = item
    _setup()
    import numpy as np
    t = time.time()
    out = fn(*args)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return name, time.time() - t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="*")
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--skip-existing", action="store_true")
    a = ap.parse_args()
    cases = {k: v for k, v in build_cases().items() if fnmatch.fnmatch(k, a.only)}
    if a.skip_existing:
        cases = {k: v for k, v in cases.items() if not os.path.exists(os.path.join(HERE, k + ".npz"))}
    print(f"{len(cases)} cases", flush=True)
    with mp.get_context("spawn").Pool(a.jobs) as pool:
        for name, dt in pool.imap_unordered(_run, list(cases.items())):
            print(f"  {name}: {dt:.1f}s", flush=True)


if __name__ == "__main__":
    main()
