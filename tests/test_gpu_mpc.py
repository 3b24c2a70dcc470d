"""GPU: the receding-horizon loop (run_iLQR_UA_MPC.py:146-174) -- once written out script-style against
the drop-in attributes exactly as the reference scripts do, once through class_files.mpc.run_mpc with
everything device-resident -- against golden vectors from the unmodified reference and the CPU oracle."""
import numpy as np
import pytest

from helpers import mpc_member_parity, write_report

from conftest import golden_names, load_golden, rel_err
from helpers import system_from_golden, SENS_FACTOR as SF

pytestmark = pytest.mark.gpu
TOL = 1e-9


def mpc_sensitivity(O, g):
    """drift of the closed loop under rounding noise (FMA build of the oracle + 1e-14 on x0)"""
    p_opt = O.problem_from_golden(g)
    p_plant = O.problem_from_golden(g, integrator=str(g["p_integrator_plant"]))
    base = O.mpc(p_opt, p_plant, g["x0"], int(g["ticks"]))
    s = dict(X_sim=0.0, U_sim=0.0, costs=0.0)
    runs = []
    with O.rounding_variant():
        runs.append(O.mpc(p_opt, p_plant, g["x0"], int(g["ticks"])))
    x0 = np.asarray(g["x0"], dtype=np.float64)
    xp = np.where(x0 != 0, x0 * (1 + 1e-14), 1e-14)
    runs.append(O.mpc(p_opt, p_plant, xp, int(g["ticks"])))
    for r in runs:
        for k, floor in (("X_sim", 0.0), ("U_sim", 1e-3), ("costs", 0.0)):
            s[k] = max(s[k], rel_err(r[k], base[k], floor=floor))
    return s


@pytest.mark.parametrize("name", golden_names("mpc_"))
def test_script_style_loop_vs_reference(name, oracle):
    from class_files.iLQR_class import iLQR
    g = load_golden(name)
    opt = system_from_golden(g)
    plant = system_from_golden(g, integrator=str(g["p_integrator_plant"]))
    n, m, N, ticks = opt.n_x, opt.n_u, int(g["N"]), int(g["ticks"])
    sol = iLQR(opt, float(g["T"]), g["x0"], np.zeros((m, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]),
               verbose=False)
    X_sim, U_sim = np.zeros((n, ticks + 1)), np.zeros((m, ticks))
    current_x = np.array(g["x0"], dtype=np.float64)
    X_sim[:, 0] = current_x
    U_guess = np.zeros((m, N))
    costs, its = [], []
    for k in range(ticks):                       # the reference's loop, verbatim modulo jnp -> np
        sol.x_0 = current_x
        sol.U = U_guess
        X_bar, U_bar, cost = sol.optimize_trajectory()
        uk = U_bar[:, 0]
        xkPlusOne = plant.f_fcn(current_x, uk)
        U_sim[:, k] = uk
        X_sim[:, k + 1] = xkPlusOne
        U_guess = np.concatenate([U_bar[:, 1:], U_bar[:, -1:]], axis=1)
        current_x = xkPlusOne
        costs.append(float(cost))
        its.append(int(sol.iterations))
    s = mpc_sensitivity(oracle, g)
    assert its == list(g["n_backward"])
    assert rel_err(X_sim, g["X_sim"]) <= max(TOL, SF * s["X_sim"])
    assert rel_err(U_sim, g["U_sim"], floor=1e-3) <= max(TOL, SF * s["U_sim"])
    assert rel_err(costs, g["costs"]) <= max(TOL, SF * s["costs"])


@pytest.mark.parametrize("name", golden_names("mpc_"))
def test_run_mpc_vs_reference(name, oracle):
    from class_files.iLQR_class import iLQR
    from class_files.mpc import run_mpc
    g = load_golden(name)
    opt = system_from_golden(g)
    plant = system_from_golden(g, integrator=str(g["p_integrator_plant"]))
    m, N, ticks = opt.n_u, int(g["N"]), int(g["ticks"])
    sol = iLQR(opt, float(g["T"]), g["x0"], np.zeros((m, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]),
               verbose=False)
    r = run_mpc(sol, plant, g["x0"], ticks, record_plans=True)
    s = mpc_sensitivity(oracle, g)
    assert list(r["iterations"]) == list(g["n_backward"])
    assert r["X_sim"].shape == g["X_sim"].shape and r["U_sim"].shape == g["U_sim"].shape
    assert rel_err(r["X_sim"], g["X_sim"]) <= max(TOL, SF * s["X_sim"])
    assert rel_err(r["U_sim"], g["U_sim"], floor=1e-3) <= max(TOL, SF * s["U_sim"])
    assert rel_err(r["costs"], g["costs"]) <= max(TOL, SF * s["costs"])
    assert rel_err(np.stack(r["X_bar"]), g["X_bar"]) <= max(TOL, SF * s["X_sim"])
    assert rel_err(sol.K, g["K_last"]) <= max(1e-8, SF * s["X_sim"])


def test_batched_mpc_vs_oracle(oracle):
    """config-3 style: many concurrent MPC instances (here 96), 8 line-search alphas, rk4 optimizer,
    backward-Euler plant; every instance against its own oracle run."""
    from class_files.iLQR_class import iLQR
    from class_files.mpc import run_mpc
    g = load_golden("mpc_cfg3_ua_T2_ticks2")
    opt = system_from_golden(g)
    plant = system_from_golden(g, integrator="backward_euler")
    B, T, N, ticks = 96, 0.4, 40, 4
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal((B, 4)) * np.array([0.1, 0.1, 0.5, 0.5])
    sol = iLQR(opt, T, x0, np.zeros((1, N)), maxiter=50, verbose=False, n_alpha=8)
    r = run_mpc(sol, plant, x0, ticks)
    p_opt = oracle.problem_from_golden(g, maxiter=50, n_alpha=8)
    p_opt.N = N
    p_plant = oracle.problem_from_golden(g, integrator="backward_euler")
    report, failures = mpc_member_parity(oracle, p_opt, p_plant, x0, ticks, np.asarray(r["X_sim"]), np.asarray(r["iterations"]))
    write_report("mpc_cfg3_style_B96_N40_ticks4", report)
    assert not failures, (failures[:5], report)
    assert report["same_flow"] >= 0.9 * B


def test_examples_run(capsys):
    """the example scripts (the reference's run scripts without plots) run end to end"""
    import importlib.util
    import os
    from conftest import ROOT
    ex = os.path.join(ROOT, "examples")
    import sys
    sys.path.insert(0, ex)
    for name, kw in (("open_loop_pendulum", {}), ("ua_mpc", dict(N_sim=3, B=256)), ("batched_swing_up", dict(B=256))):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ex, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.main(**kw)
    out = capsys.readouterr().out
    assert "Initial cost:" in out and "run_mpc, 256 instances" in out and "256 trajectories" in out
