"""CPU: the C oracle (oracle/ilqr_oracle.c) against golden vectors produced by running the UNMODIFIED
reference sources (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from helpers import SENS_FACTOR as SF
from helpers import golden_flow, rounding_sensitivity, backward_sensitivity, forward_sensitivity

TOL = 1e-9


@pytest.mark.parametrize("name", golden_names("derivs_"))
def test_point_functions(oracle, name):
    O, g = oracle, load_golden(name)
    p = O.problem_from_golden(g)
    for i, (x, u) in enumerate(zip(g["xs"], g["us"])):
        A, B = O.f_jac(p, x, u)
        lx, lu, lxx, luu, lux = O.l_derivs(p, x, u)
        lfx, lfxx = O.lf_derivs(p, x)
        got = dict(f=O.f(p, x, u), f_x=A, f_u=B, l=O.l(p, x, u), l_x=lx, l_u=lu, l_xx=lxx, l_uu=luu, l_ux=lux,
                   l_f=O.lf(p, x), l_f_x=lfx, l_f_xx=lfxx)
        for k, v in got.items():
            assert rel_err(v, g[k][i], floor=1e-3) < 1e-12, (k, i)


@pytest.mark.parametrize("name", golden_names("passes_"))
def test_passes(oracle, name):
    O, g = oracle, load_golden(name)
    p = O.problem_from_golden(g)
    z = np.zeros_like
    X0, U0, c0 = O.forward_pass(p, g["x0"], 0.0, z(g["X_nom"]), g["U_nom"], z(g["U_ff"]), z(g["K"]))
    assert rel_err(X0, g["X_nom"]) < 1e-13 and rel_err(c0, g["cost0"]) < 1e-13
    U_ff, K = O.backward_pass(p, g["X_nom"], g["U_nom"])
    assert rel_err(U_ff, g["U_ff"]) < 1e-10 and rel_err(K, g["K"]) < 1e-10
    for tag, a in (("1p0", 1.0), ("0p5", 0.5), ("0p125", 0.125)):
        Xn, Un, c = O.forward_pass(p, g["x0_b"], a, g["X_nom"], g["U_nom"], g["U_ff"], g["K"])
        assert rel_err(Xn, g["X_a" + tag]) < 1e-12 and rel_err(Un, g["U_a" + tag]) < 1e-12
        assert rel_err(c, g["cost_a" + tag]) < 1e-13


@pytest.mark.parametrize("name", golden_names("solve_"))
def test_every_iteration_on_identical_inputs(oracle, name):
    O, g = oracle, load_golden(name)
    if "it_X" not in g:
        pytest.skip("golden file has no per-iteration snapshots")
    p = O.problem_from_golden(g)
    idx, costs = golden_flow(g)
    n_it = len(idx)
    for i in range(n_it):
        Xi, Ui = g["it_X"][i], g["it_U"][i]
        U_ff, K = O.backward_pass(p, Xi, Ui)
        sK, sU = backward_sensitivity(O, p, Xi, Ui)
        assert rel_err(K, g["it_K"][i]) <= max(TOL, SF * sK), (i, rel_err(K, g["it_K"][i]), sK)
        assert rel_err(U_ff, g["it_U_ff"][i], floor=1e-6) <= max(TOL, SF * sU), (i, sU)
        if idx[i] < 0:
            continue
        a = 0.5 ** idx[i]
        Xn, Un, c = O.forward_pass(p, g["x0"], a, Xi, Ui, g["it_U_ff"][i], g["it_K"][i])
        sX, sUn, sc = forward_sensitivity(O, p, g["x0"], a, Xi, Ui, g["it_U_ff"][i], g["it_K"][i])
        X_ref = g["it_X"][i + 1] if i + 1 < n_it else g["X"]
        U_ref = g["it_U"][i + 1] if i + 1 < n_it else g["U"]
        assert rel_err(Xn, X_ref) <= max(TOL, SF * sX), (i, rel_err(Xn, X_ref), sX)
        assert rel_err(Un, U_ref, floor=1e-3) <= max(TOL, SF * sUn), (i, sUn)
        assert rel_err(c, costs[i + 1]) <= max(TOL, SF * sc), (i, sc)


@pytest.mark.parametrize("name", golden_names("solve_"))
def test_full_solves(oracle, name):
    """End-to-end optimize_trajectory() against the reference.  Bound: 1e-9, or 30x what a 1e-14 input
    perturbation does to the same solve (helpers.rounding_sensitivity) where that is larger."""
    O, g = oracle, load_golden(name)
    p = O.problem_from_golden(g)
    r = O.optimize(p, g["x0"], np.zeros((p.m, p.N)))
    ref_idx, ref_costs = golden_flow(g)
    sens = rounding_sensitivity(O, p, g["x0"])
    assert int(g["n_backward"]) == len(ref_idx)
    assert abs(r["cost0"] - ref_costs[0]) <= 1e-13 * abs(ref_costs[0])
    k = min(sens["stable_prefix"], len(ref_idx), r["iters"])
    assert k >= min(3, len(ref_idx))
    assert np.array_equal(r["alpha_idx"][:k], ref_idx[:k])
    err = np.abs(r["cost_trace"][:k] - ref_costs[1:k + 1]) / np.abs(ref_costs[1:k + 1])
    assert np.all(err <= np.maximum(TOL, SF * sens["cost"][:k])), (err, sens["cost"][:k])
    if sens["flow_stable"]:
        assert r["iters"] == len(ref_idx) and np.array_equal(r["alpha_idx"], ref_idx)
        assert rel_err(r["cost"], g["cost"]) <= max(TOL, SF * sens["cost"][-1])
        for key, floor in (("X", 0.0), ("U", 1e-3), ("K", 0.0), ("U_ff", 1e-3)):
            assert rel_err(r[key], g[key], floor=floor) <= max(TOL, SF * sens[key]), key


@pytest.mark.parametrize("name", golden_names("mpc_"))
def test_mpc(oracle, name):
    O, g = oracle, load_golden(name)
    p_opt = O.problem_from_golden(g)
    p_plant = O.problem_from_golden(g, integrator=str(g["p_integrator_plant"]))
    r = O.mpc(p_opt, p_plant, g["x0"], int(g["ticks"]))
    assert np.array_equal(r["iters"], g["n_backward"])
    assert rel_err(r["X_sim"], g["X_sim"]) < TOL and rel_err(r["U_sim"], g["U_sim"], floor=1e-3) < TOL
    assert rel_err(r["costs"], g["costs"]) < TOL
    assert rel_err(r["X_bar"], g["X_bar"]) < TOL and rel_err(r["U_bar"], g["U_bar"], floor=1e-3) < TOL
    assert rel_err(r["K"], g["K_last"]) < TOL


def test_batch_matches_single(oracle):
    from helpers import cfg2_x0, ua_oracle_problem
    O = oracle
    p = ua_oracle_problem(O, 60, maxiter=6)
    x0 = cfg2_x0(5)
    rb = O.optimize_batch(p, x0, np.zeros((5, 1, 60)), nthreads=3)
    for b in range(5):
        r = O.optimize(p, x0[b], np.zeros((1, 60)))
        assert np.array_equal(r["X"], rb["X"][b]) and r["cost"] == rb["cost"][b] and r["iters"] == rb["iters"][b]
