"""GPU parity tests: the CUDA path (through the reference-style Python API -> ctypes -> C ABI)
against (a) golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded
inputs.  Tolerances are FP64 relative errors; BASELINE.json asks for <= 1e-9 on trajectories,
gains and cost."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from helpers import system_from_golden, cfg2_x0, ua_system, ua_oracle_problem, golden_flow

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.mark.parametrize("name", golden_names("derivs_"))
def test_point_functions_vs_reference(name):
    g = load_golden(name)
    s = system_from_golden(g)
    xs, us = g["xs"], g["us"]
    got = dict(f=s.f_fcn(xs, us), f_x=s.f_x_fcn(xs, us), f_u=s.f_u_fcn(xs, us), l=s.l_fcn(xs, us),
               l_x=s.l_x_fcn(xs, us), l_u=s.l_u_fcn(xs, us), l_xx=s.l_xx_fcn(xs, us), l_uu=s.l_uu_fcn(xs, us),
               l_ux=s.l_ux_fcn(xs, us), l_f=s.l_f_fcn(xs), l_f_x=s.l_f_x_fcn(xs), l_f_xx=s.l_f_xx_fcn(xs))
    for k, v in got.items():
        assert v.shape == g[k].shape, (k, v.shape, g[k].shape)
        assert rel_err(v, g[k], floor=1e-3) < 1e-12, (k, rel_err(v, g[k], floor=1e-3))
    # single-point call, reference shapes
    A = s.f_x_fcn(xs[0], us[0])
    assert A.shape == (s.n_x, s.n_x) and rel_err(A, g["f_x"][0]) < 1e-12
    assert np.ndim(s.l_fcn(xs[0], us[0])) == 0


@pytest.mark.parametrize("name", golden_names("passes_"))
def test_backward_forward_pass_vs_reference(name):
    from class_files.iLQR_class import iLQR
    g = load_golden(name)
    s = system_from_golden(g)
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((s.n_u, int(g["N"]))), verbose=False)
    assert sol.N == int(g["N"])
    X0, U0, c0 = sol.forward_pass(g["x0"], 0.0, sol.X, g["U_nom"], sol.U_ff, sol.K)
    assert rel_err(X0, g["X_nom"]) < 1e-12 and rel_err(c0, g["cost0"]) < 1e-12
    U_ff, K = sol.backward_pass(g["X_nom"], g["U_nom"])
    assert U_ff.shape == g["U_ff"].shape and K.shape == g["K"].shape
    assert rel_err(U_ff, g["U_ff"]) < TOL, rel_err(U_ff, g["U_ff"])
    assert rel_err(K, g["K"]) < TOL, rel_err(K, g["K"])
    for tag, a in (("1p0", 1.0), ("0p5", 0.5), ("0p125", 0.125)):
        Xn, Un, c = sol.forward_pass(g["x0_b"], a, g["X_nom"], g["U_nom"], g["U_ff"], g["K"])
        assert rel_err(Xn, g["X_a" + tag]) < TOL
        assert rel_err(Un, g["U_a" + tag]) < TOL
        assert rel_err(c, g["cost_a" + tag]) < TOL


@pytest.mark.parametrize("name", golden_names("solve_"))
def test_optimize_trajectory_vs_reference(name, capsys):
    from class_files.iLQR_class import iLQR
    g = load_golden(name)
    s = system_from_golden(g)
    N = int(g["N"])
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((s.n_u, N)), tol=float(g["tol"]), maxiter=int(g["maxiter"]),
               verbose=True)
    X, U, cost = sol.optimize_trajectory()
    out = capsys.readouterr().out
    assert out.startswith("Initial cost:")
    assert X.shape == g["X"].shape and U.shape == g["U"].shape
    # same control flow as the reference: iterations and the accepted step size of every iteration
    assert int(sol.iterations) == int(g["n_backward"])
    idx, alphas, costs = sol.trace(0)
    ref_idx, ref_costs = golden_flow(g)
    assert np.array_equal(idx, ref_idx), (idx, ref_idx)
    assert rel_err(costs, ref_costs) < TOL
    assert rel_err(cost, g["cost"]) < TOL
    assert rel_err(X, g["X"]) < TOL, rel_err(X, g["X"])
    assert rel_err(U, g["U"], floor=1e-3) < TOL, rel_err(U, g["U"], floor=1e-3)
    assert rel_err(sol.K, g["K"]) < TOL, rel_err(sol.K, g["K"])
    assert rel_err(sol.U_ff, g["U_ff"], floor=1e-3) < TOL


def test_batched_solve_vs_oracle(oracle):
    """512 seeded config-2 trajectories at N=100: every trajectory's flow and result match the oracle."""
    from class_files.iLQR_class import iLQR
    B, N = 512, 100
    x0 = cfg2_x0(B)
    sol = iLQR(ua_system(), 1.0, x0, np.zeros((1, N)), maxiter=20, verbose=False)
    X, U, cost = sol.optimize_trajectory()
    ref = oracle.optimize_batch(ua_oracle_problem(oracle, N, maxiter=20), x0, np.zeros((B, 1, N)))
    same = sol.iterations == ref["iters"]
    assert same.mean() > 0.99, same.mean()          # a rounding-level tie in an accept test may flip a branch
    assert np.array_equal(sol.status[same], ref["status"][same])
    ex = np.max(np.abs(X - ref["X"]), axis=(1, 2)) / np.max(np.abs(ref["X"]), axis=(1, 2))
    ec = np.abs(cost - ref["cost"]) / np.abs(ref["cost"])
    ek = np.max(np.abs(sol.K - ref["K"]), axis=(1, 2, 3)) / np.max(np.abs(ref["K"]), axis=(1, 2, 3))
    assert np.quantile(ex[same], 0.99) < TOL and np.quantile(ec[same], 0.99) < TOL and np.quantile(ek[same], 0.99) < 1e-8
    assert int(sol.total_iterations) == int(sol.iterations.sum())


def test_fp32_mode(oracle):
    """optional FP32 mode: 1e-4 on a single rollout + backward pass (BASELINE.json north_star)."""
    from class_files.iLQR_class import iLQR
    g = load_golden("passes_ua_rk4")
    s = system_from_golden(g, dtype="float32")
    sol = iLQR(s, float(g["T"]), g["x0"], np.zeros((s.n_u, int(g["N"]))), verbose=False)
    U_ff, K = sol.backward_pass(g["X_nom"], g["U_nom"])
    assert rel_err(K, g["K"]) < 1e-4 and rel_err(U_ff, g["U_ff"]) < 1e-4
    Xn, Un, c = sol.forward_pass(g["x0_b"], 0.5, g["X_nom"], g["U_nom"], g["U_ff"], g["K"])
    assert rel_err(Xn, g["X_a0p5"]) < 1e-4 and rel_err(c, g["cost_a0p5"]) < 1e-4


def test_torch_in_torch_out():
    import torch
    from class_files.iLQR_class import iLQR
    B, N = 64, 50
    x0 = torch.as_tensor(cfg2_x0(B)).cuda()
    sol = iLQR(ua_system(), 0.5, x0, torch.zeros((1, N), dtype=torch.float64, device="cuda"), maxiter=3, verbose=False)
    X, U, cost = sol.optimize_trajectory()
    assert isinstance(X, torch.Tensor) and X.is_cuda and X.shape == (B, 4, N + 1)
    assert U.shape == (B, 1, N) and cost.shape == (B,) and sol.K.shape == (B, N, 1, 4)
    assert torch.allclose(X[:, :, 0], x0)


def test_errors():
    from class_files.iLQR_class import iLQR
    s = ua_system()
    with pytest.raises(ValueError, match="U_init must have shape"):
        iLQR(s, 1.0, np.zeros(4), np.zeros((1, 99)))
    with pytest.raises(ValueError, match="Unknown integrator"):
        ua_system(integrator="rk5")
