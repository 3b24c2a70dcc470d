"""Config 4: synthetic LTV system (n=12, m=4).  The reference has no such system; the anchor is the
closed-form answer for linear dynamics + quadratic cost (matlab/CLASSES/Linear_iLQR_CLASS.m:80-139,
matlab/main_.m:7-32): the backward pass is the discrete Riccati recursion and one alpha=1 iteration
reaches the optimum, whose cost is 1/2 x0' P0 x0.  CPU: the oracle against that closed form.
GPU: the CUDA path against the oracle and the same closed form."""
import numpy as np
import pytest

from conftest import rel_err

N_LTV = 120


def ltv_setup(B=1, seed=2):
    from class_files.systems.ltv_sys import MyLTVSystem
    s = MyLTVSystem.synthetic(seed=seed)
    rng = np.random.default_rng(seed + 100)
    x0 = rng.standard_normal((B, 12))
    phi = rng.uniform(0, 2 * np.pi, B)
    return s, x0, phi


def oracle_problem(O, s, N, **kw):
    return O.make_problem("ltv", "euler", N, s.dt, np.eye(12), 0.1 * np.eye(4), 10.0 * np.eye(12), np.zeros(12),
                          ltv=dict(Ac=s.Ac, E=s.E, Bc=s.Bc, amp=s.amp), **kw)


def riccati(s, N, phi):
    """textbook finite-horizon LQR on x+ = A_t x + B_t u with stage cost dt*(x'Qx/2 + u'Ru/2)"""
    dt = s.dt
    Q, R, P = np.eye(12) * dt, 0.1 * np.eye(4) * dt, 10.0 * np.eye(12)
    Ks = np.zeros((N, 4, 12))
    Bt = dt * s.Bc
    for t in range(N - 1, -1, -1):
        At = np.eye(12) + dt * (s.Ac + s.amp * np.sin(2 * np.pi * t / N + phi) * s.E)
        G = R + Bt.T @ P @ Bt
        K = -np.linalg.solve(G, Bt.T @ P @ At)
        Ks[t] = K
        P = Q + At.T @ P @ At + At.T @ P @ Bt @ K
        P = 0.5 * (P + P.T)
    return Ks, P


def test_oracle_ltv_is_lqr(oracle):
    O = oracle
    s, x0, phi = ltv_setup(B=3)
    p = oracle_problem(O, s, N_LTV, maxiter=3)
    for b in range(3):
        Ks, P0 = riccati(s, N_LTV, phi[b])
        r = O.optimize(p, x0[b], np.zeros((4, N_LTV)), phi=float(phi[b]))
        assert rel_err(r["K"], Ks) < 1e-9
        assert r["alpha_idx"][0] == 0                      # full step accepted
        assert abs(r["cost_trace"][0] - 0.5 * x0[b] @ P0 @ x0[b]) < 1e-9 * r["cost_trace"][0]
        assert r["status"] == "converged" and r["iters"] == 2    # second iteration cannot improve


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", ["16", "4", "32"])
def test_ltv_point_functions_and_passes_vs_oracle(oracle, monkeypatch, lanes):
    """the three Riccati kernels of the LTV model: sixteen lanes per trajectory (small batches), the register-tiled
    four-lane kernel (FP32, large batches) and the FP64 tensor-core kernel, one warp per trajectory (FP64, large
    batches); forced here by ILQR_LTV_LANES"""
    from class_files.iLQR_class import iLQR
    monkeypatch.setenv("ILQR_LTV_LANES", lanes)
    O = oracle
    B = 5
    s, x0, phi = ltv_setup(B=B)
    p = oracle_problem(O, s, N_LTV)
    rng = np.random.default_rng(5)
    us = rng.standard_normal((B, 4))
    got = s.f_fcn(x0, us, t=7, phi=phi, N=N_LTV)
    for b in range(B):
        assert rel_err(got[b], O.f(p, x0[b], us[b], t=7, phi=float(phi[b]))) < 1e-13
    sol = iLQR(s, N_LTV * s.dt, x0, np.zeros((4, N_LTV)), verbose=False, phi=phi)
    assert sol.N == N_LTV
    U_nom = rng.standard_normal((B, 4, N_LTV)) * 0.3
    X_nom, U2, c0 = sol.forward_pass(x0, 0.0, sol.X, U_nom, sol.U_ff, sol.K)
    U_ff, K = sol.backward_pass(X_nom, U_nom)
    Xn, Un, c = sol.forward_pass(x0, 0.5, X_nom, U_nom, U_ff, K)
    for b in range(B):
        Xo, Uo, co = O.forward_pass(p, x0[b], 0.0, np.zeros((12, N_LTV + 1)), U_nom[b], np.zeros((4, N_LTV)),
                                    np.zeros((N_LTV, 4, 12)), phi=float(phi[b]))
        assert rel_err(X_nom[b], Xo) < 1e-12 and rel_err(c0[b], co) < 1e-12
        Uff_o, K_o = O.backward_pass(p, Xo, U_nom[b], phi=float(phi[b]))
        assert rel_err(K[b], K_o) < 1e-9 and rel_err(U_ff[b], Uff_o) < 1e-9
        X2, U2o, c2 = O.forward_pass(p, x0[b], 0.5, Xo, U_nom[b], Uff_o, K_o, phi=float(phi[b]))
        assert rel_err(Xn[b], X2) < 1e-9 and rel_err(c[b], c2) < 1e-9
        Ks, _ = riccati(s, N_LTV, phi[b])
        assert rel_err(K[b], Ks) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("lanes,B", [("16", 64), ("4", 64), ("4", 37), ("32", 64), ("32", 37)])
def test_ltv_solve_reaches_lqr_optimum(oracle, monkeypatch, lanes, B):
    from class_files.iLQR_class import iLQR
    monkeypatch.setenv("ILQR_LTV_LANES", lanes)
    s, x0, phi = ltv_setup(B=B)
    sol = iLQR(s, N_LTV * s.dt, x0, np.zeros((4, N_LTV)), verbose=False, phi=phi, maxiter=5)
    X, U, cost = sol.optimize_trajectory()
    # the second iteration starts at the optimum: whether its line search "improves" the cost is a
    # rounding-level tie, so it ends either converged (0) or line-search-failed (1) with the same result
    assert np.all(sol.iterations == 2) and np.all(sol.status <= 1) and np.mean(sol.status == 0) > 0.75
    p = oracle_problem(oracle, s, N_LTV, maxiter=5)
    ref = oracle.optimize_batch(p, x0, np.zeros((B, 4, N_LTV)), phi=phi)
    assert rel_err(cost, ref["cost"]) < 1e-9 and rel_err(X, ref["X"]) < 1e-9 and rel_err(U, ref["U"]) < 1e-9
    for b in range(0, B, 16):
        _, P0 = riccati(s, N_LTV, phi[b])
        assert abs(cost[b] - 0.5 * x0[b] @ P0 @ x0[b]) < 1e-9 * cost[b]


@pytest.mark.gpu
def test_ltv_riccati_kernels_agree_incl_fp32_and_ragged_batches(monkeypatch):
    """the LTV Riccati kernels on the same nominal: FP64 gains of the lane-tiled kernels to 1e-12, of the tensor-core
    kernel (every block size) to 1e-11 (DMMA sums in another order), FP32 gains to 1e-4 of the FP64 ones, on a batch that
    fills no kernel's last block"""
    from class_files.iLQR_class import iLQR
    from class_files.systems.ltv_sys import MyLTVSystem
    B, N = 203, 90
    rng = np.random.default_rng(9)
    x0, phi = rng.standard_normal((B, 12)), rng.uniform(0, 2 * np.pi, B)
    U_nom = 0.3 * rng.standard_normal((B, 4, N))
    out = {}
    for dtype in ("float64", "float32"):
        s = MyLTVSystem.synthetic(seed=2, dtype=dtype)
        for lanes, wpb in (("16", ""), ("4", ""), ("32", "4"), ("32", "8")):
            if dtype == "float32" and lanes == "32":
                continue                      # FP64 only: the FP32 mode keeps the lane-tiled kernels
            monkeypatch.setenv("ILQR_LTV_LANES", lanes)
            monkeypatch.setenv("ILQR_LTV_MMA_WPB", wpb)
            sol = iLQR(s, N * s.dt, x0, np.zeros((4, N)), verbose=False, phi=phi)
            X_nom, _, _ = sol.forward_pass(x0, 0.0, sol.X, U_nom, sol.U_ff, sol.K)
            U_ff, K = sol.backward_pass(X_nom, U_nom)
            out[dtype, lanes + wpb] = (np.asarray(K, dtype=np.float64), np.asarray(U_ff, dtype=np.float64))
    for j in range(2):
        assert rel_err(out["float64", "4"][j], out["float64", "16"][j]) < 1e-12
        for wpb in ("4", "8"):
            assert rel_err(out["float64", "32" + wpb][j], out["float64", "16"][j]) < 1e-11
        assert rel_err(out["float32", "4"][j], out["float64", "16"][j], floor=1e-3) < 1e-4
        assert rel_err(out["float32", "16"][j], out["float64", "16"][j], floor=1e-3) < 1e-4


@pytest.mark.gpu
def test_ltv_dense_weights_target_and_regularisation_across_kernels(oracle, monkeypatch):
    """the paths of the LTV Riccati kernels that the diagonal-weight cases above never take: dense non-symmetric Q, R, Q_f
    with a non-zero target (general cost gradient), on all three kernels against the oracle; and a regularised solve
    (mu on the diagonal of Q_uu) on the tensor-core kernel against the sixteen-lane one"""
    from class_files.iLQR_class import iLQR
    from class_files.systems.ltv_sys import MyLTVSystem
    O = oracle
    B, N = 21, 70
    base = MyLTVSystem.synthetic(seed=4)
    rng = np.random.default_rng(11)
    Q = np.eye(12) + 0.2 * rng.standard_normal((12, 12))
    R = 0.1 * np.eye(4) + 0.02 * rng.standard_normal((4, 4))
    Qf = 10.0 * np.eye(12) + rng.standard_normal((12, 12))
    xt = 0.3 * rng.standard_normal(12)
    s = MyLTVSystem(base.dt, xt, Q, R, Qf, base.Ac, base.E, base.Bc, amp=base.amp)
    x0, phi = rng.standard_normal((B, 12)), rng.uniform(0, 2 * np.pi, B)
    U_nom = 0.3 * rng.standard_normal((B, 4, N))
    p = O.make_problem("ltv", "euler", N, s.dt, Q, R, Qf, xt, ltv=dict(Ac=s.Ac, E=s.E, Bc=s.Bc, amp=s.amp))
    got = {}
    for lanes in ("16", "4", "32"):
        monkeypatch.setenv("ILQR_LTV_LANES", lanes)
        sol = iLQR(s, N * s.dt, x0, np.zeros((4, N)), verbose=False, phi=phi)
        X_nom, _, _ = sol.forward_pass(x0, 0.0, sol.X, U_nom, sol.U_ff, sol.K)
        U_ff, K = sol.backward_pass(X_nom, U_nom)
        got[lanes] = (np.asarray(X_nom), np.asarray(U_ff), np.asarray(K))
    for b in range(0, B, 5):
        Uff_o, K_o = O.backward_pass(p, got["16"][0][b], U_nom[b], phi=float(phi[b]))
        for lanes in got:
            assert rel_err(got[lanes][2][b], K_o) < 1e-9 and rel_err(got[lanes][1][b], Uff_o) < 1e-9, lanes
    # regularised solves: the mu schedule lives in the select kernels, the kernels only add mu to Q_uu's diagonal
    res = {}
    for lanes in ("16", "32"):
        monkeypatch.setenv("ILQR_LTV_LANES", lanes)
        sol = iLQR(s, N * s.dt, x0, np.zeros((4, N)), verbose=False, phi=phi, maxiter=3, reg_init=0.05, reg_factor=4.0)
        X, U, cost = sol.optimize_trajectory()
        res[lanes] = (np.asarray(X), np.asarray(U), np.asarray(cost), np.asarray(sol.K), np.asarray(sol.iterations))
    assert np.array_equal(res["16"][4], res["32"][4])
    for j in range(4):
        assert rel_err(res["32"][j], res["16"][j]) < 1e-9
