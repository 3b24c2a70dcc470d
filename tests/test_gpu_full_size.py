"""BASELINE.json's configurations at their FULL sizes on one B200, checked through size-independent
properties (the CPU oracle cannot run them in seconds) plus oracle parity on a random sample of members:
  cfg 3  65536 concurrent under-actuated MPC instances, N=200, 8 line-search step sizes
  cfg 4  synthetic LTV n=12, m=4, N=1000, batch 262144 (solved in chunks: the gains alone are 100 GB)
  cfg 5  the per-GPU shard of the 1M-trajectory double-pendulum batch on 8 GPUs: 131072 trajectories, N=500
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from helpers import (cfg2_x0, system_from_golden, ua_oracle_problem, ua_system, member_parity, mpc_member_parity,
                     write_report)

pytestmark = pytest.mark.gpu


def _free_gb():
    import torch
    return torch.cuda.mem_get_info()[0] / 2**30


def test_cfg5_shard_131072_trajectories(oracle):
    import torch
    from class_files.iLQR_class import iLQR
    if _free_gb() < 60:
        pytest.skip("needs 60 GB of free HBM")
    B, N, iters = 131072, 500, 4
    x0 = cfg2_x0(1 << 20, seed=0)[5 * B:6 * B]           # rank 5's contiguous shard of the 1M batch
    sol = iLQR(ua_system(), 5.0, torch.as_tensor(x0).cuda(), torch.zeros((1, N), dtype=torch.float64, device="cuda"),
               tol=0.0, maxiter=iters, verbose=False)
    assert sol._handle.workspace().numel() > 20e9         # lazy schedule: 10 candidate slabs (no A_t, B_t: fused K1+K2)
    sol.enable_trace()
    X, U, cost = sol.optimize_trajectory()
    torch.cuda.synchronize()
    assert sol.total_iterations == int(sol.iterations.sum().item())
    st = sol.status.cpu().numpy()
    assert set(np.unique(st)) <= {1, 2} and (st == 2).mean() > 0.999
    assert torch.equal(X[:, :, 0], torch.as_tensor(x0).cuda())
    assert bool(torch.isfinite(cost).all())
    # X is the rollout of U: re-rolling U open loop (alpha = 0, zero gains) reproduces X and the cost bit for bit
    Xr, Ur, cr = sol.forward_pass(sol.x_0, 0.0, X, U, torch.zeros_like(sol.U_ff), torch.zeros_like(sol.K))
    assert torch.equal(Xr, X) and torch.equal(Ur, U) and torch.equal(cr, cost)
    # oracle parity on a random sample of members, member by member (helpers.member_parity): X, U, K, U_ff, cost and flow
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(B, 64, replace=False))
    pd = torch.as_tensor(pick).cuda()
    idx, tc = sol.trace_arrays()
    got = dict(X=X[pd].cpu().numpy(), U=U[pd].cpu().numpy(), K=sol.K[pd].cpu().numpy(), U_ff=sol.U_ff[pd].cpu().numpy(),
               cost=cost[pd].cpu().numpy(), iters=sol.iterations[pd].cpu().numpy(), status=st[pick], alpha_idx=idx[pick],
               cost_trace=tc[pick, 1:])
    report, failures = member_parity(oracle, ua_oracle_problem(oracle, N, tol=0.0, maxiter=iters), x0[pick], np.zeros((64, 1, N)),
                                     got)
    write_report("cfg5_shard_B131072_N500_it4_sample64", report)
    assert not failures, (failures[:5], report)


def test_cfg4_ltv_262144_chunked():
    import torch
    from class_files.chunked import solve_chunked
    from class_files.systems.ltv_sys import MyLTVSystem
    from test_ltv import riccati
    if _free_gb() < 120:
        pytest.skip("needs 120 GB of free HBM")
    B, N, chunk = 262144, 1000, 32768
    s = MyLTVSystem.synthetic(seed=2)
    rng = np.random.default_rng(2)
    x0 = rng.standard_normal((B, 12))
    phi = rng.uniform(0, 2 * np.pi, B)
    r = solve_chunked(s, N * s.dt, torch.as_tensor(x0).cuda(), torch.zeros((4, N), dtype=torch.float64, device="cuda"),
                      chunk, phi=torch.as_tensor(phi).cuda(), maxiter=2, n_alpha=4)
    cost = r["cost"].cpu().numpy()
    it = r["iterations"].cpu().numpy()
    assert np.all(it == 2) and np.all(np.isfinite(cost))
    assert torch.equal(r["X"][:, :, 0], torch.as_tensor(x0).cuda())
    # linear dynamics + quadratic cost: one full step reaches the LQR optimum 1/2 x0' P0 x0
    # (matlab/CLASSES/Linear_iLQR_CLASS.m:135-139); members of every chunk incl. the first and last trajectory
    pick = np.concatenate([[0, B - 1], np.sort(rng.choice(B, 10, replace=False))])
    for b in pick:
        _, P0 = riccati(s, N, phi[b])
        want = 0.5 * x0[b] @ P0 @ x0[b]
        assert abs(cost[b] - want) < 1e-9 * want, (b, cost[b], want)
    # quadratic scaling: the optimal cost of a linear problem is homogeneous of degree 2 in x0
    r2 = solve_chunked(s, N * s.dt, torch.as_tensor(2.0 * x0[:4096]).cuda(),
                       torch.zeros((4, N), dtype=torch.float64, device="cuda"), 4096,
                       phi=torch.as_tensor(phi[:4096]).cuda(), maxiter=2, n_alpha=4)
    assert rel_err(r2["cost"].cpu().numpy(), 4.0 * cost[:4096]) < 1e-9


def test_cfg3_65536_mpc_instances(oracle):
    import torch
    from class_files.iLQR_class import iLQR
    from class_files.mpc import run_mpc
    if _free_gb() < 30:
        pytest.skip("needs 30 GB of free HBM")
    g = load_golden("mpc_cfg3_ua_T2_ticks2")
    opt = system_from_golden(g)
    plant = system_from_golden(g, integrator="backward_euler")
    B, T, N, ticks = 65536, 2.0, 200, 2
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal((B, 4)) * np.array([0.1, 0.1, 0.5, 0.5])
    x0d = torch.as_tensor(x0).cuda()
    sol = iLQR(opt, T, x0d, torch.zeros((1, N), dtype=torch.float64, device="cuda"), maxiter=50, verbose=False, n_alpha=8)
    r = run_mpc(sol, plant, x0d, ticks)
    X_sim, U_sim = r["X_sim"], r["U_sim"]
    assert X_sim.shape == (B, 4, ticks + 1) and U_sim.shape == (B, 1, ticks)
    assert torch.equal(X_sim[:, :, 0], x0d) and bool(torch.isfinite(X_sim).all())
    # the closed loop is the plant (backward Euler) driven by the applied controls: bit exact
    for k in range(ticks):
        nxt = plant.f_fcn(X_sim[:, :, k].contiguous(), U_sim[:, :, k].contiguous())
        assert torch.equal(nxt, X_sim[:, :, k + 1])
    # oracle parity on a sample of instances (8 tries per line search in both)
    p_opt = oracle.problem_from_golden(g, maxiter=50, n_alpha=8)
    p_plant = oracle.problem_from_golden(g, integrator="backward_euler")
    pick = np.sort(rng.choice(B, 16, replace=False))
    pd = torch.as_tensor(pick).cuda()
    report, failures = mpc_member_parity(oracle, p_opt, p_plant, x0, ticks, X_sim[pd].cpu().numpy(),
                                         r["iterations"][pd].cpu().numpy(), members=pick)
    write_report("cfg3_B65536_N200_ticks2_sample16", report)
    assert not failures, (failures[:5], report)
