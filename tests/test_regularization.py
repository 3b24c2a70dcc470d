"""Levenberg-Marquardt regularisation of Q_uu scheduled on the device -- an EXTENSION (north_star item 4):
the reference has no regularisation and stops a solve whose line search fails (iLQR_class.py:304-307).
Disabled (reg_factor <= 1, the default) nothing changes.  Enabled, a failed iteration is retried with a
larger mu.  The CPU oracle restates the schedule of include/ilqr_b200.h; the GPU must follow it."""
import numpy as np
import pytest

from conftest import rel_err
from helpers import SENS_FACTOR as SF
from helpers import cfg2_x0, ua_oracle_problem, ua_system, rounding_sensitivity

N = 500
FAILING = (571, 3086)        # members of the seeded config-2 batch whose reference solve ends in a failed line search
REG = dict(reg_init=0.0, reg_factor=10.0, reg_min=1e-6, reg_max=1e6)


def _x0(b):
    return cfg2_x0(32768, seed=0)[b]


@pytest.mark.parametrize("b", FAILING)
def test_oracle_schedule_recovers_failed_line_search(oracle, b):
    O = oracle
    U0 = np.zeros((1, N))
    plain = O.optimize(ua_oracle_problem(O, N, maxiter=30), _x0(b), U0)
    assert plain["status"] == "ls_failed"
    reg = O.optimize(ua_oracle_problem(O, N, maxiter=30, **REG), _x0(b), U0)
    k = plain["iters"]
    # identical up to the failure, then retries (index -1) until a regularised step is accepted
    assert np.array_equal(reg["alpha_idx"][:k - 1], plain["alpha_idx"][:k - 1])
    assert reg["alpha_idx"][k - 1] == -1 and reg["iters"] > k
    later = reg["alpha_idx"][k:]
    assert (later >= 0).any() and reg["cost"] < 0.9 * plain["cost"]
    assert np.all(np.diff(reg["cost_trace"]) <= 0)
    # a solve that never fails is bit-identical with the schedule enabled
    ok_plain = O.optimize(ua_oracle_problem(O, 100, maxiter=30), _x0(0), np.zeros((1, 100)))
    ok_reg = O.optimize(ua_oracle_problem(O, 100, maxiter=30, **REG), _x0(0), np.zeros((1, 100)))
    assert ok_plain["status"] == "converged" and ok_reg["mu"] == 0.0
    for key in ("X", "U", "K", "cost"):
        assert np.array_equal(ok_plain[key], ok_reg[key])


def test_oracle_backward_pass_mu(oracle):
    """Q_uu + mu I: with a huge mu the gains vanish like 1/mu"""
    O = oracle
    p = ua_oracle_problem(O, 50)
    rng = np.random.default_rng(0)
    X, U = rng.standard_normal((4, 51)), rng.standard_normal((1, 50))
    import ctypes as C
    k0, K0 = O.backward_pass(p, X, U)
    k1, K1 = np.empty((1, 50)), np.empty((50, 1, 4))
    O.lib().orc_backward_pass_mu(C.byref(p), 0.0, 0.0, O._d(X), O._d(U), O._d(k1), O._d(K1))
    assert np.array_equal(K0, K1) and np.array_equal(k0, k1)
    O.lib().orc_backward_pass_mu(C.byref(p), 0.0, 1e12, O._d(X), O._d(U), O._d(k1), O._d(K1))
    assert np.max(np.abs(K1)) < 1e-6 * np.max(np.abs(K0))


@pytest.mark.gpu
@pytest.mark.parametrize("b", FAILING)
def test_gpu_follows_the_schedule(oracle, b):
    from class_files.iLQR_class import iLQR
    O = oracle
    maxiter = 16
    x0 = _x0(b)
    sol = iLQR(ua_system(), 5.0, x0, np.zeros((1, N)), maxiter=maxiter, verbose=True, **REG)
    X, U, cost = sol.optimize_trajectory()
    idx, alphas, costs = sol.trace(0)
    p = ua_oracle_problem(O, N, maxiter=maxiter, **REG)
    ref = O.optimize(p, x0, np.zeros((1, N)))
    sens = rounding_sensitivity(O, p, x0)
    k = min(sens["stable_prefix"], len(idx), ref["iters"])
    assert (ref["alpha_idx"][:k] < 0).any(), "the compared prefix must contain the failure and a retry"
    assert np.array_equal(idx[:k], ref["alpha_idx"][:k]), (idx, ref["alpha_idx"])
    err = np.abs(costs[1:k + 1] - ref["cost_trace"][:k]) / np.abs(ref["cost_trace"][:k])
    assert np.all(err <= np.maximum(1e-9, SF * sens["cost"][:k])), (err, sens["cost"][:k])
    if sens["flow_stable"]:
        assert int(sol.iterations) == ref["iters"]
        assert rel_err(sol.mu, ref["mu"], floor=1e-12) < 1e-12
        assert rel_err(cost, ref["cost"]) <= max(1e-9, SF * sens["cost"][-1])


@pytest.mark.gpu
def test_gpu_schedule_in_a_batch_and_disabled_is_exact(oracle):
    """batch containing the two failing members: per-trajectory mu, eager and lazy line-search schedules
    agree bit for bit; members that never fail are untouched by the enabled schedule."""
    from class_files.iLQR_class import iLQR
    B = 64
    x0 = cfg2_x0(32768, seed=0)[:B].copy()
    x0[5], x0[40] = _x0(FAILING[0]), _x0(FAILING[1])
    out = {}
    for name, kw, waves in (("plain", {}, ()), ("reg", REG, ()), ("reg_lazy", REG, (2, 2, 2, 4))):
        sol = iLQR(ua_system(), 5.0, x0, np.zeros((1, N)), maxiter=14, verbose=False, **kw)
        sol.set_linesearch_waves(waves)
        X, U, cost = sol.optimize_trajectory()
        out[name] = dict(X=X.copy(), cost=cost.copy(), status=sol.status.copy(), it=sol.iterations.copy(),
                         mu=None if sol.mu is None else sol.mu.copy())
    assert out["plain"]["mu"] is None
    failed = out["plain"]["status"] == 1
    assert failed[5] and failed[40]
    assert np.array_equal(out["reg"]["X"][~failed], out["plain"]["X"][~failed])
    assert np.array_equal(out["reg"]["cost"][~failed], out["plain"]["cost"][~failed])
    assert np.all(out["reg"]["status"][failed] != 1) and np.all(out["reg"]["it"][failed] > out["plain"]["it"][failed])
    assert np.all(out["reg"]["cost"][failed] <= out["plain"]["cost"][failed])
    for key in ("X", "cost", "status", "it", "mu"):
        assert np.array_equal(out["reg"][key], out["reg_lazy"][key]), key
