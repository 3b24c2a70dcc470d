"""Shared by the CPU and GPU tests: build reference-API systems / oracle problems from golden files."""
import numpy as np

PHYS_DP = ("g", "m1", "m2", "l1", "l2", "d1", "d2", "theta1", "theta2")


def system_from_golden(g, integrator=None, dtype="float64"):
    """Instantiate the drop-in System subclass for a golden file's parameter block."""
    from class_files.systems.pendulum_sys import MyPendulum
    from class_files.systems.double_pendulum_sys import MyDoublePendulum
    from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum
    kind = str(g["p_kind"])
    integ = integrator or str(g["p_integrator"])
    full = lambda v: np.diag(v) if np.ndim(v) == 1 else np.array(v)      # weights stored as diagonals or full matrices
    common = dict(dt=float(g["p_dt"]), x_target=np.array(g["p_x_target"]), Q=full(g["p_Q"]), R=full(g["p_R"]),
                  Q_f=full(g["p_Q_f"]), integrator=integ, dtype=dtype)
    if kind == "pendulum":
        return MyPendulum(g=float(g["p_g"]), l=float(g["p_l"]), d=float(g["p_d"]), **common)
    phys = {k: float(g["p_" + k]) for k in PHYS_DP}
    cls = MyDoublePendulum if kind == "double" else MyUADoublePendulum
    return cls(**phys, **common)


def cfg2_x0(count, seed=0):
    """BASELINE config 2 initial states (same generator as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    x0 = np.empty((count, 4))
    x0[:, :2] = rng.uniform(-np.pi, np.pi, size=(count, 2))
    x0[:, 2:] = rng.uniform(-2.0, 2.0, size=(count, 2))
    return x0


# parity bound = max(1e-9, SENS_FACTOR x measured rounding sensitivity); see rounding_sensitivity()
SENS_FACTOR = 30.0

UA_OL = dict(dt=0.01, Q=[1.0, 1.0, 0.1, 0.1], R=[1.0], Q_f=[1000.0, 1000.0, 100.0, 100.0],
             x_target=[np.pi, 0.0, 0.0, 0.0],
             phys=dict(g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.1, d2=0.1, theta1=1.0 / 12, theta2=1.0 / 12))


def ua_system(integrator="rk4", dtype="float64", **over):
    from class_files.systems.UA_double_pendulum_sys import MyUADoublePendulum
    p = dict(UA_OL)
    p.update(over)
    return MyUADoublePendulum(dt=p["dt"], x_target=np.array(p["x_target"]), Q=np.diag(p["Q"]), R=np.diag(p["R"]),
                              Q_f=np.diag(p["Q_f"]), integrator=integrator, dtype=dtype, **p["phys"])


def ua_oracle_problem(O, N, integrator="rk4", **kw):
    p = UA_OL
    return O.make_problem("ua", integrator, N, p["dt"], p["Q"], p["R"], p["Q_f"], p["x_target"], p["phys"], **kw)


def golden_flow(g):
    """Per-iteration (accepted try index or -1, cost after the iteration) from the call trace that
    make_golden.py recorded around the reference's forward_pass (first call = alpha-0 rollout)."""
    ta, tc = np.asarray(g["trace_alpha"]), np.asarray(g["trace_cost"])
    cost = tc[0]
    idx, costs = [], [cost]
    i = 1
    while i < len(ta):
        j, acc = 0, -1
        while i < len(ta):
            if j > 0 and ta[i] == 1.0:
                break                       # next iteration's first try
            if tc[i] <= cost:
                acc, cost = j, tc[i]
                i += 1
                break
            i += 1
            j += 1
        idx.append(acc)
        costs.append(cost)
    return np.array(idx), np.array(costs)


def rounding_sensitivity(O, p, x0, n_draws=3, eps=1e-14, seed=7):
    """How far rounding-level input noise moves a full iLQR solve of this problem.

    iLQR on the pendulum swing-ups amplifies tiny differences from iteration to iteration (the accept
    tests are discontinuous and the rollouts are chaotic), so two correct float64 implementations drift
    apart.  This runs the CPU oracle on x0, on `n_draws` copies of x0 perturbed by `eps` (relative,
    absolute for zero entries) and once from its FMA-contracted build (same inputs, different rounding in
    every operation), and returns the largest drift seen: per-iteration relative cost drift, final
    X/U/K/U_ff drift, and whether the control flow (accepted step sizes) stayed the same.
    Parity tests accept max(1e-9, SENS_FACTOR x this): 1e-9 wherever the computation is well conditioned."""
    rng = np.random.default_rng(seed)
    x0 = np.asarray(x0, dtype=np.float64)
    U0 = np.zeros((p.m, p.N))
    base = O.optimize(p, x0, U0)
    nb = base["iters"]
    out = dict(cost=np.zeros(nb), X=0.0, U=0.0, K=0.0, U_ff=0.0, flow_stable=True, stable_prefix=nb)
    for draw in range(n_draws + 1):
        if draw == n_draws:     # same inputs, every operation rounded differently (FMA-contracted build)
            with O.rounding_variant():
                r = O.optimize(p, x0, U0)
        else:
            d = rng.choice([-1.0, 1.0], size=x0.shape) * eps
            xp = np.where(x0 != 0.0, x0 * (1.0 + d), d)
            r = O.optimize(p, xp, U0)
        n = min(nb, r["iters"])
        same = r["alpha_idx"][:n] == base["alpha_idx"][:n]
        k = n if same.all() else int(np.argmin(same))
        if k < nb or r["iters"] != nb:
            out["flow_stable"] = False
        out["stable_prefix"] = min(out["stable_prefix"], k)
        dc = np.abs(r["cost_trace"][:n] - base["cost_trace"][:n]) / np.abs(base["cost_trace"][:n])
        out["cost"][:n] = np.maximum(out["cost"][:n], dc)
        for key, floor in (("X", 0.0), ("U", 1e-3), ("K", 0.0), ("U_ff", 1e-3)):
            e = float(np.max(np.abs(r[key] - base[key])) / max(float(np.max(np.abs(base[key]))), floor))
            out[key] = max(out[key], e)
    out["cost"] = np.maximum.accumulate(out["cost"])
    return out


def _perturb(rng, a, eps):
    a = np.asarray(a, dtype=np.float64)
    d = rng.choice([-1.0, 1.0], size=a.shape) * eps
    return np.where(a != 0.0, a * (1.0 + d), d)


def backward_sensitivity(O, p, X, U, n_draws=3, eps=1e-14, seed=11):
    """Drift of (K, U_ff) of one backward pass under 1e-14 input noise: the Riccati recursion of the
    stiff problems (Q_f = 1000, R dt = 1e-3) subtracts nearly equal matrices at every step, so even a
    single pass on identical inputs amplifies rounding differences between correct implementations."""
    rng = np.random.default_rng(seed)
    U_ff, K = O.backward_pass(p, X, U)
    sK = sU = 0.0
    for d in range(n_draws + 1):
        if d == n_draws:        # same inputs, every operation rounded differently (FMA-contracted build)
            with O.rounding_variant():
                u2, k2 = O.backward_pass(p, X, U)
        else:
            u2, k2 = O.backward_pass(p, _perturb(rng, X, eps), _perturb(rng, U, eps))
        sK = max(sK, float(np.max(np.abs(k2 - K)) / np.max(np.abs(K))))
        sU = max(sU, float(np.max(np.abs(u2 - U_ff)) / max(float(np.max(np.abs(U_ff))), 1e-6)))
    return sK, sU


def forward_sensitivity(O, p, x0, alpha, X, U, U_ff, K, n_draws=3, eps=1e-14, seed=13):
    """Drift of (X_new, U_new, cost) of one rollout under 1e-14 noise on x0 and the gains."""
    rng = np.random.default_rng(seed)
    Xn, Un, c = O.forward_pass(p, x0, alpha, X, U, U_ff, K)
    sX = sU = sc = 0.0
    for d in range(n_draws + 1):
        if d == n_draws:
            with O.rounding_variant():
                X2, U2, c2 = O.forward_pass(p, x0, alpha, X, U, U_ff, K)
        else:
            X2, U2, c2 = O.forward_pass(p, _perturb(rng, x0, eps), alpha, X, U, _perturb(rng, U_ff, eps),
                                        _perturb(rng, K, eps))
        sX = max(sX, float(np.max(np.abs(X2 - Xn)) / np.max(np.abs(Xn))))
        sU = max(sU, float(np.max(np.abs(U2 - Un)) / max(float(np.max(np.abs(Un))), 1e-3)))
        sc = max(sc, abs(c2 - c) / abs(c))
    return sX, sU, sc
