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


def backward_sensitivity_per_step(O, p, X, U, n_draws=3, eps=1e-14, seed=11):
    """backward_sensitivity resolved in time: drift of K[t] (relative to max|K[t]|) under the same noise draws.  The
    recursion runs from t = N-1 down, so an ill-conditioned pass is still well determined over the late steps."""
    rng = np.random.default_rng(seed)
    U_ff, K = O.backward_pass(p, X, U)
    K = np.asarray(K)
    N = K.shape[0]
    scale = np.maximum(np.max(np.abs(K.reshape(N, -1)), axis=1), 1e-300)
    s = np.zeros(N)
    for d in range(n_draws + 1):
        if d == n_draws:
            with O.rounding_variant():
                _, k2 = O.backward_pass(p, X, U)
        else:
            _, k2 = O.backward_pass(p, _perturb(rng, X, eps), _perturb(rng, U, eps))
        s = np.maximum(s, np.max(np.abs(np.asarray(k2) - K).reshape(N, -1), axis=1) / scale)
    return s, scale


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


# ----------------------------------------------------------------------------------------------------------
# Per-member parity of a whole batch (no quantiles, no unbounded tail): every member is either at
# max(1e-9, SENS_FACTOR x its own rounding sensitivity) on X, U, K, U_ff and cost with the oracle's control flow, or
# its control flow differs from the oracle's at an iteration where the oracle's own flow flips under rounding noise.
# ----------------------------------------------------------------------------------------------------------
KEYS = (("X", 0.0), ("U", 1e-3), ("K", 0.0), ("U_ff", 1e-3))


def member_rel_err(a, b, floor=0.0):
    """per member: max |a-b| / max(max |b|, floor); arrays are batch-major"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ax = tuple(range(1, a.ndim))
    num = np.max(np.abs(a - b), axis=ax) if ax else np.abs(a - b)
    den = np.max(np.abs(b), axis=ax) if ax else np.abs(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = num / np.maximum(np.maximum(den, floor), 1e-300)
    return np.where(np.isnan(num) & np.isnan(den), 0.0, e)       # both NaN (a diverged member): same outcome


def flow_prefix(idx_a, it_a, idx_b, it_b):
    """per member: number of leading iterations on which two runs made the same line-search decision; equals
    max(it_a, it_b) when the flows are identical"""
    B, M = idx_a.shape
    col = np.arange(M)[None, :]
    a = np.where(col < it_a[:, None], idx_a, -2)
    b = np.where(col < it_b[:, None], idx_b, -2)
    diff = a != b
    first = np.where(diff.any(axis=1), diff.argmax(axis=1), np.maximum(it_a, it_b))
    return first.astype(np.int64)


def batch_sensitivity(O, p, x0, U0, n_draws=4, eps=1e-14, seed=7, phi=None, base=None):
    """rounding_sensitivity() for every member of a batch: the oracle on `n_draws` copies of x0 perturbed by eps
    (relative) and once from its FMA-contracted build.  Returns the base run and, per member, the largest relative
    drift of X, U, K, U_ff, final cost and of the cost after every iteration, and `prefix`: the number of leading
    iterations whose line-search decision no draw changed."""
    rng = np.random.default_rng(seed)
    x0 = np.asarray(x0, dtype=np.float64)
    B = x0.shape[0]
    if base is None:
        base = O.optimize_batch(p, x0, U0, phi=phi, trace=True)
    M = base["alpha_idx"].shape[1]
    s = {k: np.zeros(B) for k, _ in KEYS}
    s["cost"] = np.zeros(B)
    s["cost_it"] = np.zeros((B, M))
    s["prefix"] = base["iters"].astype(np.int64).copy()
    for draw in range(n_draws + 1):
        if draw == n_draws:
            with O.rounding_variant():
                r = O.optimize_batch(p, x0, U0, phi=phi, trace=True)
        else:
            r = O.optimize_batch(p, _perturb(rng, x0, eps), U0, phi=phi, trace=True)
        pre = flow_prefix(r["alpha_idx"], r["iters"], base["alpha_idx"], base["iters"])
        pre = np.where((r["iters"] == base["iters"]) & (r["status"] == base["status"]) & (pre >= base["iters"]),
                       base["iters"], np.minimum(pre, base["iters"]))
        s["prefix"] = np.minimum(s["prefix"], pre)
        for k, floor in KEYS:
            s[k] = np.maximum(s[k], member_rel_err(r[k], base[k], floor))
        s["cost"] = np.maximum(s["cost"], member_rel_err(r["cost"], base["cost"]))
        with np.errstate(invalid="ignore", divide="ignore"):
            d = np.abs(r["cost_trace"] - base["cost_trace"]) / np.abs(base["cost_trace"])
        s["cost_it"] = np.maximum(s["cost_it"], np.nan_to_num(d, nan=0.0))
    s["cost_it"] = np.maximum.accumulate(s["cost_it"], axis=1)
    return base, s


def member_parity(O, p, x0, U0, got, n_draws=4, deep_draws=24, tol=1e-9, phi=None, base=None):
    """Compare a GPU batch solve `got` (dict X, U, K, U_ff, cost, iters, status, alpha_idx (B,maxiter), cost_trace
    (B,maxiter) = cost after each iteration) with the oracle member by member.  Returns (report, failures):
    `failures` lists every member that is neither within max(tol, SENS_FACTOR x its own sensitivity) on all of X, U, K,
    U_ff, cost with the oracle's control flow, nor explained by a control-flow flip the oracle itself shows under
    1e-14 input noise / FMA contraction at or before the same iteration (members whose flip is not reproduced by
    the first `n_draws` draws get `deep_draws` more)."""
    x0 = np.asarray(x0, dtype=np.float64)
    B = x0.shape[0]
    base, s = batch_sensitivity(O, p, x0, U0, n_draws=n_draws, phi=phi, base=base)
    M = base["alpha_idx"].shape[1]
    it_g, it_o = np.asarray(got["iters"]).astype(np.int64), base["iters"].astype(np.int64)
    pre = flow_prefix(np.asarray(got["alpha_idx"])[:, :M], it_g, base["alpha_idx"], it_o)
    same = (it_g == it_o) & (np.asarray(got["status"]) == base["status"]) & (pre >= it_o)
    err = {k: member_rel_err(got[k], base[k], floor) for k, floor in KEYS}
    err["cost"] = member_rel_err(got["cost"], base["cost"])
    with np.errstate(invalid="ignore", divide="ignore"):
        e_it = np.nan_to_num(np.abs(np.asarray(got["cost_trace"])[:, :M] - base["cost_trace"]) / np.abs(base["cost_trace"]),
                             nan=0.0)
    # members whose flow differs: the oracle must flip at or before the same iteration under rounding noise
    flipped = np.flatnonzero(~same)
    unexplained = [int(b) for b in flipped if s["prefix"][b] > pre[b]]
    if unexplained and deep_draws > 0:
        sub = np.array(unexplained)
        U0s = np.broadcast_to(U0, (B, p.m, p.N))[sub]
        _, s2 = batch_sensitivity(O, p, x0[sub], U0s, n_draws=deep_draws, seed=101,
                                  phi=None if phi is None else np.asarray(phi)[sub])
        for j, b in enumerate(sub):
            s["prefix"][b] = min(s["prefix"][b], s2["prefix"][j])
            s["cost_it"][b] = np.maximum(s["cost_it"][b], s2["cost_it"][j])
        unexplained = [int(b) for b in flipped if s["prefix"][b] > pre[b]]
    failures = [dict(member=b, kind="unexplained control-flow flip", iteration=int(pre[b]),
                     oracle_stable_prefix=int(s["prefix"][b])) for b in unexplained]
    # before a flip (and for members without one: everywhere) the per-iteration costs must agree
    col = np.arange(M)[None, :]
    upto = np.where(same, it_o, np.minimum(pre, it_o))
    bound_it = np.maximum(tol, SENS_FACTOR * s["cost_it"])
    bad_it = (e_it > bound_it) & (col < upto[:, None])
    for b in np.flatnonzero(bad_it.any(axis=1)):
        i = int(bad_it[b].argmax())
        failures.append(dict(member=int(b), kind="cost before any flip", iteration=i, err=float(e_it[b, i]),
                             sens=float(s["cost_it"][b, i])))
    worst = {}
    for k in ("X", "U", "K", "U_ff", "cost"):
        bound = np.maximum(tol, SENS_FACTOR * s[k])
        bad = same & (err[k] > bound)
        for b in np.flatnonzero(bad):
            failures.append(dict(member=int(b), kind=k, err=float(err[k][b]), sens=float(s[k][b])))
        es = np.where(same, err[k], 0.0)
        w = int(es.argmax())
        worst[k] = dict(member=w, err=float(es[w]), sens=float(s[k][w]), frac_at_tol=float((es[same] <= tol).mean()) if same.any() else None,
                        max_over_bound=float(np.max(np.where(same, err[k] / bound, 0.0))))
    report = dict(members=int(B), same_flow=int(same.sum()), flipped=int(len(flipped)),
                  flips_explained_by_oracle_noise=int(len(flipped) - len(unexplained)), unexplained=int(len(unexplained)),
                  n_draws=n_draws, deep_draws=deep_draws, tol=tol, sens_factor=SENS_FACTOR, worst=worst,
                  flip_iteration_hist=np.bincount(pre[~same], minlength=M + 1).tolist() if len(flipped) else [])
    return report, failures


def gpu_result(sol, X, U, cost):
    """dict for member_parity() from a batched iLQR solver with enable_trace() on"""
    idx, tc = sol.trace_arrays()
    return dict(X=np.asarray(X), U=np.asarray(U), K=np.asarray(sol.K), U_ff=np.asarray(sol.U_ff), cost=np.asarray(cost),
                iters=np.asarray(sol.iterations), status=np.asarray(sol.status), alpha_idx=idx, cost_trace=tc[:, 1:])


def write_report(name, report):
    """parity distributions go to gpurun_out/ on the GPU box (merged back by gpurun) and are committed under profiles/"""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_r02.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = report
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def mpc_member_parity(O, p_opt, p_plant, x0, ticks, X_sim, iters, n_draws=4, deep_draws=24, tol=1e-9, members=None):
    """Closed-loop MPC runs member by member against the oracle (orc_mpc): X_sim (B, n, ticks+1), iters (B, ticks) from
    the GPU.  A member with the oracle's per-tick iteration counts must be within max(tol, SENS_FACTOR x its own drift
    under 1e-14 noise / FMA contraction); a member whose counts differ must be one whose counts the oracle itself
    changes under that noise.  Returns (report, failures)."""
    x0 = np.asarray(x0, dtype=np.float64)
    members = np.arange(x0.shape[0]) if members is None else np.asarray(members)
    rng = np.random.default_rng(17)
    failures, errs, sens, same_n, flipped, explained = [], [], [], 0, 0, 0

    def drift(b, base, nd, seed_rng):
        s, stable = 0.0, True
        for d in range(nd + 1):
            if d == nd:
                with O.rounding_variant():
                    r = O.mpc(p_opt, p_plant, x0[b], ticks)
            else:
                r = O.mpc(p_opt, p_plant, _perturb(seed_rng, x0[b], 1e-14), ticks)
            if not np.array_equal(r["iters"], base["iters"]):
                stable = False
            s = max(s, float(np.max(np.abs(r["X_sim"] - base["X_sim"])) / np.max(np.abs(base["X_sim"]))))
        return s, stable

    for j, b in enumerate(members):
        base = O.mpc(p_opt, p_plant, x0[b], ticks)
        s, stable = drift(b, base, n_draws, rng)
        e = float(np.max(np.abs(X_sim[j] - base["X_sim"])) / np.max(np.abs(base["X_sim"])))
        if np.array_equal(base["iters"], iters[j]):
            same_n += 1
            errs.append(e)
            sens.append(s)
            if e > max(tol, SENS_FACTOR * s):
                failures.append(dict(member=int(b), kind="X_sim", err=e, sens=s))
        else:
            flipped += 1
            if stable:
                _, stable = drift(b, base, deep_draws, np.random.default_rng(1000 + int(b)))
            if stable:
                failures.append(dict(member=int(b), kind="unexplained iteration-count difference",
                                     got=np.asarray(iters[j]).tolist(), oracle=base["iters"].tolist()))
            else:
                explained += 1
    errs = np.array(errs) if errs else np.zeros(1)
    report = dict(members=int(len(members)), same_flow=same_n, flipped=flipped, flips_explained_by_oracle_noise=explained,
                  worst_err=float(errs.max()), frac_at_tol=float((errs <= tol).mean()), tol=tol, sens_factor=SENS_FACTOR,
                  max_sens=float(max(sens)) if sens else 0.0)
    return report, failures
