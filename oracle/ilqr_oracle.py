"""ctypes front end of oracle/libilqr_oracle.so -- TEST INFRASTRUCTURE ONLY.

The CPU float64 restatement of the reference hot path (see ilqr_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Array conventions are the reference's: X (n,N+1), U (m,N), U_ff (m,N), K (N,m,n)
(/root/reference/python/class_files/iLQR_class.py:54-61).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libilqr_oracle.so")

NMAX, MMAX = 12, 4
MODELS = {"pendulum": 0, "double": 1, "ua": 2, "ltv": 3}
INTEGRATORS = {"euler": 0, "midpoint": 1, "rk4": 2, "backward_euler": 3}
STATUS = {0: "converged", 1: "ls_failed", 2: "maxiter"}
DIMS = {"pendulum": (2, 1), "double": (4, 2), "ua": (4, 1)}


class Problem(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("integrator", C.c_int), ("n", C.c_int), ("m", C.c_int), ("N", C.c_int),
        ("n_alpha", C.c_int), ("maxiter", C.c_int),
        ("dt", C.c_double), ("tol", C.c_double), ("alpha_factor", C.c_double), ("min_alpha", C.c_double),
        ("phys", C.c_double * 16),
        ("Q", C.c_double * (NMAX * NMAX)), ("R", C.c_double * (MMAX * MMAX)),
        ("Qf", C.c_double * (NMAX * NMAX)), ("x_target", C.c_double * NMAX),
        ("Ac", C.c_double * (NMAX * NMAX)), ("E", C.c_double * (NMAX * NMAX)),
        ("Bc", C.c_double * (NMAX * MMAX)), ("ltv_amp", C.c_double),
        ("reg_init", C.c_double), ("reg_factor", C.c_double), ("reg_min", C.c_double), ("reg_max", C.c_double),
    ]


def build(force=False):
    """Compile the oracle with gcc (also done by __graft_entry__.build())."""
    src = os.path.join(_HERE, "ilqr_oracle.c")
    fma = _LIB_PATH.replace(".so", "_fma.so")
    if force or not all(os.path.exists(f) and os.path.getmtime(f) >= os.path.getmtime(src) for f in (_LIB_PATH, fma)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "all"])
    return _LIB_PATH


_libs = {}
_variant = [""]


class rounding_variant:
    """Context manager: run the oracle from libilqr_oracle_fma.so, the same source compiled with
    -ffp-contract=fast -mfma, i.e. with a different rounding in nearly every operation.  The distance
    between the two builds measures how strongly a computation amplifies rounding-level noise
    (tests/helpers.py uses it to calibrate parity bounds)."""

    def __init__(self, name="fma"):
        self.name = name

    def __enter__(self):
        _variant.append(self.name)

    def __exit__(self, *a):
        _variant.pop()


def lib():
    v = _variant[-1]
    if v not in _libs:
        build()
        L = C.CDLL(_LIB_PATH if not v else _LIB_PATH.replace(".so", f"_{v}.so"))
        P, D, I = C.POINTER(Problem), C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_f_cont.argtypes = [P, C.c_int, C.c_double, D, D, D]
        L.orc_f_cont_jac.argtypes = [P, C.c_int, C.c_double, D, D, D, D]
        L.orc_f.argtypes = [P, C.c_int, C.c_double, D, D, D]
        L.orc_f.restype = C.c_int
        L.orc_f_jac.argtypes = [P, C.c_int, C.c_double, D, D, D, D]
        L.orc_l.argtypes = [P, D, D]
        L.orc_l.restype = C.c_double
        L.orc_lf.argtypes = [P, D]
        L.orc_lf.restype = C.c_double
        L.orc_l_derivs.argtypes = [P, D, D, D, D, D, D, D]
        L.orc_lf_derivs.argtypes = [P, D, D, D]
        L.orc_backward_pass.argtypes = [P, C.c_double, D, D, D, D]
        L.orc_forward_pass.argtypes = [P, C.c_double, D, C.c_double, D, D, D, D, D, D]
        L.orc_forward_pass.restype = C.c_double
        L.orc_optimize.argtypes = [P, C.c_double, D, D, D, D, D, I, I, D, I, D]
        L.orc_optimize.restype = C.c_double
        L.orc_optimize_ex.argtypes = [P, C.c_double, D, D, D, D, D, I, I, D, I, D, D]
        L.orc_optimize_ex.restype = C.c_double
        L.orc_backward_pass_mu.argtypes = [P, C.c_double, C.c_double, D, D, D, D]
        L.orc_optimize_batch.argtypes = [P, C.c_int, D, D, D, D, D, D, D, D, I, I, C.c_int]
        L.orc_optimize_batch_trace.argtypes = [P, C.c_int, D, D, D, D, D, D, D, D, I, I, C.c_int, I, D, D]
        L.orc_mpc.argtypes = [P, P, C.c_double, D, C.c_int, D, D, D, I, D, D, D, D, D, D]
        L.orc_max_threads.restype = C.c_int
        _libs[v] = L
    return _libs[v]


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _arr(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64))


def _full(M, n):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 1:
        M = np.diag(M)
    assert M.shape == (n, n), M.shape
    return M


def horizon(T, dt):
    """N = len(arange(0, T+dt, dt)) - 1  (iLQR_class.py:46-47)."""
    return len(np.arange(0, T + dt, dt)) - 1


def make_problem(kind, integrator, N, dt, Q, R, Q_f, x_target, phys=None, tol=1e-5, maxiter=100,
                 alpha_factor=0.5, min_alpha=1e-8, n_alpha=10, ltv=None, reg_init=0.0, reg_factor=0.0, reg_min=1e-6,
                 reg_max=1e10):
    """kind in MODELS; phys = dict of the system's physical parameters; Q/R/Q_f diag vectors or full."""
    p = Problem()
    if kind == "ltv":
        n, m = ltv["Ac"].shape[0], ltv["Bc"].shape[1]
    else:
        n, m = DIMS[kind]
    p.model, p.integrator, p.n, p.m, p.N = MODELS[kind], INTEGRATORS[integrator], n, m, int(N)
    p.n_alpha, p.maxiter = int(n_alpha), int(maxiter)
    p.dt, p.tol, p.alpha_factor, p.min_alpha = float(dt), float(tol), float(alpha_factor), float(min_alpha)
    p.reg_init, p.reg_factor, p.reg_min, p.reg_max = float(reg_init), float(reg_factor), float(reg_min), float(reg_max)
    phys = dict(phys or {})
    if kind == "pendulum":
        vals = [phys.get("g", 9.81), phys.get("l", 1.0), phys.get("d", 0.01)]
    elif kind in ("double", "ua"):
        vals = [phys.get("g", 9.81), phys.get("m1", 1.0), phys.get("m2", 1.0), phys.get("l1", 1.0),
                phys.get("l2", 1.0), phys.get("d1", 0.01), phys.get("d2", 0.01),
                phys.get("theta1", 0.0), phys.get("theta2", 0.0)]
    else:
        vals = []
    for i, v in enumerate(vals):
        p.phys[i] = float(v)
    for name, M, k in (("Q", Q, n), ("R", R, m), ("Qf", Q_f, n)):
        flat = _full(M, k).ravel()
        getattr(p, name)[: flat.size] = flat.tolist()
    xt = _arr(x_target).ravel()
    p.x_target[:n] = xt.tolist()
    if kind == "ltv":
        p.Ac[: n * n] = _arr(ltv["Ac"]).ravel().tolist()
        p.E[: n * n] = _arr(ltv["E"]).ravel().tolist()
        p.Bc[: n * m] = _arr(ltv["Bc"]).ravel().tolist()
        p.ltv_amp = float(ltv.get("amp", 0.1))
    return p


def problem_from_golden(g, integrator=None, **over):
    """Build a Problem from the p_* entries of a tests/golden/*.npz file."""
    kind = str(g["p_kind"])
    integ = integrator or str(g["p_integrator"])
    names = ("g", "l", "d") if kind == "pendulum" else ("g", "m1", "m2", "l1", "l2", "d1", "d2", "theta1", "theta2")
    phys = {k: float(g["p_" + k]) for k in names}
    N = int(g["N"]) if "N" in g else 1
    kw = dict(tol=float(g["tol"]) if "tol" in g else 1e-5, maxiter=int(g["maxiter"]) if "maxiter" in g else 100)
    kw.update(over)
    return make_problem(kind, integ, N, float(g["p_dt"]), g["p_Q"], g["p_R"], g["p_Q_f"], g["p_x_target"], phys, **kw)


# ------------------------------------------------------------------ point functions
def f(p, x, u, t=0, phi=0.0):
    x, u = _arr(x), _arr(u)
    out = np.empty(p.n)
    lib().orc_f(C.byref(p), t, phi, _d(x), _d(u), _d(out))
    return out


def f_jac(p, x, u, t=0, phi=0.0):
    x, u = _arr(x), _arr(u)
    A, B = np.empty((p.n, p.n)), np.empty((p.n, p.m))
    lib().orc_f_jac(C.byref(p), t, phi, _d(x), _d(u), _d(A), _d(B))
    return A, B


def f_cont_jac(p, x, u, t=0, phi=0.0):
    x, u = _arr(x), _arr(u)
    A, B = np.empty((p.n, p.n)), np.empty((p.n, p.m))
    lib().orc_f_cont_jac(C.byref(p), t, phi, _d(x), _d(u), _d(A), _d(B))
    return A, B


def l(p, x, u):
    x, u = _arr(x), _arr(u)
    return lib().orc_l(C.byref(p), _d(x), _d(u))


def lf(p, x):
    x = _arr(x)
    return lib().orc_lf(C.byref(p), _d(x))


def l_derivs(p, x, u):
    x, u = _arr(x), _arr(u)
    n, m = p.n, p.m
    lx, lu, lxx, luu, lux = np.empty(n), np.empty(m), np.empty((n, n)), np.empty((m, m)), np.empty((m, n))
    lib().orc_l_derivs(C.byref(p), _d(x), _d(u), _d(lx), _d(lu), _d(lxx), _d(luu), _d(lux))
    return lx, lu, lxx, luu, lux


def lf_derivs(p, x):
    x = _arr(x)
    lfx, lfxx = np.empty(p.n), np.empty((p.n, p.n))
    lib().orc_lf_derivs(C.byref(p), _d(x), _d(lfx), _d(lfxx))
    return lfx, lfxx


# ------------------------------------------------------------------ passes / solves
def backward_pass(p, X, U, phi=0.0):
    X, U = _arr(X), _arr(U)
    U_ff, K = np.empty((p.m, p.N)), np.empty((p.N, p.m, p.n))
    lib().orc_backward_pass(C.byref(p), phi, _d(X), _d(U), _d(U_ff), _d(K))
    return U_ff, K


def forward_pass(p, x0, alpha, X_old, U_old, U_ff, K, phi=0.0):
    x0, X_old, U_old, U_ff, K = map(_arr, (x0, X_old, U_old, U_ff, K))
    Xn, Un = np.empty((p.n, p.N + 1)), np.empty((p.m, p.N))
    c = lib().orc_forward_pass(C.byref(p), phi, _d(x0), float(alpha), _d(X_old), _d(U_old), _d(U_ff), _d(K),
                               _d(Xn), _d(Un))
    return Xn, Un, c


def optimize(p, x0, U_init, state=None, phi=0.0):
    """optimize_trajectory() on a fresh solver (state=None) or on persistent (X,K,U_ff) state."""
    n, m, N = p.n, p.m, p.N
    x0 = _arr(x0)
    U = _arr(U_init).copy()
    if state is None:
        X, K, U_ff = np.zeros((n, N + 1)), np.zeros((N, m, n)), np.zeros((m, N))
    else:
        X, K, U_ff = (_arr(a).copy() for a in state)
    iters, status = C.c_int(0), C.c_int(0)
    cost0 = C.c_double(0.0)
    tr_a = np.full(max(p.maxiter, 1), -2, dtype=np.int32)
    tr_c = np.full(max(p.maxiter, 1), np.nan)
    mu = C.c_double(0.0)
    cost = lib().orc_optimize_ex(C.byref(p), phi, _d(x0), _d(X), _d(U), _d(K), _d(U_ff), C.byref(iters),
                                 C.byref(status), C.byref(cost0), _i(tr_a), _d(tr_c), C.byref(mu))
    return dict(X=X, U=U, K=K, U_ff=U_ff, cost=cost, cost0=cost0.value, iters=iters.value, mu=mu.value,
                status=STATUS[status.value], alpha_idx=tr_a[: iters.value], cost_trace=tr_c[: iters.value])


def optimize_batch(p, x0, U_init, phi=None, nthreads=0, trace=False):
    """x0 (B,n); U_init (B,m,N).  Returns dict of batch-major arrays.  trace=True adds the per-member control
    flow: alpha_idx (B,maxiter) accepted try index per iteration (-1 failed, -2 not run), cost_trace (B,maxiter)
    cost after each iteration (NaN where not run), cost0 (B,) cost of the alpha = 0 rollout."""
    n, m, N = p.n, p.m, p.N
    x0 = _arr(x0)
    B = x0.shape[0]
    U_init = _arr(np.broadcast_to(U_init, (B, m, N)))
    X, U = np.empty((B, n, N + 1)), np.empty((B, m, N))
    K, U_ff = np.empty((B, N, m, n)), np.empty((B, m, N))
    cost = np.empty(B)
    iters, status = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    phi_p = _d(_arr(phi)) if phi is not None else None
    if trace:
        mi = max(p.maxiter, 1)
        tr_a = np.full((B, mi), -2, dtype=np.int32)
        tr_c = np.full((B, mi), np.nan)
        c0 = np.empty(B)
        lib().orc_optimize_batch_trace(C.byref(p), B, phi_p, _d(x0), _d(U_init), _d(X), _d(U), _d(K), _d(U_ff), _d(cost),
                                       _i(iters), _i(status), int(nthreads), _i(tr_a), _d(tr_c), _d(c0))
        return dict(X=X, U=U, K=K, U_ff=U_ff, cost=cost, iters=iters, status=status, alpha_idx=tr_a, cost_trace=tr_c,
                    cost0=c0)
    lib().orc_optimize_batch(C.byref(p), B, phi_p, _d(x0), _d(U_init), _d(X), _d(U), _d(K), _d(U_ff), _d(cost),
                             _i(iters), _i(status), int(nthreads))
    return dict(X=X, U=U, K=K, U_ff=U_ff, cost=cost, iters=iters, status=status)


def mpc(p_opt, p_plant, x0, ticks, U_init=None, phi=0.0):
    n, m, N = p_opt.n, p_opt.m, p_opt.N
    x0 = _arr(x0)
    X, K, U_ff = np.zeros((n, N + 1)), np.zeros((N, m, n)), np.zeros((m, N))
    U = np.zeros((m, N)) if U_init is None else _arr(U_init).copy()
    X_sim, U_sim = np.zeros((n, ticks + 1)), np.zeros((m, ticks))
    costs, iters = np.zeros(ticks), np.zeros(ticks, dtype=np.int32)
    Xb, Ub = np.zeros((ticks, n, N + 1)), np.zeros((ticks, m, N))
    lib().orc_mpc(C.byref(p_opt), C.byref(p_plant), phi, _d(x0), ticks, _d(X_sim), _d(U_sim), _d(costs), _i(iters),
                  _d(X), _d(U), _d(K), _d(U_ff), _d(Xb), _d(Ub))
    return dict(X_sim=X_sim, U_sim=U_sim, costs=costs, iters=iters, X=X, U=U, K=K, U_ff=U_ff, X_bar=Xb, U_bar=Ub)


def max_threads():
    return lib().orc_max_threads()
