"""System base class -- drop-in for the reference's class_files/systems/system_base.py.

The reference builds a discrete step from `_f_cont_fcn` with one of four integrators and gets
every derivative from JAX autodiff (system_base.py:25-251).  Here the shipped systems are
device models inside libilqr_b200.so: the same twelve public callables
(`f_fcn, f_x_fcn, f_u_fcn, l_fcn, l_x_fcn, l_u_fcn, l_xx_fcn, l_uu_fcn, l_ux_fcn, l_f_fcn,
l_f_x_fcn, l_f_xx_fcn`, system_base.py:223-251) evaluate on the GPU with analytic Jacobians,
for one point (x (n,), u (m,)) exactly as in the reference, or for a batch of points
(x (P,n), u (P,m)).

The shipped subclasses describe themselves to the kernels through `_device_model()`.  A USER-DEFINED
subclass keeps the reference's contract -- implement `_f_cont_fcn`, `_l_fcn`, `_l_f_fcn`
(system_base.py:255-275) -- written against `class_files.symbolic` instead of `jax.numpy`: the three
methods are traced once symbolically, differentiated analytically and compiled into a device model of
their own (class_files/codegen.py), the counterpart of the reference's jit + autodiff factory
(system_base.py:203-251).
"""
import ctypes as C
from abc import ABC

import numpy as np
import torch

from .. import _cabi
from .. import _device as D

_SUPPORTED = ("rk4", "midpoint", "euler", "backward_euler")


def _square(M, k, name):
    M = np.asarray(M.detach().cpu().numpy() if isinstance(M, torch.Tensor) else M, dtype=np.float64)
    if M.ndim == 0:
        M = M.reshape(1, 1)
    if M.shape != (k, k):
        raise ValueError(f"{name} must have shape {(k, k)}, but got {M.shape}")
    return M


class System(ABC):
    def __init__(self, n_x: int, n_u: int, dt: float, use_jit: bool = True, integrator: str = "rk4",
                 dtype: str = "float64"):
        if not (1 <= int(n_x) <= _cabi.NMAX and 1 <= int(n_u) <= _cabi.MMAX):
            # the kernels hold a trajectory's value function in registers / shared memory: fixed upper bounds
            raise ValueError(f"n_x must be in 1..{_cabi.NMAX} and n_u in 1..{_cabi.MMAX} (ILQR_NMAX / ILQR_MMAX of "
                             f"include/ilqr_b200.h), but got n_x={n_x}, n_u={n_u}")
        self.n_x = n_x
        self.n_u = n_u
        self.dt = dt
        self.use_jit = use_jit          # kept for signature parity; kernels are always compiled
        if integrator not in _SUPPORTED:
            # same message as the reference, system_base.py:197-198
            raise ValueError(f"Unknown integrator: '{integrator}'. Supported: 'rk4', 'midpoint', 'euler', "
                             f"'backward_euler'.")
        if dtype not in _cabi.DTYPES:
            raise ValueError(f"dtype must be 'float64' or 'float32', got {dtype!r}")
        self.integrator = integrator
        self.dtype = dtype
        self._point_handles = {}
        # public callables, same names as system_base.py:223-251
        self.f_fcn = self._f
        self.f_x_fcn = lambda x, u, **kw: self._jac(x, u, **kw)[0]      # kw: phi, N (MyLTVSystem only)
        self.f_u_fcn = lambda x, u, **kw: self._jac(x, u, **kw)[1]
        self.l_fcn = lambda x, u: self._cost(x, u, "l")
        self.l_x_fcn = lambda x, u: self._cost(x, u, "lx")
        self.l_u_fcn = lambda x, u: self._cost(x, u, "lu")
        self.l_xx_fcn = lambda x, u: self._cost(x, u, "lxx")
        self.l_uu_fcn = lambda x, u: self._cost(x, u, "luu")
        self.l_ux_fcn = lambda x, u: self._cost(x, u, "lux")
        self.l_f_fcn = lambda x: self._cost(x, None, "lf")
        self.l_f_x_fcn = lambda x: self._cost(x, None, "lfx")
        self.l_f_xx_fcn = lambda x: self._cost(x, None, "lfxx")

    # ---- what a subclass provides -----------------------------------------------------
    def _device_model(self):
        """-> (model name in _cabi.MODELS, list of physical parameters in ilqr_problem_t.phys order)"""
        if self._is_user_defined():
            return "user", []
        raise NotImplementedError(
            f"{type(self).__name__} has no device model: implement _f_cont_fcn, _l_fcn and _l_f_fcn "
            "(written with class_files.symbolic) or use one of the shipped systems.")

    def _is_user_defined(self):
        cls = type(self)
        return all(getattr(cls, name) is not getattr(System, name) for name in ("_f_cont_fcn", "_l_fcn", "_l_f_fcn"))

    def _library(self):
        """the shared library holding this system's kernels: libilqr_b200.so for the shipped systems, a
        generated one (class_files/codegen.py) for a user-defined subclass"""
        if type(self)._device_model is System._device_model and self._is_user_defined():
            if getattr(self, "_user_lib", None) is None:
                from .. import codegen
                self._user_lib = codegen.build_library(self)
            return self._user_lib
        return _cabi.load()

    def _cost_weights(self):
        if type(self)._device_model is System._device_model and self._is_user_defined():
            n, m = self.n_x, self.n_u           # the generated cost carries its own weights
            return np.zeros((n, n)), np.zeros((m, m)), np.zeros((n, n)), np.zeros(n)
        return self.Q, self.R, self.Q_f, self.x_target

    # ---- problem struct ---------------------------------------------------------------
    def make_problem(self, N, B, tol=1e-5, maxiter=100, alpha_factor=0.5, min_alpha=1e-8, n_alpha=10,
                     reg_init=0.0, reg_factor=0.0, reg_min=1e-6, reg_max=1e10):
        model, phys = self._device_model()
        n, m = self.n_x, self.n_u
        Q, R, Q_f, x_t = self._cost_weights()
        p = _cabi.Problem()
        p.model, p.integrator, p.dtype = _cabi.MODELS[model], _cabi.INTEGRATORS[self.integrator], _cabi.DTYPES[self.dtype]
        p.n, p.m, p.N, p.B = n, m, int(N), int(B)
        p.n_alpha, p.maxiter = int(n_alpha), int(maxiter)
        p.dt, p.tol, p.alpha_factor, p.min_alpha = float(self.dt), float(tol), float(alpha_factor), float(min_alpha)
        _cabi.fill(p.phys, phys)
        _cabi.fill(p.Q, _square(Q, n, "Q").ravel())
        _cabi.fill(p.R, _square(R, m, "R").ravel())     # integer R (run_iLQR_UA_MPC.py:53) is promoted here
        _cabi.fill(p.Qf, _square(Q_f, n, "Q_f").ravel())
        xt = np.asarray(x_t.detach().cpu().numpy() if isinstance(x_t, torch.Tensor) else x_t, dtype=np.float64).ravel()
        if xt.shape != (n,):
            raise ValueError(f"x_target must have shape {(n,)}, but got {xt.shape}")
        _cabi.fill(p.x_target, xt)
        p.reg_init, p.reg_factor, p.reg_min, p.reg_max = float(reg_init), float(reg_factor), float(reg_min), float(reg_max)
        return p

    # ---- point evaluations ------------------------------------------------------------
    def _points(self, x, u):
        """-> (xd [n][P], ud [m][P] or None, P, single, torch_out)"""
        D.require_cuda()
        tdt = D.torch_dtype(self.dtype)
        torch_out = D.is_torch(x) and x.is_cuda
        xd = D.to_device(x, tdt)
        single = xd.ndim == 1
        xd = xd.reshape(-1, self.n_x)
        P = xd.shape[0]
        ud = None
        if u is not None:
            ud = D.to_device(u, tdt).reshape(-1, self.n_u)
            if ud.shape[0] != P:
                raise ValueError(f"x and u disagree on the number of points: {P} vs {ud.shape[0]}")
            ud = ud.t().contiguous()
        return xd.t().contiguous(), ud, P, single, torch_out

    def _handle(self, P, N=1):
        h = self._point_handles.get((P, N))
        if h is None:
            h = D.Handle(self.make_problem(N=N, B=P), lib=self._library())
            self._point_handles[(P, N)] = h
        return h

    def _out(self, t, single, torch_out):
        """t is batch-LAST; move the batch axis first and drop it for a single point."""
        t = t.movedim(-1, 0)
        if single:
            t = t[0]
        if torch_out:
            return t
        return D.host(t.cpu().numpy())

    def _phi(self, phi, P, dtype):
        if phi is None:
            return None
        t = D.to_device(phi, dtype).reshape(-1)
        return (t.expand(P) if t.shape[0] == 1 else t).contiguous()

    def _f(self, x, u, t=0, phi=None, N=1):
        """f(x,u); `t`, `phi` and the horizon `N` (period of the modulation) only matter for MyLTVSystem"""
        xd, ud, P, single, tout = self._points(x, u)
        h = self._handle(P, N)
        xn = torch.empty_like(xd)
        ph = self._phi(phi, P, xd.dtype)
        h.check(h.lib.ilqr_step(h.h, int(t), D.ptr(ph), D.ptr(xd), D.ptr(ud), D.ptr(xn), D.stream_ptr()))
        return self._out(xn, single, tout)

    def _jac(self, x, u, phi=None, N=1):
        """(f_x, f_u) at (x,u), time index 0; `phi` and the horizon `N` only matter for MyLTVSystem (as in _f)"""
        xd, ud, P, single, tout = self._points(x, u)
        h = self._handle(P, N)
        n, m = self.n_x, self.n_u
        X = torch.stack([xd, xd])            # [N+1=2][n][P]; only t=0 is linearized
        U = ud.reshape(1, m, P)
        A = torch.empty((1, n, n, P), dtype=xd.dtype, device="cuda")
        Bd = torch.empty((1, n, m, P), dtype=xd.dtype, device="cuda")
        ph = self._phi(phi, P, xd.dtype)
        h.check(h.lib.ilqr_linearize(h.h, D.ptr(ph), D.ptr(X), D.ptr(U), D.ptr(A), D.ptr(Bd), D.stream_ptr()))
        return self._out(A[0], single, tout), self._out(Bd[0], single, tout)

    def _cost(self, x, u, which):
        n, m = self.n_x, self.n_u
        if u is None:
            u = torch.zeros((x.shape[0], m)) if getattr(x, "ndim", 1) == 2 else torch.zeros(m)
        xd, ud, P, single, tout = self._points(x, u)
        h = self._handle(P)
        X = torch.stack([xd, xd])
        U = ud.reshape(1, m, P)
        shapes = dict(l=(1, P), lx=(1, n, P), lu=(1, m, P), lxx=(1, n, n, P), luu=(1, m, m, P), lux=(1, m, n, P),
                      lf=(P,), lfx=(n, P), lfxx=(n, n, P))
        out = torch.empty(shapes[which], dtype=xd.dtype, device="cuda")
        order = ["l", "lx", "lu", "lxx", "luu", "lux", "lf", "lfx", "lfxx"]
        args = [D.ptr(out) if k == which else C.c_void_p(0) for k in order]
        h.check(h.lib.ilqr_cost_expansion(h.h, D.ptr(X), D.ptr(U), *args, D.stream_ptr()))
        if which in ("lf", "lfx", "lfxx"):
            return self._out(out, single, tout)
        return self._out(out[0], single, tout)

    # ---- the reference's abstract methods ----------------------------------------------
    # (system_base.py:255-275).  The shipped subclasses implement them on the device; they are not
    # abstract here so that the device-backed subclasses need not carry Python bodies.
    def _f_cont_fcn(self, x, u):
        raise NotImplementedError("continuous dynamics live in the CUDA device model of this system")

    def _l_fcn(self, x, u):
        return self.l_fcn(x, u)

    def _l_f_fcn(self, x):
        return self.l_f_fcn(x)
