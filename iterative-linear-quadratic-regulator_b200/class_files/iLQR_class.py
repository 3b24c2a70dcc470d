"""iLQR -- drop-in for the reference's class_files/iLQR_class.py, batched and running on B200.

Same constructor, attributes and methods as the reference class (iLQR_class.py:18-76, 122-161,
193-247, 250-313): `iLQR(system, T, x_0, U_init, tol, maxiter, alpha_factor, min_alpha, verbose)`,
`.backward_pass(X, U) -> (U_ff, K)`, `.forward_pass(x_0, alpha, X, U, U_ff, K) -> (X, U, cost)`,
`.optimize_trajectory() -> (X, U, cost)`, mutable `.x_0 .U .X .K .U_ff`.

Extension: a leading batch axis.  `x_0` of shape (B, n_x) (and optionally `U_init` of shape
(B, n_u, N)) solves B independent problems at once; every array then carries the batch axis first:
X (B, n_x, N+1), U (B, n_u, N), U_ff (B, n_u, N), K (B, N, n_u, n_x), cost (B,).  With a plain
(n_x,) `x_0` all shapes are exactly the reference's.

numpy/list inputs give numpy outputs (copied to the host); CUDA torch tensors in give CUDA torch
tensors out (device-resident views, no copy).  All numerics run in libilqr_b200.so; torch is only
the buffer allocator.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from . import _device as D
from .systems.system_base import System


class iLQR:
    def __init__(self, system: System, T: float, x_0, U_init, tol: float = 1e-5, maxiter: int = 100,
                 alpha_factor: float = 0.5, min_alpha: float = 1e-8, verbose: bool = True, n_alpha: int = 10,
                 phi=None, reg_init: float = 0.0, reg_factor: float = 0.0, reg_min: float = 1e-6,
                 reg_max: float = 1e10):
        self.system = system
        self.T = T
        self.tol = tol
        self.maxiter = maxiter
        self.alpha_factor = alpha_factor
        self.min_alpha = min_alpha
        self.verbose = verbose
        self.n_alpha = n_alpha            # line-search tries; the reference hard-codes 10 (iLQR_class.py:281)

        self.n_x = system.n_x
        self.n_u = system.n_u
        self.dt = system.dt

        # iLQR_class.py:46-47 -- numpy's arange length rule is jnp.arange's
        self.tspan = D.host(np.arange(0, T + self.dt, self.dt))
        self.N = len(self.tspan) - 1

        x0_shape = tuple(x_0.shape) if hasattr(x_0, "shape") else np.shape(x_0)
        if len(x0_shape) == 1:
            self.batched, self.B = False, 1
        elif len(x0_shape) == 2:
            self.batched, self.B = True, int(x0_shape[0])
        else:
            raise ValueError(f"x_0 must have shape ({self.n_x},) or (B, {self.n_x}), but got {x0_shape}")
        if x0_shape[-1] != self.n_x:
            raise ValueError(f"x_0 must have shape ({self.n_x},) or (B, {self.n_x}), but got {x0_shape}")

        # iLQR_class.py:50-52
        expected_shape = (self.n_u, self.N)
        u_shape = tuple(U_init.shape) if hasattr(U_init, "shape") else np.shape(U_init)
        if u_shape != expected_shape and not (self.batched and u_shape == (self.B,) + expected_shape):
            raise ValueError(f"U_init must have shape {expected_shape}, but got {u_shape}")

        self._torch_out = D.is_torch(x_0) and x_0.is_cuda
        self._tdt = D.torch_dtype(system.dtype)
        # Extension (no counterpart in the reference): Levenberg-Marquardt regularisation Q_uu + mu I with the
        # per-trajectory mu scheduled on the device; reg_factor <= 1 (default) keeps the reference behaviour
        self.reg_factor = reg_factor
        self._handle = D.Handle(system.make_problem(self.N, self.B, tol=tol, maxiter=maxiter,
                                                    alpha_factor=alpha_factor, min_alpha=min_alpha, n_alpha=n_alpha,
                                                    reg_init=reg_init, reg_factor=reg_factor, reg_min=reg_min,
                                                    reg_max=reg_max), lib=system._library())
        n, m, N, B = self.n_x, self.n_u, self.N, self.B
        dev = dict(dtype=self._tdt, device="cuda")
        # device state, batch-innermost (include/ilqr_b200.h)
        self._x0 = torch.empty((n, B), **dev)
        self._X = torch.zeros((N + 1, n, B), **dev)          # iLQR_class.py:55
        self._U = torch.empty((N, m, B), **dev)
        self._K = torch.zeros((N, m, n, B), **dev)           # :59
        self._k = torch.zeros((N, m, B), **dev)              # :61
        self._cost = torch.zeros((B,), **dev)
        self._iters = torch.zeros((B,), dtype=torch.int32, device="cuda")
        self._status = torch.full((B,), 3, dtype=torch.int32, device="cuda")
        self._trace = None
        self._mu = None
        if reg_factor > 1.0:
            self._mu = torch.full((B,), float(reg_init), **dev)
            self._handle.check(self._handle.lib.ilqr_set_mu_buffer(self._handle.h, D.ptr(self._mu)))
        # per-trajectory phase of the synthetic LTV system (config 4); unused by the pendulum models
        self._phi = None
        if phi is not None:
            self._phi = D.to_device(phi, self._tdt).reshape(-1).contiguous()
            if self._phi.shape[0] != self.B:
                raise ValueError(f"phi must hold {self.B} value(s), got {tuple(self._phi.shape)}")
        self._async = None
        self.x_0 = x_0
        self.U = U_init
        self.total_iterations = 0

    # ------------------------------------------------------------------ layout helpers
    def _in_time_major(self, a, dims):
        """reference layout ((B,) dims..., time) -> device [time][dims...][B]"""
        t = D.to_device(a, self._tdt)
        want = len(dims) + 1
        if t.ndim == want:
            t = t.unsqueeze(0).expand(self.B, *t.shape)
        elif t.ndim != want + 1 or t.shape[0] != self.B:
            raise ValueError(f"expected an array of shape {dims + ('time',)} with an optional batch axis of {self.B}, "
                             f"got {tuple(t.shape)}")
        # (B, d1.., T) -> (T, d1.., B)
        perm = [t.ndim - 1] + list(range(1, t.ndim - 1)) + [0]
        return t.permute(*perm).contiguous()

    def _out_time_major(self, t):
        """device [time][dims...][B] -> reference layout ((B,) dims..., time)"""
        perm = [t.ndim - 1] + list(range(1, t.ndim - 1)) + [0]
        v = t.permute(*perm)
        return self._finish(v)

    def _finish(self, v):
        if not self.batched:
            v = v[0]
        if self._torch_out:
            return v
        return D.to_host(v)

    # ------------------------------------------------------------------ attributes
    @property
    def x_0(self):
        return self._finish(self._x0.t())

    @x_0.setter
    def x_0(self, value):
        t = D.to_device(value, self._tdt).reshape(-1, self.n_x)
        if t.shape[0] != self.B:
            raise ValueError(f"x_0 must hold {self.B} initial state(s), got {tuple(t.shape)}")
        self._x0.copy_(t.t())

    @property
    def X(self):
        return self._out_time_major(self._X)

    @X.setter
    def X(self, value):
        self._X.copy_(self._in_time_major(value, (self.n_x,)))

    @property
    def U(self):
        return self._out_time_major(self._U)

    @U.setter
    def U(self, value):
        self._U.copy_(self._in_time_major(value, (self.n_u,)))

    @property
    def U_ff(self):
        return self._out_time_major(self._k)

    @U_ff.setter
    def U_ff(self, value):
        self._k.copy_(self._in_time_major(value, (self.n_u,)))

    @property
    def K(self):
        # device [N][m][n][B] -> (B, N, m, n)
        return self._finish(self._K.permute(3, 0, 1, 2))

    @K.setter
    def K(self, value):
        self._K.copy_(self._K_in(value))

    def _K_in(self, value):
        t = D.to_device(value, self._tdt)
        if t.ndim == 3:
            t = t.unsqueeze(0).expand(self.B, *t.shape)
        return t.permute(1, 2, 3, 0).contiguous()

    @property
    def cost(self):
        return self._finish(self._cost) if self.batched else self._scalar(self._cost)

    @property
    def iterations(self):
        """backward passes executed per trajectory in the last optimize_trajectory()"""
        v = self._iters if self.batched else self._iters[0]
        return v if self._torch_out else v.cpu().numpy()

    @property
    def status(self):
        """per-trajectory exit status code (0 converged, 1 line search failed, 2 maxiter)"""
        v = self._status if self.batched else self._status[0]
        return v if self._torch_out else v.cpu().numpy()

    @property
    def mu(self):
        """per-trajectory regularisation after the last solve (None unless reg_factor > 1)"""
        if self._mu is None:
            return None
        v = self._mu if self.batched else self._mu[0]
        return v if self._torch_out else v.cpu().numpy()

    def _scalar(self, t):
        if self._torch_out:
            return t[0]
        return D.host(t[0].cpu().numpy())

    # ------------------------------------------------------------------ passes
    def backward_pass(self, X_nom, U_nom):
        """(U_ff, K) about the nominal (X_nom, U_nom); iLQR_class.py:122-161."""
        h = self._handle
        X = self._in_time_major(X_nom, (self.n_x,))
        U = self._in_time_major(U_nom, (self.n_u,))
        K = torch.empty_like(self._K)
        k = torch.empty_like(self._k)
        ws = h.workspace()
        h.check(h.lib.ilqr_backward_pass(h.h, D.ptr(self._phi), D.ptr(X), D.ptr(U), D.ptr(K), D.ptr(k), D.ptr(ws), ws.numel(),
                                         D.stream_ptr()))
        return self._out_time_major(k), self._finish(K.permute(3, 0, 1, 2))

    def forward_pass(self, x_0, alpha, X_old, U_old, U_ff, K):
        """(X_new, U_new, cost) for one step size; iLQR_class.py:193-247."""
        h = self._handle
        x0 = D.to_device(x_0, self._tdt).reshape(-1, self.n_x)
        if x0.shape[0] == 1 and self.B > 1:
            x0 = x0.expand(self.B, self.n_x)
        x0 = x0.t().contiguous()
        Xo = self._in_time_major(X_old, (self.n_x,))
        Uo = self._in_time_major(U_old, (self.n_u,))
        k = self._in_time_major(U_ff, (self.n_u,))
        Kd = self._K_in(K)
        Xn, Un = torch.empty_like(self._X), torch.empty_like(self._U)
        cost = torch.empty_like(self._cost)
        h.check(h.lib.ilqr_rollout(h.h, D.ptr(self._phi), D.ptr(x0), float(alpha), D.ptr(Xo), D.ptr(Uo), D.ptr(k), D.ptr(Kd),
                                   D.ptr(Xn), D.ptr(Un), D.ptr(cost), D.stream_ptr()))
        c = self._finish(cost) if self.batched else self._scalar(cost)
        return self._out_time_major(Xn), self._out_time_major(Un), c

    # ------------------------------------------------------------------ solve
    def solve_device(self, sync=True):
        """Run optimize_trajectory() on the device state in place.  Returns sum_b iterations when
        `sync`, else None (fully asynchronous on the current stream)."""
        h = self._handle
        ws = h.workspace()
        tot = C.c_int64(0)
        h.check(h.lib.ilqr_solve(h.h, D.ptr(self._phi), D.ptr(self._x0), D.ptr(self._X), D.ptr(self._U), D.ptr(self._K),
                                 D.ptr(self._k), D.ptr(self._cost), D.ptr(self._iters), D.ptr(self._status), D.ptr(ws),
                                 ws.numel(), D.stream_ptr(), C.byref(tot) if sync else None))
        if sync:
            self.total_iterations = int(tot.value)
            return self.total_iterations
        return None

    def optimize_trajectory(self):
        """Runs the full iLQR loop (iLQR_class.py:250-313) for every trajectory of the batch."""
        if self.verbose and self.B <= 64:
            self.enable_trace()
        self.solve_device(sync=True)
        if self.verbose:
            self._report()
        cost = self._finish(self._cost) if self.batched else self._scalar(self._cost)
        return self.X, self.U, cost

    def enable_trace(self):
        """Record the control flow of every following solve on the device (ilqr_set_trace): per iteration the
        accepted try index and the cost.  Always on for verbose solves of up to 64 trajectories."""
        if self._trace is None:
            ta = torch.full((max(self.maxiter, 1), self.B), -2, dtype=torch.int32, device="cuda")
            tc = torch.full((self.maxiter + 1, self.B), float("nan"), dtype=self._tdt, device="cuda")
            self._trace = (ta, tc)
            h = self._handle
            h.check(h.lib.ilqr_set_trace(h.h, D.ptr(ta), D.ptr(tc)))

    # ------------------------------------------------------------------ pipelined host-in / host-out solves
    def optimize_trajectory_async(self):
        """optimize_trajectory() without waiting for the results: enqueues the solve, then the device-side transposes
        into the reference layout and, on a second stream, the device->host copies into pinned memory.  Returns a
        PendingSolve whose .result() gives (X, U, cost) exactly as optimize_trajectory() would.

        The next solve may be set up (x_0 / U / reset_state) and started while the copies of this one are still in
        flight -- with two result slots alternating, the D2H traffic of solve i runs under the kernels of solve i+1:

            pending = None
            for x0 in batches:
                sol.x_0 = x0; sol.U = U_init; sol.reset_state()
                nxt = sol.optimize_trajectory_async()
                if pending is not None:
                    X, U, cost = pending.result()
                pending = nxt

        The arrays returned by .result() are views of the slot's pinned buffers: they stay valid until the second
        following optimize_trajectory_async() call re-uses the slot (copy them if they must live longer).  (The
        reference returns JAX arrays, which are futures too: its scripts call .block_until_ready() on them.)"""
        if self._torch_out:
            raise RuntimeError("optimize_trajectory_async() is the host-in/host-out path; with CUDA tensors use "
                               "solve_device(sync=False)")
        if self._async is None:
            self._async = D.AsyncResults(self)
        self.solve_device(sync=False)
        return self._async.enqueue()

    def trace_arrays(self):
        """(alpha_idx (B, maxiter), cost_trace (B, maxiter + 1)) of the last solve as numpy arrays: accepted try
        index per iteration (-1 = line search failed; entries at or beyond iterations[b] are stale) and the cost
        after the initial rollout and after every iteration.  Needs enable_trace()."""
        if self._trace is None:
            raise RuntimeError("no trace recorded (call enable_trace() before solving)")
        ta, tc = self._trace
        return ta.t().cpu().numpy(), tc.t().cpu().numpy()

    def trace(self, b=0):
        """(accepted alpha per iteration, cost after each iteration incl. the initial rollout) of
        trajectory b in the last solve; needs verbose=True and B <= 64."""
        if self._trace is None:
            raise RuntimeError("no trace recorded (construct with verbose=True and B <= 64)")
        it = int(self._iters[b].item())
        ta, tc = self._trace
        idx = ta[:it, b].cpu().numpy()
        alphas = np.where(idx >= 0, self.alpha_factor ** np.maximum(idx, 0), np.nan)
        return idx, alphas, tc[: it + 1, b].cpu().numpy()

    def _report(self):
        """The reference's verbose messages (iLQR_class.py:262,269,296,306,311), printed after the
        solve because the loop itself never returns to the host."""
        if self.B <= 64 and self._trace is not None and not self.batched:
            idx, alphas, costs = self.trace(0)
            st = int(self._status[0].item())
            print(f"Initial cost: {costs[0]:.4f}")
            for i, (w, a) in enumerate(zip(idx, alphas)):
                if w >= 0:
                    print(f"  Iter {i+1} (alpha={a:.2e}): Cost improved to {costs[i+1]:.4f}")
                else:
                    print(f"Warning: Line search failed at iteration {i+1}. Cost did not improve.")
            if st == 0:
                print(f"Converged at iteration {len(idx)}")
            if st == 2 or (st == 0 and len(idx) == self.maxiter - 1):
                print(f"Warning: Reached max iterations ({self.maxiter}) without converging.")
            return
        st = self._status.cpu().numpy()
        it = self._iters.cpu().numpy()
        c = self._cost.cpu().numpy()
        print(f"iLQR batch of {self.B}: converged {int((st == 0).sum())}, line-search failed {int((st == 1).sum())}, "
              f"max iterations {int((st == 2).sum())}; iterations mean {it.mean():.1f} max {int(it.max())}; "
              f"cost mean {np.nanmean(c):.4f}")

    # ------------------------------------------------------------------ MPC helpers
    def mpc_shift(self):
        """U_guess = [U[:,1:], U[:,-1:]] in place on the device and return u_0 = U[:,0]
        (run_iLQR_UA_MPC.py:157,168) as a device tensor [m][B]."""
        h = self._handle
        u0 = torch.empty((self.n_u, self.B), dtype=self._tdt, device="cuda")
        h.check(h.lib.ilqr_mpc_shift(h.h, D.ptr(self._U), D.ptr(u0), D.stream_ptr()))
        return u0

    def launches(self):
        return self._handle.launches()

    def set_linesearch_waves(self, sizes):
        """Line-search schedule of optimize_trajectory (see ilqr_set_linesearch_waves): `sizes` = tries per
        lazily evaluated wave, e.g. (2, 2, 2, 4); () or None = all tries eagerly.  The accepted step size of
        every trajectory is the same under every schedule; only the amount of speculative work changes."""
        h = self._handle
        sizes = [int(v) for v in (sizes or ())]
        arr = (C.c_int32 * max(len(sizes), 1))(*sizes)
        h.check(h.lib.ilqr_set_linesearch_waves(h.h, len(sizes), arr))

    def linesearch_waves(self):
        """tries per lazily evaluated wave, () for the eager schedule (ilqr_get_linesearch_waves)"""
        h = self._handle
        arr = (C.c_int32 * 8)()
        n = h.lib.ilqr_get_linesearch_waves(h.h, arr)
        return tuple(int(arr[i]) for i in range(n))

    def first_wave(self):
        """eager schedule: step sizes rolled out for every trajectory in the first wave"""
        h = self._handle
        arr = (C.c_int32 * 8)()
        n = h.lib.ilqr_get_linesearch_waves(h.h, arr)
        return int(arr[0]) if n == 0 else 0

    def set_profiling(self, enable=True):
        """chain CUDA events between the solve's kernels (see ilqr_set_profiling)"""
        h = self._handle
        h.check(h.lib.ilqr_set_profiling(h.h, 1 if enable else 0))

    def kernel_times(self):
        """{kernel class: (total ms, launches)} accumulated since set_profiling(True)"""
        h = self._handle
        n = len(_cabi.KERNEL_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        h.check(h.lib.ilqr_get_kernel_times(h.h, ms, cnt))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_cabi.KERNEL_CLASSES)}

    def reset_state(self):
        """fresh-solver state: X = K = U_ff = 0 (iLQR_class.py:55-61); U and x_0 are left as they are"""
        self._X.zero_()
        self._K.zero_()
        self._k.zero_()
