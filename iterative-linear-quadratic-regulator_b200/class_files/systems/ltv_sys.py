"""MyLTVSystem -- synthetic linear time-varying system of BASELINE.json config 4 (not a reference class:
the closest upstream artefact is matlab/CLASSES/Linear_iLQR_CLASS.m, an LTI LQR special case).

    x_{t+1} = x_t + dt * ((Ac + amp * sin(2 pi t / N + phi_b) * E) x_t + Bc u_t),   n_x = 12, n_u = 4,

with a per-trajectory phase phi_b.  A_t and B_t are generated inside the kernels from (Ac, E, Bc, phi_b);
they are never stored per trajectory.  Forward Euler only.  For a linear system one iLQR
iteration with alpha = 1 is exact (Linear_iLQR_CLASS.m:135-139), which is the known-answer test.
"""
import numpy as np

from .system_base import System


class MyLTVSystem(System):
    TIME_VARYING = True       # f depends on the time index and a per-trajectory phase (System._f / ._jac take t, phi)

    def __init__(self, dt, x_target, Q, R, Q_f, Ac, E, Bc, amp: float = 0.1, use_jit: bool = True,
                 integrator: str = "euler", dtype: str = "float64"):
        self.Ac = np.asarray(Ac, dtype=np.float64)
        self.E = np.asarray(E, dtype=np.float64)
        self.Bc = np.asarray(Bc, dtype=np.float64)
        self.amp = float(amp)
        self.n_x, self.n_u = self.Ac.shape[0], self.Bc.shape[1]
        if (self.n_x, self.n_u) != (12, 4) or self.E.shape != (12, 12) or self.Bc.shape != (12, 4):
            raise ValueError("MyLTVSystem is built for n_x = 12, n_u = 4 (Ac, E: 12x12, Bc: 12x4)")
        if integrator != "euler":
            raise ValueError(f"Unknown integrator: '{integrator}'. MyLTVSystem supports 'euler' only.")
        self.x_target = x_target
        self.Q = Q
        self.R = R
        self.Q_f = Q_f
        super().__init__(self.n_x, self.n_u, dt, use_jit=use_jit, integrator=integrator, dtype=dtype)

    def _device_model(self):
        return "ltv", []

    def make_problem(self, N, B, **kw):
        p = super().make_problem(N, B, **kw)
        from .. import _cabi
        _cabi.fill(p.Ac, self.Ac.ravel())
        _cabi.fill(p.E, self.E.ravel())
        _cabi.fill(p.Bc, self.Bc.ravel())
        p.ltv_amp = self.amp
        return p

    @staticmethod
    def synthetic(dt=0.01, seed=2, dtype="float64", shift=1.0):
        """The seeded config-4 instance: Ac = G / rho(G) - shift * I and E = H / rho(H) with G, H standard normal,
        Q = I, R = 0.1 I, Q_f = 10 I.  shift = 1 makes Ac Hurwitz.  With an open-loop unstable Ac (shift = 0) the
        reference's backward recursion -- V_xx = Q_xx + Q_ux' K, neither symmetrised nor regularised
        (iLQR_class.py:113-114) -- is itself numerically unstable over N = 1000 steps: the CPU oracle and the
        GPU both end in NaN gains for part of the batch, so the full-size configuration uses the stable system."""
        rng = np.random.default_rng(seed)
        Ac = rng.standard_normal((12, 12))
        Ac *= 1.0 / max(abs(np.linalg.eigvals(Ac)))
        Ac = Ac - shift * np.eye(12)
        E = rng.standard_normal((12, 12))
        E *= 1.0 / max(abs(np.linalg.eigvals(E)))
        Bc = rng.standard_normal((12, 4))
        return MyLTVSystem(dt, np.zeros(12), np.eye(12), 0.1 * np.eye(4), 10.0 * np.eye(12), Ac, E, Bc, amp=0.1,
                           dtype=dtype)
