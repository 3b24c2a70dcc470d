"""MyPendulum -- drop-in for the reference's class_files/systems/pendulum_sys.py:12-98.

x = [theta, theta_dot], u = [tau];  x_dot = [x2, u - d*x2 - (g/l) sin(x1)]  (pendulum_sys.py:60-75);
l = (1/2 dx'Q dx + 1/2 u'R u) dt, l_f = 1/2 dx'Q_f dx  (pendulum_sys.py:77-98).
The dynamics, their analytic Jacobians and the cost run inside libilqr_b200.so (PendulumSys,
csrc/ilqr_systems.cuh).  The reference module's `__main__` integrator demo (:101-313) is
plotting/printing and is not reproduced.
"""
from .system_base import System     # (the reference spells this import absolutely, pendulum_sys.py:10; same module either way)


class MyPendulum(System):
    def __init__(self, dt, x_target, Q, R, Q_f, g: float = 9.81, l: float = 1.0, d: float = 0.01,
                 use_jit: bool = True, integrator: str = "rk4", dtype: str = "float64"):
        self.n_x = 2
        self.n_u = 1
        self.g = g
        self.l = l
        self.d = d
        self.x_target = x_target
        self.Q = Q
        self.R = R
        self.Q_f = Q_f
        super().__init__(self.n_x, self.n_u, dt, use_jit=use_jit, integrator=integrator, dtype=dtype)

    def _device_model(self):
        return "pendulum", [self.g, self.l, self.d]
