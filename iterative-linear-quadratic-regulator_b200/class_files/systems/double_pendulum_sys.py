"""MyDoublePendulum -- drop-in for the reference's class_files/systems/double_pendulum_sys.py:9-205.

Fully actuated double pendulum, x = [q1, q2, q1_dot, q2_dot], u = [tau1, tau2]; q_ddot solves
M(q) q_ddot = h(q, q_dot, tau) (double_pendulum_sys.py:84-111, mass matrix :138-160, right-hand side
:162-206); quadratic stage/terminal cost (:114-134).  Runs as DoublePendulumSys<T,2> inside
libilqr_b200.so (csrc/ilqr_systems.cuh).
"""
from .system_base import System


class MyDoublePendulum(System):
    _MODEL = "double_pendulum"
    _N_U = 2

    def __init__(self, dt, x_target, Q, R, Q_f, g: float = 9.81, m1: float = 1.0, m2: float = 1.0,
                 l1: float = 1.0, l2: float = 1.0, d1: float = 0.01, d2: float = 0.01, theta1: float = 0.0,
                 theta2: float = 0.0, use_jit: bool = True, integrator: str = "rk4", dtype: str = "float64"):
        self.n_x = 4
        self.n_u = self._N_U
        self.g = g
        self.m1 = m1
        self.m2 = m2
        self.l1 = l1
        self.l2 = l2
        self.d1 = d1
        self.d2 = d2
        self.theta1 = theta1
        self.theta2 = theta2
        self.x_target = x_target
        self.Q = Q
        self.R = R
        self.Q_f = Q_f
        super().__init__(self.n_x, self.n_u, dt, use_jit=use_jit, integrator=integrator, dtype=dtype)

    def _device_model(self):
        return self._MODEL, [self.g, self.m1, self.m2, self.l1, self.l2, self.d1, self.d2, self.theta1, self.theta2]
